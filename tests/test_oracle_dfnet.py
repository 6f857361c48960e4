"""Pin oracle/dfnet_oracle.py against vectors from the unmodified reference DFNet (CPU only)."""
import numpy as np
import pytest

from helpers import sd_checksum, synthetic_dfnet
from oracle import dfnet_oracle as DO


@pytest.fixture(scope="module")
def g():
    import os
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "dfnet_golden.npz"))


def relmax(a, b):
    return float(np.abs(np.asarray(a, np.float64) - b).max() / max(np.abs(b).max(), 1e-12))


@pytest.mark.parametrize("tag,cls,L", [("dfnet", "DFNet", 3), ("dfnet_s", "DFNet_s", 1)])
def test_dfnet_forward_oracle(g, tag, cls, L):
    net = synthetic_dfnet(cls)
    assert sd_checksum(net.state_dict()) == bytes(g[f"{tag}_sha"]).decode()
    P = {k: v.numpy() for k, v in net.state_dict().items()}
    x = g[f"{tag}_x"]
    feats, pose = DO.dfnet_forward(P, x, n_levels=L, single=False, return_pose=True, upH=48, upW=64)
    assert relmax(pose, g[f"{tag}_pose"]) < 1e-4
    for nm, f in (("t", feats[0]), ("r", feats[1])):
        assert f.shape == (L, 1, 128, 48, 64)
        assert relmax(f[:, :, ::8, ::4, ::4], g[f"{tag}_feat_{nm}_sub"]) < 1e-4
        st = g[f"{tag}_feat_{nm}_stats"]
        assert abs(np.abs(f).sum(dtype=np.float64) - st[1]) / st[1] < 1e-4
    fs, none = DO.dfnet_forward(P, x, n_levels=L, single=True, return_pose=False, upH=30, upW=40)
    assert none is None and relmax(fs[0][:, :, ::8, ::4, ::4], g[f"{tag}_feat_s_sub"]) < 1e-4
    if tag == "dfnet":
        ft = feats[0].transpose(1, 0, 2, 3, 4).reshape(1, 384, 48, 64)[0]
        fr = feats[1].transpose(1, 0, 2, 3, 4).reshape(1, 384, 48, 64)[0]
        assert abs(DO.feature_loss(fr, ft, False) - g["loss_per_channel_false"]) < 2e-5
        assert abs(DO.feature_loss(fr, ft, True) - g["loss_per_channel_true"]) < 2e-5
        assert abs(DO.feature_loss(fr[:128], ft[:128], False) - g["loss_lvl0_false"]) < 2e-5


@pytest.mark.parametrize("k", [0, 1, 2, 3])
def test_triplet_loss_oracle(g, k):
    loss, case = DO.triplet_loss_hnm_plus(g[f"trip_{k}_f1"], g[f"trip_{k}_f2"], 1.0)
    assert case == k                      # the four constructed inputs exercise the four cases
    assert abs(float(loss) - float(np.asarray(g[f"trip_{k}_loss"]).reshape(-1)[0])) < 1e-5


@pytest.mark.parametrize("tag,cls,L", [("dfnet", "DFNet", 3), ("dfnet_s", "DFNet_s", 1)])
def test_dfnet_forward_oracle_train_mode_batchnorm(g, tag, cls, L):
    """Heads under model.train() (run_feature.py without freezeBN): batch statistics over the whole (target + render)
    batch and the running-statistics update, against the unmodified reference."""
    net = synthetic_dfnet(cls)
    P = {k: v.numpy().copy() for k, v in net.state_dict().items()}
    for l in range(L):
        for k in ("weight", "bias", "running_mean", "running_var"):
            P[f"adaptation_layers.adapt_layer_{l}.3.{k}"] = g[f"{tag}_bntrain_init_{l}_{k}"]
    run = {}
    feats, _ = DO.dfnet_forward(P, g[f"{tag}_x"], n_levels=L, single=False, return_pose=False, upH=48, upW=64,
                                bn_train=True, bn_running=run)
    for nm, f in (("t", feats[0]), ("r", feats[1])):
        assert relmax(f[:, :, ::8, ::4, ::4], g[f"{tag}_bntrain_feat_{nm}_sub"]) < 1e-4
        st = g[f"{tag}_bntrain_feat_{nm}_stats"]
        assert abs(np.abs(f).sum(dtype=np.float64) - st[1]) / st[1] < 1e-4
    for l in range(L):
        p = f"adaptation_layers.adapt_layer_{l}.3."
        assert relmax(run[p + "running_mean"], g[f"{tag}_bntrain_running_mean_{l}"]) < 1e-5
        assert relmax(run[p + "running_var"], g[f"{tag}_bntrain_running_var_{l}"]) < 1e-5
