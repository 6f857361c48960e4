"""Golden vectors for the DFNet feature path from the UNMODIFIED reference (/root/reference).

    python tests/golden/make_golden_dfnet.py   ->  tests/golden/dfnet_golden.npz

VGG-16 weights are seeded random (no pretrained weights offline); they are not stored: tests rebuild
them with tests/helpers.synthetic_dfnet (torchvision vgg16(weights=None) under the same seed) and
re-check the SHA-256 recorded here."""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
for _m in ["imageio", "matplotlib", "matplotlib.pyplot", "pytorch3d", "pytorch3d.transforms", "efficientnet_pytorch",
           "torchsummary", "kornia", "transforms3d", "transforms3d.euler", "transforms3d.quaternions", "pykalman",
           "configargparse"]:
    sys.modules.setdefault(_m, types.ModuleType(_m))
sys.modules["efficientnet_pytorch"].EfficientNet = object
sys.modules["torchsummary"].summary = lambda *a, **k: None
sys.path[:0] = ["/root/reference/script", "/root/reference"]

import torch  # noqa: E402
import torchvision  # noqa: E402

_orig_vgg16 = torchvision.models.vgg16
torchvision.models.vgg16 = lambda pretrained=False, **kw: _orig_vgg16(weights=None)

from feature import dfnet as ref_dfnet  # noqa: E402
from helpers import sd_checksum, synthetic_dfnet  # noqa: E402

torch.set_num_threads(8)
G = {}


def put(k, v):
    G[k] = np.ascontiguousarray(v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else v)


def main():
    for tag, cls, mycls in (("dfnet", ref_dfnet.DFNet, "DFNet"), ("dfnet_s", ref_dfnet.DFNet_s, "DFNet_s")):
        torch.manual_seed(0)
        ref = cls().eval()
        mine = synthetic_dfnet(mycls)
        sa, sb = ref.state_dict(), mine.state_dict()
        assert list(sa) == list(sb)
        for k in sa:
            assert torch.equal(sa[k], sb[k]), k
        put(f"{tag}_sha", np.frombuffer(sd_checksum(sa).encode(), np.uint8))
        rng = np.random.RandomState(5)
        x = torch.from_numpy(rng.rand(2, 3, 48, 64).astype(np.float32))
        put(f"{tag}_x", x)
        with torch.no_grad():
            feats, pose = ref(x, return_feature=True, isSingleStream=False, return_pose=True, upsampleH=48, upsampleW=64)
            feats_s, _ = ref(x, return_feature=True, isSingleStream=True, return_pose=False, upsampleH=30, upsampleW=40)
            _, pose_only = ref(x, return_feature=False)
        put(f"{tag}_pose", pose)
        assert torch.equal(pose, pose_only)
        for nm, f in (("t", feats[0]), ("r", feats[1]), ("s", feats_s[0])):
            put(f"{tag}_feat_{nm}_sub", f[:, :, ::8, ::4, ::4])
            put(f"{tag}_feat_{nm}_stats", torch.stack([f.sum(), f.abs().sum(), f.abs().max(), (f * f).sum()]).double())
        # train-mode BatchNorm in the heads (run_feature.py without freezeBN): batch statistics over the whole batch,
        # running statistics updated in place.  Non-trivial BatchNorm state so that the fold is visible.
        ref_t = cls()
        ref_t.load_state_dict(ref.state_dict())
        g = torch.Generator().manual_seed(77)
        bn_state = {}
        for l in range(3 if tag == "dfnet" else 1):
            bn = getattr(ref_t.adaptation_layers, f"adapt_layer_{l}")[3]
            with torch.no_grad():
                bn.weight.copy_(0.5 + torch.rand(128, generator=g))
                bn.bias.copy_(0.2 * torch.randn(128, generator=g))
                bn.running_mean.copy_(0.1 * torch.randn(128, generator=g))
                bn.running_var.copy_(0.5 + torch.rand(128, generator=g))
            for k in ("weight", "bias", "running_mean", "running_var"):
                bn_state[f"{l}.{k}"] = getattr(bn, k).detach().clone()
                put(f"{tag}_bntrain_init_{l}_{k}", bn_state[f"{l}.{k}"])
        ref_t.train()
        with torch.no_grad():
            feats_bn, _ = ref_t(x, return_feature=True, isSingleStream=False, return_pose=False, upsampleH=48, upsampleW=64)
        for nm, f in (("t", feats_bn[0]), ("r", feats_bn[1])):
            put(f"{tag}_bntrain_feat_{nm}_sub", f[:, :, ::8, ::4, ::4])
            put(f"{tag}_bntrain_feat_{nm}_stats", torch.stack([f.sum(), f.abs().sum(), f.abs().max(), (f * f).sum()]).double())
        for l in range(3 if tag == "dfnet" else 1):
            bn = getattr(ref_t.adaptation_layers, f"adapt_layer_{l}")[3]
            put(f"{tag}_bntrain_running_mean_{l}", bn.running_mean)
            put(f"{tag}_bntrain_running_var_{l}", bn.running_var)
            assert int(bn.num_batches_tracked) == 1
        if tag == "dfnet":
            from feature.direct_feature_matching import feature_loss, preprocess_features_for_loss
            ft = preprocess_features_for_loss(feats[0])[0]
            fr = preprocess_features_for_loss(feats[1])[0]
            put("loss_per_channel_false", feature_loss(fr, ft, per_channel=False))
            put("loss_per_channel_true", feature_loss(fr, ft, per_channel=True))
            put("loss_lvl0_false", feature_loss(fr[:128], ft[:128], per_channel=False))
    # triplet loss with in-triplet hard negative mining (feature/misc.py:399-435), all four cases
    from feature.misc import triplet_loss_hard_negative_mining_plus as ref_triplet
    rng = np.random.RandomState(11)
    base = rng.randn(3, 4, 8, 6, 10).astype(np.float32)
    variants = {
        0: (base, np.roll(base, -1, 1) + 0.01 * rng.randn(*base.shape).astype(np.float32)),   # f1 ~ roll(f2): case 0
        1: (np.roll(base, -1, 1) + 0.01 * rng.randn(*base.shape).astype(np.float32), base),   # f2 ~ roll(f1): case 1
        2: (np.broadcast_to(base[:, :1], base.shape).copy() + 0.01 * rng.randn(*base.shape).astype(np.float32),
            rng.randn(*base.shape).astype(np.float32)),                                       # f1 constant over b: case 2
        3: (rng.randn(*base.shape).astype(np.float32),
            np.broadcast_to(base[:, :1], base.shape).copy() + 0.01 * rng.randn(*base.shape).astype(np.float32)),
    }
    for k, (a, b) in variants.items():
        put(f"trip_{k}_f1", a), put(f"trip_{k}_f2", b)
        put(f"trip_{k}_loss", ref_triplet(torch.from_numpy(a), torch.from_numpy(b), margin=1.0))
    out = os.path.join(HERE, "dfnet_golden.npz")
    np.savez_compressed(out, **G)
    print("wrote", out, os.path.getsize(out) / 1e3, "KB")


if __name__ == "__main__":
    main()
