"""Drop-in proof (VERDICT r01 next #1d): the reference's OWN, unmodified callers - `feature.direct_feature_matching.
train_on_batch` and `models.rendering.render_path`, imported from baseline/_ref - run with only `render`, the DFNet
modules and `feature_loss` redirected to dfnet_b200, and reproduce (a) the golden step the reference produced on its own
stack (tests/golden/make_golden_train.py) and (b) dfnet_b200's own render_path.  Skipped when baseline/_ref was not
populated (python baseline/make_ref.py, run by __graft_entry__.build() wherever /root/reference exists)."""
import os
import types

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from baseline import ref_shims
from helpers import pose_head_init_, sd_checksum, synthetic_dfnet, synthetic_nets, train_case

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_shims.available(), reason="baseline/_ref not populated")]


def dev():
    return torch.device("cuda:0")


class _Recorder:
    def __init__(self, model):
        self.model, self.grads = model, None

    def step(self):
        self.grads = {n: p.grad.detach().clone() for n, p in self.model.named_parameters() if p.grad is not None}

    def zero_grad(self):
        self.model.zero_grad()


@pytest.mark.parametrize("case,mma", [("lvl0", "fp32"), ("lvl0", "f16"), ("lvl012", "f16")])
def test_reference_train_on_batch_unmodified_on_the_shims(monkeypatch, case, mma):
    ref_shims.activate(cpu_default_tensor_type=False)
    import feature.direct_feature_matching as RD      # the reference's module, unmodified
    from dfnet_b200 import dfnet as my_dfnet
    from dfnet_b200 import rendering as my_rendering
    assert os.path.realpath(RD.__file__).startswith(os.path.realpath(ref_shims.reference_root()))
    monkeypatch.setattr(RD, "render", my_rendering.render)
    monkeypatch.setattr(RD, "feature_loss", my_dfnet.feature_loss)
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "train_golden.npz"))
    cfg = train_case(case)
    Fnet = pose_head_init_(synthetic_dfnet("DFNet", seed=0))
    Gnet = synthetic_dfnet("DFNet", seed=1)
    assert sd_checksum(Fnet.state_dict()) == bytes(g[f"{case}_F_sha"]).decode()
    Fnet, Gnet = Fnet.to(dev()), Gnet.to(dev()).eval()
    for p in Gnet.parameters():
        p.requires_grad_(False)
    Fnet.train()
    Fnet = RD.freeze_bn_layer_train(Fnet)              # the reference's own helper on our module tree
    for m in Fnet.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.weight.requires_grad_(False), m.bias.requires_grad_(False)   # train.py:111-112
    mods, _ = synthetic_nets(cfg["D"], cfg["W"])
    c, f, ea, et = [m.to(dev()) for m in mods]
    for m in (c, f, ea, et):
        for p in m.parameters():
            p.requires_grad_(False)
    kw = dict(network_query_fn=None, perturb=0.0, N_importance=cfg["Nf"], network_fine=f, N_samples=cfg["Nc"], network_fn=c,
              use_viewdirs=True, white_bkgd=False, raw_noise_std=0.0, embedding_a=ea, embedding_t=et, test_time=True,
              ndc=False, lindisp=False, near=cfg["near"], far=cfg["far"], mma=mma)
    rec = _Recorder(Fnet)
    try:
        loss, psnr = RD.train_on_batch(cfg["args"], cfg["data"], Fnet, Gnet, cfg["pose"], cfg["hist"], cfg["hwf"], rec, True,
                                       dev(), cfg["world"], **kw)
    finally:
        torch.set_default_tensor_type("torch.FloatTensor")
    want_loss, want_psnr = float(g[f"{case}_loss"].reshape(-1)[0]), float(g[f"{case}_psnr"].reshape(-1)[0])
    tol = 1e-3
    assert abs(float(loss.reshape(-1)[0]) - want_loss) < tol * abs(want_loss), (loss, want_loss)
    assert abs(float(np.asarray(psnr).reshape(-1)[0]) - want_psnr) < tol * abs(want_psnr), (psnr, want_psnr)
    names = bytes(g[f"{case}_grad_names"]).decode().split("\n")
    assert sorted(rec.grads) == names
    worst_cos, worst_norm = (1.0, ""), (0.0, "")
    for n in names:
        gg = rec.grads[n].flatten()
        sub = gg[:: max(1, gg.numel() // 4096)][:4096].double().cpu()
        want = torch.from_numpy(g[f"{case}_g_{n}_sub"]).double()
        worst_cos = min(worst_cos, (float(F.cosine_similarity(sub, want, dim=0)), n))
        worst_norm = max(worst_norm, (abs(float(gg.norm()) / float(g[f"{case}_g_{n}_stats"][0]) - 1.0), n))
    print("reference train_on_batch on the shims:", case, mma, "loss", float(loss.reshape(-1)[0]), want_loss, "worst cos",
          worst_cos, "worst norm dev", worst_norm)
    assert worst_cos[0] > 0.95 and worst_norm[0] < 0.10, (worst_cos, worst_norm)


def test_reference_render_path_unmodified_on_the_shims(monkeypatch):
    ref_shims.activate(cpu_default_tensor_type=False)
    import models.rendering as RR
    from dfnet_b200 import rendering as my_rendering
    monkeypatch.setattr(RR, "render", my_rendering.render)
    mods, _ = synthetic_nets(8, 256)
    c, f, ea, et = [m.to(dev()) for m in mods]
    kw = dict(network_query_fn=None, perturb=False, N_importance=32, network_fine=f, N_samples=16, network_fn=c,
              use_viewdirs=True, white_bkgd=False, raw_noise_std=0.0, embedding_a=ea, embedding_t=et, test_time=True,
              ndc=False, lindisp=False, near=0.0, far=2.5)
    H, W, focal = 24, 32, 30.0
    c2w = np.array([[0.9962, -0.0872, 0.0, 0.0], [0.0872, 0.9962, 0.0, 0.0], [0.0, 0.0, 1.0, 1.0], [0, 0, 0, 1]], np.float32)
    poses = torch.tensor(np.stack([c2w] * 2), device=dev())
    poses[1, 0, 3] = 0.2
    hists = torch.tensor([[5., 10, 20, 30, 15, 10, 5, 3, 1, 1]] * 2, device=dev())
    gt = np.random.RandomState(0).rand(2, H, W, 3).astype(np.float32)
    args = types.SimpleNamespace()
    with torch.no_grad():
        rgbs_ref, disps_ref = RR.render_path(args, poses, (H, W, focal), 32768, kw, gt_imgs=gt, savedir=None, img_ids=hists)
    rgbs, disps = my_rendering.render_path(args, poses, (H, W, focal), 32768, kw, gt_imgs=gt, savedir=None, img_ids=hists)
    assert np.array_equal(rgbs, rgbs_ref) and np.array_equal(disps, disps_ref)
