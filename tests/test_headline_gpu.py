"""Parity of the paths the headline numbers rest on (VERDICT r01 "weak" 1-2): the end-to-end host call
`dfb_render_image_host` that bench.py's `e2e` leg times, the full-size fp16 tensor-core render against the ORACLE (not
against the repo's own fp32 kernels), BASELINE config[4]'s shape, a trained-like (sharp) field, and the §8b shims
`render_path` / `create_nerf` with their file formats."""
import os
import types

import numpy as np
import pytest
import torch

from helpers import fit_synthetic_scene, np_params, rel_err, synthetic_nets
from oracle import nerf_oracle as O

pytestmark = pytest.mark.gpu
HIST = np.array([5, 10, 20, 30, 15, 10, 5, 3, 1, 1], np.float32)
C2W = np.array([[0.9962, -0.0872, 0.0, 0.0], [0.0872, 0.9962, 0.0, 0.0], [0.0, 0.0, 1.0, 1.0]], np.float32)


def dev():
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def ctx():
    from dfnet_b200 import ops
    mods, nets = synthetic_nets(8, 256)
    h = ops.NerfHandle(*[m.to(dev()) for m in mods])
    return ops, h, nets


def _oracle_subset(nets, H, W, focal, near, far, Nc, Nf, sel):
    o, d = O.get_rays(H, W, focal, C2W)
    rec = O.make_ray_records(o.reshape(-1, 3)[sel], d.reshape(-1, 3)[sel], near, far, HIST[None])
    O.set_linear_backend("torch")   # same arithmetic, threaded GEMMs: keeps 2 000 rays x 192 samples within seconds
    try:
        return O.render_rays(rec, nets, Nc, Nf, test_time=True)
    finally:
        O.set_linear_backend("numpy")


def test_render_image_host_is_the_device_path_and_matches_the_oracle(ctx):
    """a10: the render_path step (host pose in, host image out) at BASELINE config[1] size.  Its three host outputs must
    be bit-identical to dfb_render_fwd's device outputs (checks the staging offsets and the D2H copies) and a 2 000-ray
    subset must match the oracle at the north star's 1e-3."""
    from dfnet_b200 import _lib
    ops, h, nets = ctx
    H, W, focal, near, far, Nc, Nf = 480, 640, 585.0, 0.0, 2.5, 64, 128
    cfg = _lib.RenderCfg(N_samples=Nc, N_importance=Nf, test_time=1, perturb=0, mma_kind=_lib.MMA_KINDS["f16"],
                         lindisp=0, raw_noise_std=0.0)
    c2w_h, hist_h = torch.tensor(C2W).pin_memory(), torch.tensor(HIST).pin_memory()
    rgb_h, disp_h, acc_h = torch.empty(H * W, 3).pin_memory(), torch.empty(H * W).pin_memory(), torch.empty(H * W).pin_memory()
    rgb_h.fill_(-7.0), disp_h.fill_(-7.0), acc_h.fill_(-7.0)
    h.render_image_host(cfg, c2w_h, H, W, focal, near, far, hist_h, rgb_h, disp_h, acc_h, dev())
    torch.cuda.synchronize()
    o = h.render(Nc, Nf, True, c2w=torch.tensor(C2W, device=dev()), H=H, W=W, focal=focal, near=near, far=far,
                 hist=torch.tensor(HIST, device=dev()), mma="f16")
    torch.cuda.synchronize()
    assert torch.equal(rgb_h, o["rgb"].cpu()) and torch.equal(disp_h, o["disp"].cpu()) and torch.equal(acc_h, o["acc"].cpu())
    sel = np.arange(0, H * W, 153)[:2000]
    want = _oracle_subset(nets, H, W, focal, near, far, Nc, Nf, sel)
    e_rgb = rel_err(rgb_h.numpy()[sel], want["rgb_map"])
    e_acc = rel_err(acc_h.numpy()[sel], want["acc_map"])
    print("render_image_host vs oracle (2000 rays, f16): rgb", e_rgb, "acc", e_acc)
    assert e_rgb < 1e-3 and e_acc < 1e-3
    ok = want["disp_map"] < 1e3       # disp = 1/depth is ill-conditioned where depth -> 0
    assert ok.mean() > 0.9 and rel_err(disp_h.numpy()[sel][ok], want["disp_map"][ok]) < 2e-3
    # a histogram of the wrong length is refused before anything is read
    with pytest.raises(_lib.DfbError):
        h.render_image_host(cfg, c2w_h, H, W, focal, near, far, torch.zeros(3).pin_memory(), rgb_h, disp_h, acc_h, dev())


@pytest.mark.parametrize("name,H,W,focal,near,far,Nc,Nf,n", [
    ("cfg2", 480, 640, 585.0, 0.0, 2.5, 64, 128, 2000),
    ("cfg5", 1080, 1920, 1674.0, 0.0, 20.0, 64, 192, 800)])
def test_full_size_tensor_core_render_vs_oracle(ctx, name, H, W, focal, near, far, Nc, Nf, n):
    """The fp16 tcgen05 render of a whole image (BASELINE config[1] and config[4] shapes) against the oracle on a ray
    subset spread over the image, both cta_group variants."""
    ops, h, nets = ctx
    sel = np.linspace(0, H * W - 1, n).astype(np.int64)
    want = _oracle_subset(nets, H, W, focal, near, far, Nc, Nf, sel)
    for cg in ("2", "1"):
        os.environ["DFB_TC_CTA_GROUP"] = cg
        try:
            o = h.render(Nc, Nf, True, c2w=torch.tensor(C2W, device=dev()), H=H, W=W, focal=focal, near=near, far=far,
                         hist=torch.tensor(HIST, device=dev()), mma="f16")
            torch.cuda.synchronize()
        finally:
            os.environ.pop("DFB_TC_CTA_GROUP", None)
        e_rgb = rel_err(o["rgb"].cpu().numpy()[sel], want["rgb_map"])
        e_acc = rel_err(o["acc"].cpu().numpy()[sel], want["acc_map"])
        print(name, "cta_group", cg, "f16 vs oracle: rgb", e_rgb, "acc", e_acc)
        assert e_rgb < 1e-3 and e_acc < 1e-3, (name, cg)


def test_native_128_wide_program_vs_oracle_and_embedding(monkeypatch):
    """The reference's default networks (netwidth 128, 64+64 samples, models/options.py:31-33,56-57) at 640x480: the native
    128-wide tcgen05 program against the ORACLE at the north star's 1e-3 on a 2 000-ray subset, against the same network
    zero-padded into the 8x256 program (DFB_TC_NATIVE128=0: same function, different MMA shapes), with early ray
    termination, and on a ragged ray count."""
    from dfnet_b200 import ops
    mods, nets = synthetic_nets(8, 128)
    h = ops.NerfHandle(*[m.to(dev()) for m in mods])
    H, W, focal, near, far, Nc, Nf = 480, 640, 585.0, 0.0, 2.5, 64, 64
    kw = dict(c2w=torch.tensor(C2W, device=dev()), H=H, W=W, focal=focal, near=near, far=far, hist=torch.tensor(HIST, device=dev()))
    l0 = ops.lib.dfb_launch_count()
    nat = {k: v.clone() for k, v in h.render(Nc, Nf, True, mma="f16", **kw).items()}
    torch.cuda.synchronize()
    assert ops.lib.dfb_launch_count() > l0
    monkeypatch.setenv("DFB_TC_NATIVE128", "0")
    emb = {k: v.clone() for k, v in h.render(Nc, Nf, True, mma="f16", **kw).items()}
    torch.cuda.synchronize()
    monkeypatch.delenv("DFB_TC_NATIVE128")
    sel = np.linspace(0, H * W - 1, 2000).astype(np.int64)
    want = _oracle_subset(nets, H, W, focal, near, far, Nc, Nf, sel)
    for name, o in (("native", nat), ("embedded", emb)):
        e_rgb = rel_err(o["rgb"].cpu().numpy()[sel], want["rgb_map"])
        e_acc = rel_err(o["acc"].cpu().numpy()[sel], want["acc_map"])
        print(name, "128-wide f16 vs oracle: rgb", e_rgb, "acc", e_acc)
        assert e_rgb < 1e-3 and e_acc < 1e-3, name
    assert rel_err(nat["rgb"].cpu().numpy(), emb["rgb"].cpu().numpy()) < 1e-3
    # bf16 operands and the extras path (raw, depth: unfused compositing) on a ragged ray count
    o, d = O.get_rays(H, W, focal, C2W)
    idx = np.arange(0, H * W, 307)[:997]
    rec = torch.tensor(O.make_ray_records(o.reshape(-1, 3)[idx], d.reshape(-1, 3)[idx], near, far, HIST[None]), device=dev())
    ref = h.render(Nc, Nf, True, rays=rec, mma="fp32", want=("raw", "depth"))
    for mma, tol in (("f16", 1e-3), ("bf16", 1e-2)):
        got = h.render(Nc, Nf, True, rays=rec, mma=mma, want=("raw", "depth"))
        for k in ("rgb", "acc", "depth"):
            assert rel_err(got[k].cpu().numpy(), ref[k].cpu().numpy()) < tol, (mma, k)
    # opt-in early ray termination runs on the same program
    ert = h.render(Nc, Nf, True, mma="f16", ert_eps=1e-3, **kw)
    assert float((ert["rgb"] - nat["rgb"]).abs().max()) < 5e-3


def test_pair_kernels_are_deterministic_over_repeats(ctx):
    """The cta_group::2 kernels hand tiles between the two CTAs of a pair through barriers with CTA-scope release / acquire
    (tc_common.cuh, mbar_arrive_cluster): a missing ordering would show up as run-to-run differences.  30 renders of the
    640x480 image (fused compositing), 10 training forwards with masks + saved-mask backwards: bit-identical every time."""
    ops, h, nets = ctx
    H, W, focal, near, far, Nc, Nf = 480, 640, 585.0, 0.0, 2.5, 64, 128
    kw = dict(c2w=torch.tensor(C2W, device=dev()), H=H, W=W, focal=focal, near=near, far=far, hist=torch.tensor(HIST, device=dev()))
    first = {k: v.clone() for k, v in h.render(Nc, Nf, True, mma="f16", **kw).items()}
    for _ in range(30):
        o = h.render(Nc, Nf, True, mma="f16", **kw)
        for k in ("rgb", "disp", "acc"):
            assert torch.equal(o[k], first[k]), k
    oo, dd = O.get_rays(H, W, focal, C2W)
    idx = np.arange(0, H * W, 41)[:4099]      # ragged: not a multiple of the 128-sample tile
    rec = torch.tensor(O.make_ray_records(oo.reshape(-1, 3)[idx], dd.reshape(-1, 3)[idx], near, far, HIST[None]), device=dev())
    g = torch.randn(rec.shape[0], 3, device=dev()) * 1e-6
    ref = None
    for _ in range(10):
        t = h.render(Nc, Nf, True, rays=rec, mma="f16", want=("z_vals", "raw", "relu_masks"))
        grads = h.render_backward(rec, t["z_vals"], t["raw"], g, mma="f16", relu_masks=t["relu_masks"])
        m = t["relu_masks"].view(-1, 12, 8, 128)       # the 128-wide transient layers (9..11) write 4 of their 8 words
        cur = [t["raw"].clone(), torch.cat([m[:, :9].reshape(-1), m[:, 9:, :4].reshape(-1)])] + [x.clone() for x in grads]
        if ref is None:
            ref = cur
        for name, a, b in zip(("raw", "relu_masks", "g_rays_o", "g_rays_d", "g_viewdirs"), cur, ref):
            assert torch.equal(a, b), (name, int((a != b).sum()), a.numel())


def test_split_precision_coarse_pass_vs_oracle(ctx):
    """mma="f16s" on the benchmark field: the coarse weights (which decide the sample indices) match the oracle like the
    fp32 kernels do, indices almost never flip, the image stays inside 1e-3."""
    ops, h, nets = ctx
    H, W, focal, near, far, Nc, Nf = 480, 640, 585.0, 0.0, 2.5, 64, 128
    sel = np.linspace(0, H * W - 1, 1500).astype(np.int64)
    o, d = O.get_rays(H, W, focal, C2W)
    rec_np = O.make_ray_records(o.reshape(-1, 3)[sel], d.reshape(-1, 3)[sel], near, far, HIST[None])
    O.set_linear_backend("torch")
    try:
        want = O.render_rays(rec_np, nets, Nc, Nf, test_time=True, return_internals=True)
    finally:
        O.set_linear_backend("numpy")
    rec = torch.tensor(rec_np, device=dev())
    res = {}
    for mma in ("f16s", "f16"):
        g = h.render(Nc, Nf, True, rays=rec, mma=mma, want=("weights_coarse", "inds"))
        dw = float(np.abs(g["weights_coarse"].cpu().numpy() - want["_internals"]["weights_coarse"]).max())
        flips = float((g["inds"].cpu().numpy() != want["_internals"]["inds"]).mean())
        e = rel_err(g["rgb"].cpu().numpy(), want["rgb_map"])
        res[mma] = (dw, flips, e)
        print(mma, "coarse weights max abs err", dw, "index flip rate", flips, "rgb rel err", e)
        assert e < 1e-3
    assert res["f16s"][0] < 2e-6 and res["f16s"][1] < 2e-3
    assert res["f16s"][0] < 0.05 * res["f16"][0]


def test_ray_record_width_is_checked(ctx):
    """a record that is not [.., 11 + hist_bin] wide (e.g. a stray img_idx column, or the reference's default
    empty img_idx) must raise, not be read misaligned."""
    from dfnet_b200._lib import DfbError
    ops, h, _ = ctx
    bad = torch.zeros(16, 22, device=dev())
    with pytest.raises(DfbError):
        h.render(64, 128, True, rays=bad, mma="f16")
    with pytest.raises(DfbError):
        h.render(64, 128, True, c2w=torch.tensor(C2W, device=dev()), H=2, W=2, focal=1.0, hist=torch.zeros(0), mma="f16")
    with pytest.raises(DfbError):
        h.render_backward(bad, torch.zeros(16, 192, device=dev()), torch.zeros(16, 192, 9, device=dev()),
                          torch.zeros(16, 3, device=dev()), mma="f16")


@pytest.fixture(scope="module")
def fitted():
    """Trained-like field: sharp density steps (sigma 0 / 40), striped colours; fitted on the GPU with plain torch."""
    c, f, ea, et = fit_synthetic_scene(steps=400, batch=16384, device=dev())
    nets = dict(coarse=np_params(c), fine=np_params(f), emb_a=ea.weight.detach().cpu().numpy(),
                emb_t=et.weight.detach().cpu().numpy(), D=8, skips=(4,), beta_min=0.1)
    return (c, f, ea, et), nets


def test_trained_like_field_precision_ladder(fitted):
    """VERDICT r01 weak #2c.  A field with structure (the seeded gain-1.6 field renders rgb in [0.51, 0.55]): rendered
    rgb spans ~[0.1, 0.95].  Gates:
      * hidden activations stay far inside fp16's range (max |pre-activation| recorded from the oracle);
      * fp32 kernels: <= 1e-3 of the oracle on every ray, <= 1e-4 on 99 % of them (hierarchical sampling is itself
        ill-conditioned on a sharp field: a 1e-6 difference in a coarse weight moves fine samples, and the highest
        positional-encoding band turns a 1e-6 depth shift into a 1e-3 phase shift; measured max 3.6e-4, mean 1.7e-6);
      * "f16s" (split-precision coarse pass + fp16 fine pass): coarse weights within 1.2e-5 of the oracle (fp16: 4.9e-3);
        what remains is the fine network's own operand rounding: measured mean 4.9e-4, p99 1.5e-3, max 3.0e-3
        (gated at 1e-3 / 3e-3 / 1e-2);
      * "f16" (everything fp16): the coarse pass' operand rounding (2^-11) shifts the coarse weights by ~5e-3, which
        moves fine samples across the density steps: mean ~1e-3, ~15 % of the rays above 1e-3, worst ray ~9e-2
        (measured, printed, loosely gated).  An oracle emulation of the roundings attributes > 90 % of that to the coarse
        network alone, which is why the split-precision kind exists."""
    from dfnet_b200 import ops
    mods, nets = fitted
    H, W, focal, near, far, Nc, Nf = 480, 640, 585.0, 0.0, 2.5, 64, 128
    sel = np.linspace(0, H * W - 1, 1500).astype(np.int64)
    amax = [0.0]
    lin0 = O._linear

    def lin_track(x, w, b):
        y = lin0(x, w, b)
        amax[0] = max(amax[0], float(np.abs(y).max()))
        return y
    O._linear = lin_track
    try:
        want = _oracle_subset(nets, H, W, focal, near, far, Nc, Nf, sel)
    finally:
        O._linear = lin0
    print("trained-like field: max |pre-activation| =", amax[0], " rgb range", want["rgb_map"].min(), want["rgb_map"].max())
    assert amax[0] < 65504 / 16
    assert want["rgb_map"].max() - want["rgb_map"].min() > 0.5      # the field has structure
    h = ops.NerfHandle(*mods)
    o, d = O.get_rays(H, W, focal, C2W)
    rec = torch.tensor(O.make_ray_records(o.reshape(-1, 3)[sel], d.reshape(-1, 3)[sel], near, far, HIST[None]), device=dev())
    got32 = h.render(Nc, Nf, True, rays=rec, mma="fp32")
    r32 = (np.abs(got32["rgb"].cpu().numpy() - want["rgb_map"]) / np.maximum(np.abs(want["rgb_map"]), 1e-3)).max(1)
    print(f"fp32 kernels vs oracle on the trained-like field: mean {r32.mean():.2e} p99 {np.percentile(r32, 99):.2e} max {r32.max():.2e}")
    assert r32.max() < 1e-3 and np.percentile(r32, 99) < 1e-4
    for mma in ("f16s", "f16", "bf16"):
        got = h.render(Nc, Nf, True, rays=rec, mma=mma, want=("weights_coarse",))
        r = np.abs(got["rgb"].cpu().numpy() - want["rgb_map"]) / np.maximum(np.abs(want["rgb_map"]), 1e-3)
        per_ray = r.max(1)
        print(f"{mma} tensor-core kernels vs oracle on the trained-like field: mean {per_ray.mean():.2e} p99 "
              f"{np.percentile(per_ray, 99):.2e} max {per_ray.max():.2e} frac>1e-3 {(per_ray > 1e-3).mean():.3f}")
        assert np.isfinite(r).all()
        if mma == "f16s":    # split-precision coarse pass + fp16 fine pass: the sample placement is fp32-exact
            assert per_ray.mean() < 1e-3 and np.percentile(per_ray, 99) < 3e-3 and per_ray.max() < 1e-2
        if mma == "f16":     # everything fp16: the coarse pass' rounding moves samples across surfaces
            assert per_ray.mean() < 3e-3 and np.percentile(per_ray, 99) < 5e-2


def test_render_path_shim_writes_the_reference_files(ctx, tmp_path):
    """§8b: render_path on top of dfb_render_image_host - same return values as per-image render() calls, the
    reference's file names, 8-bit PNGs that decode to to8b(rgb)."""
    import cv2
    from dfnet_b200 import rendering
    from dfnet_b200.nerfw import to8b
    mods, _ = synthetic_nets(8, 256)
    c, f, ea, et = [m.to(dev()) for m in mods]
    kw = dict(network_query_fn=None, perturb=False, N_importance=32, network_fine=f, N_samples=16, network_fn=c,
              use_viewdirs=True, white_bkgd=False, raw_noise_std=0.0, embedding_a=ea, embedding_t=et, test_time=True,
              ndc=False, lindisp=False, near=0.0, far=2.5)
    H, W, focal = 24, 32, 30.0
    poses = torch.tensor(np.stack([np.concatenate([C2W, [[0, 0, 0, 1]]], 0)] * 3).astype(np.float32), device=dev())
    poses[1, 0, 3] = 0.1
    poses[2, 1, 3] = -0.1
    hists = torch.tensor(np.stack([HIST, HIST[::-1].copy(), HIST]), device=dev())
    gt = np.random.RandomState(0).rand(3, H, W, 3).astype(np.float32)
    args = types.SimpleNamespace()
    rgbs, disps = rendering.render_path(args, poses, (H, W, focal), 32768, kw, gt_imgs=gt, savedir=str(tmp_path), img_ids=hists)
    assert rgbs.shape == (3, H, W, 3) and disps.shape == (3, H, W)
    for i in range(3):
        rgb, disp, acc, _ = rendering.render(H, W, focal, c2w=poses[i, :3, :4], img_idx=hists[i], **kw)
        assert np.array_equal(rgbs[i], rgb.cpu().numpy()) and np.array_equal(disps[i], disp.cpu().numpy())
        png = cv2.imread(str(tmp_path / f"{i:03d}.png"))[..., ::-1]
        assert np.array_equal(png, to8b(rgbs[i]))
        assert np.array_equal(cv2.imread(str(tmp_path / f"{i:03d}_GT.png"))[..., ::-1], to8b(gt[i]))
        assert np.array_equal(cv2.imread(str(tmp_path / f"{i:03d}_disp.png"), cv2.IMREAD_UNCHANGED),
                              to8b(disps[i] / np.max(disps[i])))
    # render_factor and the general (non test-time) branch
    kw_train = dict(kw, test_time=False)
    r2, d2 = rendering.render_path(args, poses[:1], (H, W, focal), 32768, kw_train, render_factor=2, img_ids=hists, mma="fp32")
    assert r2.shape == (1, H // 2, W // 2, 3) and np.isfinite(r2).all()


def test_create_nerf_and_tar_checkpoint_roundtrip(tmp_path):
    """§8b: create_nerf(args) -> kwargs / optimizer; the .tar checkpoint dict of run_nerf.py:150-167 written by
    save_checkpoint is found (newest *tar* file of basedir/expname) and reloaded bit for bit."""
    from dfnet_b200 import nerfw, rendering
    exp = tmp_path / "exp"
    exp.mkdir()
    args = types.SimpleNamespace(NeRFH=True, encode_hist=True, multires=10, multires_views=4, use_viewdirs=True, i_embed=0,
                                 reduce_embedding=-1, N_vocab=1000, netdepth=8, netwidth=128, N_importance=64, N_samples=64,
                                 in_channels_a=50, in_channels_t=20, no_grad_update=False, lrate=5e-4, basedir=str(tmp_path),
                                 expname="exp", ft_path=None, no_reload=False, perturb=1.0, white_bkgd=False,
                                 raw_noise_std=0.0, dataset_type="7Scenes", no_ndc=True, lindisp=False, multi_gpu=False)
    ktr, kte, start, grad_vars, opt = nerfw.create_nerf(args)
    assert start == 0 and isinstance(opt, torch.optim.Adam) and len(grad_vars) == 24 + 38 + 2
    assert ktr["test_time"] is False and kte["test_time"] is True and kte["perturb"] is False and ktr["perturb"] == 1.0
    assert ktr["ndc"] is False and next(ktr["network_fn"].parameters()).is_cuda
    with torch.no_grad():   # make the four modules distinguishable from a fresh construction
        for k in ("network_fn", "network_fine", "embedding_a", "embedding_t"):
            for p in ktr[k].parameters():
                p.add_(torch.randn_like(p) * 0.01)
    nerfw.save_checkpoint(str(exp / "000100.tar"), 1234, ktr, opt)
    nerfw.save_checkpoint(str(exp / "000050.tar"), 50, {"network_fn": nerfw.NeRFW("coarse", W=128),
                                                        "network_fine": None}, None)   # older: must not be picked
    ktr2, kte2, start2, _, _ = nerfw.create_nerf(args)
    assert start2 == 1234
    for k in ("network_fn", "network_fine", "embedding_a", "embedding_t"):
        a, b = ktr[k].state_dict(), ktr2[k].state_dict()
        assert list(a) == list(b) and all(torch.equal(a[n], b[n]) for n in a)
    # the reloaded networks render the same image
    img = [rendering.render(12, 16, 15.0, c2w=torch.tensor(C2W, device=dev()), img_idx=torch.tensor(HIST, device=dev()),
                            near=0.0, far=2.5, **kk)[0] for kk in (kte, kte2)]
    assert torch.equal(img[0], img[1])
    args.no_grad_update = True
    assert nerfw.create_nerf(args)[3:] == (None, None)


@pytest.mark.parametrize("n_rays,Nc,Nf", [(5000, 64, 128), (777, 16, 24), (1301, 64, 192), (300, 64, 64), (97, 20, 13), (70000, 64, 128)])
@pytest.mark.parametrize("cg", ["2", "1"])
def test_fused_compositing_matches_the_raw_round_trip(ctx, n_rays, Nc, Nf, cg):
    """N1: compositing fused into the heads epilogue of the fine tcgen05 kernel (per-warp segment records + k_composite_partials)
    against the unfused path (raw [N,S,9] through HBM + k_composite_fine_tt, float64 scans) on the SAME network outputs:
    sample counts where rays are warp-aligned (192, 256, 128) and ragged ones (40, 33), ray counts that leave partial
    tiles, more rays than one internal chunk.  fp32 association differs, nothing else: <= 2e-5 relative."""
    ops, h, _ = ctx
    if n_rays > 20000 and cg == "1":
        pytest.skip("large case once")
    rng = np.random.RandomState(n_rays)
    o = np.tile(np.array([[0.0, 0.0, 1.0]], np.float32), (n_rays, 1))
    d = (rng.randn(n_rays, 3) * 0.3 + np.array([0, 0, -1.0])).astype(np.float32)
    rec = torch.tensor(O.make_ray_records(o, d, 0.0, 2.5, HIST[None]), device=dev())
    out = {}
    for fuse in ("1", "0"):
        os.environ["DFB_TC_FUSE_COMPOSITE"] = fuse
        os.environ["DFB_TC_CTA_GROUP"] = cg
        try:
            r = h.render(Nc, Nf, True, rays=rec, mma="f16")
            out[fuse] = {k: v.cpu().numpy() for k, v in r.items()}
        finally:
            os.environ.pop("DFB_TC_FUSE_COMPOSITE", None), os.environ.pop("DFB_TC_CTA_GROUP", None)
    for k in ("rgb", "acc"):
        assert rel_err(out["1"][k], out["0"][k]) < 2e-5, k
    ok = out["0"]["disp"] < 1e3
    assert rel_err(out["1"]["disp"][ok], out["0"]["disp"][ok]) < 1e-4
    assert np.isfinite(out["1"]["disp"]).all()


def test_early_ray_termination_opt_in(ctx, fitted):
    """N1: opt-in early ray termination (ert_eps): fine samples behind the depth where the COARSE transmittance falls
    below ert_eps are not evaluated (compacted sample list, rays keep a prefix of their sorted samples).
      * ert_eps -> tiny (nothing terminates): n_live == S everywhere and the image equals the parity path's to fp32 noise;
      * on a field with surfaces a large share of the samples is skipped and the image moves by at most ~ert_eps;
      * the option is refused where it cannot apply (train mode, fp32 kernels)."""
    from dfnet_b200 import ops
    from dfnet_b200._lib import DfbError
    _, h_smooth, _ = ctx
    mods, _ = fitted
    h_sharp = ops.NerfHandle(*mods)
    H, W, focal, near, far, Nc, Nf = 120, 160, 146.0, 0.0, 2.5, 64, 128
    c2w, hist = torch.tensor(C2W, device=dev()), torch.tensor(HIST, device=dev())
    for name, h in (("smooth", h_smooth), ("sharp", h_sharp)):
        base = h.render(Nc, Nf, True, c2w=c2w, H=H, W=W, focal=focal, near=near, far=far, hist=hist, mma="f16")
        rgb0, acc0, disp0 = base["rgb"].clone(), base["acc"].clone(), base["disp"].clone()
        off = h.render(Nc, Nf, True, c2w=c2w, H=H, W=W, focal=focal, near=near, far=far, hist=hist, mma="f16", ert_eps=1e-30,
                       want=("n_live",))
        if name == "smooth":      # (on the sharp field the coarse transmittance reaches exactly 0 in fp32 behind a surface)
            assert int(off["n_live"].min()) == Nc + Nf
        assert rel_err(off["rgb"].cpu().numpy(), rgb0.cpu().numpy()) < 2e-5
        for eps in (1e-2, 1e-3):
            o = h.render(Nc, Nf, True, c2w=c2w, H=H, W=W, focal=focal, near=near, far=far, hist=hist, mma="f16", ert_eps=eps,
                         want=("n_live",))
            frac = float(o["n_live"].float().mean()) / (Nc + Nf)
            d_rgb = float((o["rgb"] - rgb0).abs().max())
            m_rgb = float((o["rgb"] - rgb0).abs().mean())
            d_acc = float((o["acc"] - acc0).abs().max())
            print(f"ERT {name} eps={eps:g}: evaluated {100 * frac:.1f} % of the fine samples, |d rgb| mean {m_rgb:.2e} max {d_rgb:.2e}, "
                  f"max |d acc| {d_acc:.2e}")
            assert int(o["n_live"].min()) >= 1 and int(o["n_live"].max()) <= Nc + Nf
            # the criterion is the COARSE network's transmittance: where the two networks disagree about a surface the
            # dropped tail still carries fine-network weight, so the bound is statistical (mean <= eps), not per ray
            assert m_rgb < eps and d_rgb < 0.1 and d_acc < 0.25
            assert torch.isfinite(o["disp"]).all()
            if name == "sharp":       # hierarchical sampling already puts 2/3 of the samples at the surface: the dead tail is
                assert frac < 0.99    # only the coarse-grid samples behind it (measured 93.7 % / 94.7 % evaluated)
    with pytest.raises(DfbError):
        h_smooth.render(Nc, Nf, False, c2w=c2w, H=8, W=8, focal=8.0, near=near, far=far, hist=hist, mma="f16", ert_eps=1e-3)
    with pytest.raises(DfbError):
        h_smooth.render(Nc, Nf, True, c2w=c2w, H=8, W=8, focal=8.0, near=near, far=far, hist=hist, mma="fp32", ert_eps=1e-3)
