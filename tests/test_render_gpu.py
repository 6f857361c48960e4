"""GPU parity tests of the render hot path, through the C ABI, against the oracle and the
reference-generated golden vectors.  Tolerances: bit-exact for indices and for op-level
float results whose operation order is pinned; 1e-3 relative (north star) for the tensor-core
paths; 1e-4 for the fp32 path."""
import numpy as np
import pytest
import torch

from helpers import rel_err, synthetic_nets
from oracle import nerf_oracle as O

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda:0")


def T(x):
    return torch.as_tensor(np.ascontiguousarray(x)).to(dev())


@pytest.fixture(scope="module")
def ops():
    from dfnet_b200 import ops as _ops
    return _ops


def to_dev(mods):
    return [m.to(dev()) if m is not None else None for m in mods]


def test_get_rays_bit_exact(ops, golden):
    H, W, f = golden["rays_hwf"]
    o, d = ops.get_rays(int(H), int(W), float(f), T(golden["rays_c2w"]))
    assert np.array_equal(o.cpu().numpy(), golden["rays_o"])
    assert np.array_equal(d.cpu().numpy(), golden["rays_d"])


def test_sample_pdf_bit_exact(ops, golden):
    s, i = ops.sample_pdf(T(golden["pdf_bins"]), T(golden["pdf_w"]), 128, det=True)
    assert np.array_equal(i.cpu().numpy(), golden["pdf_det_inds"])
    assert np.array_equal(s.cpu().numpy(), golden["pdf_det_samples"])
    s, i = ops.sample_pdf(T(golden["pdf_bins"]), T(golden["pdf_w"]), 128, det=False, u=T(golden["pdf_u_rand"]))
    assert np.array_equal(i.cpu().numpy(), golden["pdf_rand_inds"])
    assert np.array_equal(s.cpu().numpy(), golden["pdf_rand_samples"])


@pytest.mark.parametrize("nb,nf", [(63, 128), (63, 192), (15, 24), (127, 64), (600, 33), (2, 5)])
def test_sample_pdf_vs_oracle_sizes(ops, nb, nf):
    rng = np.random.RandomState(nb * 1000 + nf)
    N = 257
    bins = np.sort(rng.rand(N, nb).astype(np.float32) * 3, -1)
    w = rng.rand(N, nb - 1).astype(np.float32) ** 5
    w[0] = 0
    w[1, : (nb - 1) // 2] = 0
    u = rng.rand(N, nf).astype(np.float32)
    u[2, :3] = [0.0, 1.0, 0.5][: min(3, nf)] + [0.5] * max(0, 3 - nf) if nf >= 3 else u[2, :3]
    for det in (True, False):
        so, io = O.sample_pdf(bins, w, nf, det=det, u=u)
        s, i = ops.sample_pdf(T(bins), T(w), nf, det=det, u=T(u))
        assert np.array_equal(i.cpu().numpy(), io), (nb, nf, det)
        assert np.array_equal(s.cpu().numpy(), so), (nb, nf, det)


def test_sample_pdf_empty(ops):
    s, i = ops.sample_pdf(torch.zeros(0, 63, device=dev()), torch.zeros(0, 62, device=dev()), 16, det=True)
    assert s.shape == (0, 16) and i.shape == (0, 16)


@pytest.mark.parametrize("case", ["coarse_test", "coarse_train", "fine_test", "fine_train"])
def test_raw2outputs(ops, golden, case):
    raw, z = golden["r2o_raw"], golden["r2o_z"]
    sel = dict(coarse_test=raw[..., 3:4], coarse_train=raw[..., :4], fine_test=raw, fine_train=raw)[case]
    o = ops.raw2outputs(T(sel), T(z), "fine" if case.startswith("fine") else "coarse", case.endswith("test"))
    n = 0
    for nm in ["rgb", "disp", "acc", "weights", "depth", "transient_sigmas", "beta"]:
        key = f"r2o_{case}_{nm}"
        if key in golden:
            assert np.allclose(o[nm].cpu().numpy(), golden[key], rtol=2e-5, atol=2e-6), nm
            n += 1
    assert n >= 2


def test_raw2outputs_transmittance_matches_f64_scan(ops):
    """Weights must follow ATen's float64-accumulated cumprod exactly given identical alphas:
    a long ray (S=256) with small alphas exposes float32-scan drift."""
    rng = np.random.RandomState(5)
    S = 256
    raw = np.abs(rng.randn(64, S, 1)).astype(np.float32) * 0.05
    z = np.sort(rng.rand(64, S).astype(np.float32) * 20, -1)
    want = O.raw2outputs_nerfw(raw, z, test_time=True, typ="coarse")["weights"]
    got = ops.raw2outputs(T(raw), T(z), "coarse", True)["weights"].cpu().numpy()
    # alpha itself carries 1-ulp exp differences (relative 1e-5 at alpha ~ 4e-3); a float32 scan
    # would add a drift that grows with the sample index instead
    assert np.allclose(got, want, rtol=2e-5, atol=3e-7)
    big = want > 1e-3
    err = np.where(big, np.abs(got - want) / np.maximum(want, 1e-3), 0.0)
    assert err[:, 192:].sum() / max(big[:, 192:].sum(), 1) < 2 * err[:, :64].sum() / max(big[:, :64].sum(), 1) + 2e-6


@pytest.mark.parametrize("tag,D,W", [("s", 4, 64), ("b", 8, 256)])
def test_nerfw_forward_fp32(ops, golden, tag, D, W):
    (c, f, ea, et), _ = synthetic_nets(D, W)
    h = ops.NerfHandle(*to_dev([c, f, ea, et]))
    x = T(golden[f"mlp_{tag}_x"])
    s = h.nerfw_forward(0, 0, x[:, :63].contiguous())
    assert rel_err(s.cpu().numpy(), golden[f"mlp_{tag}_sigma_only"]) < 5e-5
    st = h.nerfw_forward(0, 1, x[:, :90].contiguous())
    assert rel_err(st.cpu().numpy(), golden[f"mlp_{tag}_coarse_static"]) < 5e-5
    full = h.nerfw_forward(1, 2, x)
    assert rel_err(full.cpu().numpy(), golden[f"mlp_{tag}_fine_full"]) < 5e-5


def test_nerfw_module_forward_uses_cuda_path(ops, golden):
    (c, f, ea, et), _ = synthetic_nets(4, 64)
    c, f = to_dev([c, f])
    x = T(golden["mlp_s_x"])
    before = ops.lib.dfb_launch_count()
    out = f(x, output_transient=True)
    assert ops.lib.dfb_launch_count() > before
    assert rel_err(out.cpu().numpy(), golden["mlp_s_fine_full"]) < 5e-5
    out = c(x[:, :63].contiguous(), sigma_only=True)
    assert rel_err(out.cpu().numpy(), golden["mlp_s_sigma_only"]) < 5e-5


def _render_kwargs(mods, Nc, Nf, test_time, perturb=0.0):
    c, f, ea, et = to_dev(mods)
    return dict(network_query_fn=None, perturb=perturb, N_importance=Nf, network_fine=f, N_samples=Nc, network_fn=c,
                use_viewdirs=True, white_bkgd=False, raw_noise_std=0.0, embedding_a=ea, embedding_t=et,
                test_time=test_time, ndc=False, lindisp=False)


def test_render_cfg1_shape_fp32(golden):
    """BASELINE config[0] shape: coarse-only 4x64 network, train-mode path."""
    from dfnet_b200 import rendering
    mods, _ = synthetic_nets(4, 64, fine=False)
    kw = _render_kwargs(mods, 64, 0, False)
    rgb, disp, acc, ex = rendering.render(8, 8, 8.0, c2w=T(golden["e2e_a_c2w"]), img_idx=T(golden["hist"]),
                                          near=0.0, far=2.5, mma="fp32", **kw)
    assert rgb.shape == (8, 8, 3) and disp.shape == (8, 8)
    assert rel_err(rgb.cpu().numpy(), golden["e2e_a_rgb"]) < 1e-4
    assert rel_err(disp.cpu().numpy(), golden["e2e_a_disp"]) < 1e-4
    assert rel_err(acc.cpu().numpy(), golden["e2e_a_acc"]) < 1e-4


@pytest.mark.parametrize("mma,tol,flip_tol", [("fp32", 1e-4, 0.01), ("f16", 1e-3, 0.05), ("bf16", 1e-2, 0.2)])
def test_render_cfg2_shape(ops, golden, mma, tol, flip_tol):
    """BASELINE config[1] shape (8x256, 64+128, test_time) on a 6x8 image, all precisions."""
    mods, _ = synthetic_nets(8, 256)
    h = ops.NerfHandle(*to_dev(mods))
    o = h.render(64, 128, True, c2w=T(golden["e2e_b_c2w"]), H=6, W=8, focal=7.3125, near=0.0, far=2.5,
                 hist=T(golden["hist"]), mma=mma, want=("inds", "z_samples", "weights_coarse", "z_vals"))
    torch.cuda.synchronize()
    assert rel_err(o["rgb"].cpu().numpy().reshape(6, 8, 3), golden["e2e_b_rgb"]) < tol
    assert rel_err(o["disp"].cpu().numpy().reshape(6, 8), golden["e2e_b_disp"]) < tol
    assert rel_err(o["acc"].cpu().numpy().reshape(6, 8), golden["e2e_b_acc"]) < tol
    inds = o["inds"].cpu().numpy()
    flips = (inds != golden["e2e_b_inds"]).mean()
    assert flips < flip_tol, flips
    # the sampler must be bit-exact on ITS OWN coarse weights (op-level gate for the index claim)
    zc = np.broadcast_to(O.linspace_f32(0, 1, 64)[None] * np.float32(2.5), (48, 64))
    _, io = O.sample_pdf(0.5 * (zc[:, 1:] + zc[:, :-1]), o["weights_coarse"].cpu().numpy()[:, 1:-1], 128, det=True)
    assert np.array_equal(inds, io)
    z = o["z_vals"].cpu().numpy()
    assert (np.diff(z, axis=-1) >= 0).all()


@pytest.mark.parametrize("mma,tol", [("f16", 1e-3), ("bf16", 1e-2)])
def test_render_cfg2_shape_one_cta_kernel(ops, golden, monkeypatch, mma, tol):
    """The 1-CTA variant of the tcgen05 kernel (DFB_TC_CTA_GROUP=1; the default is the cta_group::2 pair kernel):
    same gates as above, and both variants agree to rounding."""
    mods, _ = synthetic_nets(8, 256)
    h = ops.NerfHandle(*to_dev(mods))
    kw = dict(c2w=T(golden["e2e_b_c2w"]), H=6, W=8, focal=7.3125, near=0.0, far=2.5, hist=T(golden["hist"]), mma=mma)
    pair = h.render(64, 128, True, **kw)
    monkeypatch.setenv("DFB_TC_CTA_GROUP", "1")
    one = h.render(64, 128, True, **kw)
    torch.cuda.synchronize()
    for k, g in (("rgb", "e2e_b_rgb"), ("disp", "e2e_b_disp"), ("acc", "e2e_b_acc")):
        assert rel_err(one[k].cpu().numpy().reshape(golden[g].shape), golden[g]) < tol, k
        assert rel_err(one[k].cpu().numpy(), pair[k].cpu().numpy()) < tol, k


@pytest.mark.parametrize("W,Nf", [(128, 64), (64, 128), (192, 64)])
def test_render_narrow_networks_on_tensor_cores(ops, golden, W, Nf):
    """Networks narrower than 256 (the reference's shipped default is netwidth=128, 64+64 samples) run on the tcgen05
    kernels embedded in 8x256 with zero weights — exactly the same function.  Forward against the fp32 kernels (which
    are pinned to the reference for these widths), backward (saved masks and recompute) against the fp32 kernels."""
    from dfnet_b200 import rendering
    mods, _ = synthetic_nets(8, W)
    dmods = to_dev(mods)
    h = ops.NerfHandle(*dmods)
    assert h.tc_train
    kw = dict(c2w=T(golden["e2e_b_c2w"]), H=12, W=16, focal=14.6, near=0.0, far=2.5, hist=T(golden["hist"]))
    ref = h.render(64, Nf, True, mma="fp32", **kw)
    for mma, tol in (("f16", 1e-3), ("bf16", 1e-2)):
        got = h.render(64, Nf, True, mma=mma, **kw)
        torch.cuda.synchronize()
        for k in ("rgb", "disp", "acc"):
            assert rel_err(got[k].cpu().numpy(), ref[k].cpu().numpy()) < tol, (mma, k)
    # differentiable path: rendering.render -> saved-mask tcgen05 backward vs the fp32 kernels
    rays = golden["e2e_c_rays"]
    grads = {}
    for mma in ("f16", "fp32"):
        ro = T(rays[0]).clone().requires_grad_(True)
        rd = T(rays[1]).clone().requires_grad_(True)
        kw2 = _render_kwargs(mods, 64, Nf, True)
        kw2["network_fn"], kw2["network_fine"], kw2["embedding_a"], kw2["embedding_t"] = dmods
        rgb, _, _, _ = rendering.render(4, 6, 5.0, rays=(ro, rd), img_idx=T(golden["hist"]), near=0.0, far=2.5, mma=mma, **kw2)
        torch.manual_seed(0)
        (rgb * torch.randn_like(rgb)).sum().backward()
        grads[mma] = (ro.grad.clone(), rd.grad.clone())
    for a, b in zip(grads["f16"], grads["fp32"]):
        assert torch.isfinite(a).all() and _cos(a, b) > 0.995, _cos(a, b)
    # recompute variant of the backward kernel on the padded network
    rec = T(O.make_ray_records(rays[0], rays[1], 0.0, 2.5, golden["hist"]))
    out = h.render(64, Nf, True, rays=rec, mma="f16", want=("z_vals", "raw", "relu_masks"))
    g = torch.randn(rec.shape[0], 3, device=dev()) * 1e-6
    a = h.render_backward(rec, out["z_vals"], out["raw"], g, mma="f16", relu_masks=out["relu_masks"])
    b = h.render_backward(rec, out["z_vals"], out["raw"], g, mma="f16")
    c = h.render_backward(rec, out["z_vals"], out["raw"], g, mma="fp32")
    for x, y, z in zip(a, b, c):
        assert _cos(x, y) > 0.999 and _cos(x, z) > 0.995, (_cos(x, y), _cos(x, z))


def test_render_train_mode_extras_fp32(golden):
    from dfnet_b200 import rendering
    mods, _ = synthetic_nets(8, 64)
    kw = _render_kwargs(mods, 16, 24, False)
    rays = T(golden["e2e_c_rays"])
    rgb, disp, acc, ex = rendering.render(4, 6, 5.0, rays=(rays[0], rays[1]), img_idx=T(golden["hist"]), near=0.0,
                                          far=2.5, retraw=True, mma="fp32", **kw)
    got = dict(rgb=rgb, disp=disp, acc=acc, **ex)
    for k in ["rgb", "disp", "acc", "rgb0", "disp0", "acc0", "z_std", "transient_sigmas", "beta", "raw"]:
        assert rel_err(got[k].cpu().numpy(), golden[f"e2e_c_{k}"]) < 1e-4, k


def test_render_stratified_with_reference_draws(ops, golden):
    mods, _ = synthetic_nets(8, 64)
    h = ops.NerfHandle(*to_dev(mods))
    rays = golden["e2e_c_rays"]
    rec = O.make_ray_records(rays[0], rays[1], 0.0, 2.5, golden["hist"])
    o = h.render(16, 24, False, rays=T(rec), perturb=True, t_rand=T(golden["e2e_d_t_rand"]), u=T(golden["e2e_d_u"]),
                 mma="fp32", want=("rgb0", "beta", "z_std", "inds"))
    for k, g in [("rgb", "rgb"), ("disp", "disp"), ("acc", "acc"), ("rgb0", "rgb0"), ("beta", "beta"),
                 ("z_std", "z_std")]:
        assert rel_err(o[k].cpu().numpy(), golden[f"e2e_d_{g}"]) < 1e-4, k
    assert (o["inds"].cpu().numpy() != golden["e2e_d_inds"]).mean() < 0.01


def test_render_error_behaviour(ops):
    mods, _ = synthetic_nets(4, 64, fine=False)
    h = ops.NerfHandle(*to_dev(mods))
    from dfnet_b200._lib import DfbError
    with pytest.raises(DfbError):  # reference returns rgb_map=None here and crashes
        h.render(64, 0, True, c2w=torch.eye(4, device=dev())[:3], H=2, W=2, focal=1.0, hist=torch.zeros(10), mma="fp32")
    with pytest.raises(DfbError):  # no fine network loaded
        h.render(64, 16, True, c2w=torch.eye(4, device=dev())[:3], H=2, W=2, focal=1.0, hist=torch.zeros(10), mma="fp32")


@pytest.mark.parametrize("kind,N,K", [(1, 256, 64), (1, 256, 256), (1, 128, 128), (2, 128, 320), (1, 32, 16)])
def test_umma_descriptor_selftest(ops, kind, N, K):
    """One-tile tcgen05 GEMM through the kernel's shared-memory descriptors / TMEM loads."""
    import ctypes as C
    rng = np.random.RandomState(N + K)
    A = rng.randn(128, K).astype(np.float32)
    B = rng.randn(N, K).astype(np.float32)
    dt = torch.float16 if kind == 1 else torch.bfloat16
    Ar = torch.tensor(A).to(dt).float().numpy()
    Br = torch.tensor(B).to(dt).float().numpy()
    want = Ar.astype(np.float64) @ Br.astype(np.float64).T
    errs = {}
    for variant in (0,):  # variant 1 (LBO/SBO swapped) reads outside the CTA's shared memory: kept for bring-up only
        D = torch.zeros(128, N, device=dev())
        At, Bt = T(A), T(B)  # keep the device copies alive across the call
        ops.check(ops.lib.dfb_debug_umma_gemm(C.c_void_p(At.data_ptr()), C.c_void_p(Bt.data_ptr()), N, K, kind,
                                              variant, C.c_void_p(D.data_ptr()), None))
        torch.cuda.synchronize()
        errs[variant] = float(np.abs(D.cpu().numpy() - want).max())
    print("umma selftest max abs err by descriptor variant:", errs)
    assert errs[0] < 1e-3 * np.sqrt(K), errs


def test_render_cfg5_shape_and_default_width(ops):
    """BASELINE config[4] sampling (64+192 -> S=256, far=20) on the tensor-core path and the reference's
    default network width 128 (models/options.py:30-33, fp32 path) against the oracle."""
    hist = np.array([[5, 10, 20, 30, 15, 10, 5, 3, 1, 1]], np.float32)
    c2w = np.array([[0.9848, 0.0, 0.1736, 0.2], [0.0, 1.0, 0.0, -0.1], [-0.1736, 0.0, 0.9848, 1.0]], np.float32)
    mods, nets = synthetic_nets(8, 256)
    h = ops.NerfHandle(*to_dev(mods))
    want = O.render(5, 7, 9.0, nets, 64, 192, 0.0, 20.0, c2w=c2w, hist=hist, test_time=True)
    got = h.render(64, 192, True, c2w=T(c2w), H=5, W=7, focal=9.0, near=0.0, far=20.0, hist=T(hist), mma="f16")
    torch.cuda.synchronize()
    assert rel_err(got["rgb"].cpu().numpy().reshape(5, 7, 3), want["rgb_map"]) < 1e-3
    assert rel_err(got["acc"].cpu().numpy().reshape(5, 7), want["acc_map"]) < 1e-3
    mods, nets = synthetic_nets(8, 128)
    h = ops.NerfHandle(*to_dev(mods))
    want = O.render(5, 7, 9.0, nets, 64, 64, 0.0, 2.5, c2w=c2w, hist=hist, test_time=True)
    got = h.render(64, 64, True, c2w=T(c2w), H=5, W=7, focal=9.0, near=0.0, far=2.5, hist=T(hist), mma="f16")  # falls to fp32: W != 256
    torch.cuda.synchronize()
    assert rel_err(got["rgb"].cpu().numpy().reshape(5, 7, 3), want["rgb_map"]) < 1e-4
    assert rel_err(got["disp"].cpu().numpy().reshape(5, 7), want["disp_map"]) < 1e-4


def _torch_render_rgb(mods, rays_o, rays_d, z_vals, hist):
    """Differentiable float64 restatement of the test-time fine render for FIXED depths z_vals (the
    reference detaches z_samples, models/rendering.py:302): nerfw.py:62-95,297-354 and rendering.py:169-212."""
    import torch.nn.functional as F
    c, f, ea, et = mods
    P = {k: v.double() for k, v in f.state_dict().items()}

    def embed(x, L):
        out = [x]
        for l in range(L):
            out += [torch.sin(x * 2.0 ** l), torch.cos(x * 2.0 ** l)]
        return torch.cat(out, -1)
    N, S = z_vals.shape
    vd = rays_d / rays_d.norm(dim=-1, keepdim=True)
    pts = rays_o[:, None] + rays_d[:, None] * z_vals[..., None]
    idx = hist.long()
    a = ea.weight.double()[idx].reshape(1, -1).expand(N * S, -1)
    tt = et.weight.double()[idx].reshape(1, -1).expand(N * S, -1)
    xyz = embed(pts.reshape(-1, 3), 10)
    dirs = embed(vd, 4)[:, None].expand(N, S, 27).reshape(-1, 27)
    h = xyz
    for i in range(f.D):
        if i in f.skips:
            h = torch.cat([xyz, h], 1)
        h = F.relu(F.linear(h, P[f"xyz_encoding_{i+1}.0.weight"], P[f"xyz_encoding_{i+1}.0.bias"]))
    sig = F.softplus(F.linear(h, P["static_sigma.0.weight"], P["static_sigma.0.bias"]))[:, 0]
    fin = F.linear(h, P["xyz_encoding_final.weight"], P["xyz_encoding_final.bias"])
    de = F.relu(F.linear(torch.cat([fin, dirs, a], 1), P["dir_encoding.0.weight"], P["dir_encoding.0.bias"]))
    rgb = torch.sigmoid(F.linear(de, P["static_rgb.0.weight"], P["static_rgb.0.bias"]))
    t = torch.cat([fin, tt], 1)
    for k in (0, 2, 4, 6):
        t = F.relu(F.linear(t, P[f"transient_encoding.{k}.weight"], P[f"transient_encoding.{k}.bias"]))
    tsig = F.softplus(F.linear(t, P["transient_sigma.0.weight"], P["transient_sigma.0.bias"]))[:, 0]
    trgb = torch.sigmoid(F.linear(t, P["transient_rgb.0.weight"], P["transient_rgb.0.bias"]))
    sig, tsig, rgb, trgb = sig.reshape(N, S), tsig.reshape(N, S), rgb.reshape(N, S, 3), trgb.reshape(N, S, 3)
    deltas = torch.cat([z_vals[:, 1:] - z_vals[:, :-1], torch.full_like(z_vals[:, :1], 1e2)], -1)
    a_s, a_t = 1 - torch.exp(-deltas * sig), 1 - torch.exp(-deltas * tsig)
    al = 1 - torch.exp(-deltas * (sig + tsig))
    T = torch.cumprod(torch.cat([torch.ones_like(al[:, :1]), 1 - al], -1)[:, :-1], -1)
    return ((a_s * T)[..., None] * rgb).sum(1) + ((a_t * T)[..., None] * trgb).sum(1)


@pytest.mark.parametrize("D,W,Nc,Nf", [(8, 64, 16, 24), (8, 256, 64, 128)])
def test_render_backward_wrt_rays_matches_autograd(ops, golden, D, W, Nc, Nf):
    """dfb_render_bwd against float64 autograd of a torch restatement (fixed depths, as the reference detaches them)."""
    from dfnet_b200 import rendering
    mods, _ = synthetic_nets(D, W)
    dmods = to_dev(mods)
    rays = golden["e2e_c_rays"]
    hist = T(golden["hist"])
    ro = T(rays[0]).clone().requires_grad_(True)
    rd = T(rays[1]).clone().requires_grad_(True)
    kw = _render_kwargs(mods, Nc, Nf, True)
    kw["network_fn"], kw["network_fine"], kw["embedding_a"], kw["embedding_t"] = dmods
    rgb, disp, acc, _ = rendering.render(4, 6, 5.0, rays=(ro, rd), img_idx=hist, near=0.0, far=2.5, mma="fp32", **kw)
    torch.manual_seed(0)
    wgt = torch.randn_like(rgb)
    (rgb * wgt).sum().backward()
    g_o, g_d = ro.grad.clone(), rd.grad.clone()
    # reference gradient: same depths, float64 autograd
    h = ops.handle_for(*dmods)
    rec = O.make_ray_records(rays[0], rays[1], 0.0, 2.5, golden["hist"])
    z = h.render(Nc, Nf, True, rays=T(rec), mma="fp32", want=("z_vals",))["z_vals"].double()
    ro64 = T(rays[0]).double().requires_grad_(True)
    rd64 = T(rays[1]).double().requires_grad_(True)
    rgb64 = _torch_render_rgb(dmods, ro64, rd64, z, hist.reshape(-1))
    assert rel_err(rgb.detach().cpu().numpy(), rgb64.detach().cpu().numpy()) < 1e-4
    (rgb64 * wgt.double()).sum().backward()
    for got, want in ((g_o, ro64.grad), (g_d, rd64.grad)):
        scale = float(want.abs().max())
        assert float((got.double() - want).abs().max()) < 2e-3 * scale, (float((got.double() - want).abs().max()), scale)


def _cos(a, b):
    a, b = a.double().reshape(-1), b.double().reshape(-1)
    return float((a * b).sum() / (a.norm() * b.norm() + 1e-300))


def _ballot_to_tc_layout(m):
    """[P,12,8] ReLU mask words: fp32 kernels' ballot layout (word j bit l = column 32j+l) -> the tcgen05 kernel's
    layout (bit 16*(c&1) + (c>>1) of the same word = column 32j+c): even columns to the low half, odd to the high."""
    m = m.to(torch.int64) & 0xFFFFFFFF
    out = torch.zeros_like(m)
    for c in range(32):
        out |= ((m >> c) & 1) << (16 * (c & 1) + (c >> 1))
    return (out - ((out >> 31) & 1) * (1 << 32)).to(torch.int32)


@pytest.mark.parametrize("mma,tol,flip", [("f16", 3e-3, 5e-3), ("bf16", 2e-2, 3e-2)])
def test_render_backward_tcgen05_vs_fp32_kernels(ops, golden, mma, tol, flip):
    """dfb_render_bwd_mma (tcgen05 forward recompute + input-gradient chain, 8x256 network) against the fp32 kernels
    on the SAME saved forward state (z_vals, raw) and a loss-scale upstream gradient (|g| ~ 1e-7: exercises the
    per-row power-of-two scaling of the 16-bit gradient operands).  2000 rays x 192 samples = 10 passes per CTA.

    The gradient is discontinuous in the forward values (ReLU masks): a 16-bit forward flips ~0.1 % of the masks of
    an fp32 forward and every flip moves a sample's gradient by ~1/sqrt(active units) (measured un-pinned: 4 %
    relative L2, cosine 0.9993).  The arithmetic of the chain is therefore compared with the masks PINNED to the fp32
    kernels' (dfb_debug_bwd_masks); the recompute itself is gated by its flip rate, and the un-pinned result by its
    cosine."""
    import ctypes as C
    from dfnet_b200._lib import lib, check
    mods, _ = synthetic_nets(8, 256)
    h = ops.handle_for(*to_dev(mods))
    Hh, Ww = 40, 50
    rng = np.random.RandomState(7)
    c2w = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 1.0]], np.float32)
    o, d = ops.get_rays(Hh, Ww, 45.0, T(c2w))
    rec = T(O.make_ray_records(o.reshape(-1, 3).cpu().numpy(), d.reshape(-1, 3).cpu().numpy(), 0.0, 2.5, golden["hist"]))
    out = h.render(64, 128, True, rays=rec, mma=mma, want=("z_vals", "raw"))
    g_rgb = T((rng.randn(Hh * Ww, 3) * 1e-7).astype(np.float32))
    g_rgb[5] = 0.0                                            # a ray without gradient
    P = Hh * Ww * 192
    m_simt = torch.zeros(P, 12, 8, dtype=torch.int32, device=dev())
    m_tc = torch.zeros(P, 12, 8, dtype=torch.int32, device=dev())
    vp = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    try:
        check(lib.dfb_debug_bwd_masks(vp(m_simt), None, None))
        want = h.render_backward(rec, out["z_vals"], out["raw"], g_rgb, mma="fp32")
        torch.cuda.synchronize()
        m_in = _ballot_to_tc_layout(m_simt).contiguous()
        check(lib.dfb_debug_bwd_masks(None, vp(m_in), vp(m_tc)))
        got = h.render_backward(rec, out["z_vals"], out["raw"], g_rgb, mma=mma)
        torch.cuda.synchronize()
    finally:
        check(lib.dfb_debug_bwd_masks(None, None, None))
    # (1) the chain on identical ReLU patterns
    for nm, a, b in zip(("g_o", "g_d", "g_vd"), got, want):
        assert torch.isfinite(a).all(), nm
        scale = float(b.abs().max())
        err = float((a - b).abs().max())
        assert err < tol * scale, (nm, err, scale)
        assert float(a[5].abs().max()) == 0.0
    # (2) the forward recompute: fraction of ReLU decisions that differ from the fp32 recompute
    valid = torch.ones(12, 8, dtype=torch.bool, device=dev())
    valid[9:, 4:] = False                                     # 128-wide layers use 4 words
    x = ((m_tc ^ m_in).to(torch.int64) & 0xFFFFFFFF)[:, valid]
    flips = sum(int(((x >> b) & 1).sum()) for b in range(32))
    rate = flips / (P * (9 * 256 + 3 * 128))
    assert rate < flip, rate
    # (3) un-pinned (the product path): direction of the gradient
    free = h.render_backward(rec, out["z_vals"], out["raw"], g_rgb, mma=mma)
    for nm, a, b in zip(("g_o", "g_d", "g_vd"), free, want):
        assert _cos(a, b) > (0.998 if mma == "f16" else 0.99), (nm, _cos(a, b))
    # ragged tail: a ray count that leaves the last tile / pass partially filled
    n = 777
    want = h.render_backward(rec[:n], out["z_vals"][:n], out["raw"][:n], g_rgb[:n], mma="fp32")
    got = h.render_backward(rec[:n], out["z_vals"][:n], out["raw"][:n], g_rgb[:n], mma=mma)
    for a, b, c in zip(got, want, free):
        assert _cos(a, b) > (0.998 if mma == "f16" else 0.99)
        assert torch.equal(a, c[:n])                          # samples are independent of their tile position


@pytest.mark.parametrize("mma", ["f16", "bf16"])
def test_render_backward_saved_masks(ops, golden, mma):
    """Training forward that saves its ReLU masks (DfbRenderExtras::relu_masks) + dfb_render_bwd_saved (no forward
    recompute): bit-identical to the recompute kernel when that one is pinned to the same masks, same direction as the
    fp32 kernels; the saved masks differ from the recompute's own masks only by rounding-level flips."""
    import ctypes as C
    from dfnet_b200._lib import lib, check
    mods, _ = synthetic_nets(8, 256)
    h = ops.handle_for(*to_dev(mods))
    Hh, Ww = 40, 50
    c2w = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 1.0]], np.float32)
    o, d = ops.get_rays(Hh, Ww, 45.0, T(c2w))
    rec = T(O.make_ray_records(o.reshape(-1, 3).cpu().numpy(), d.reshape(-1, 3).cpu().numpy(), 0.0, 2.5, golden["hist"]))
    plain = h.render(64, 128, True, rays=rec, mma=mma, want=("z_vals", "raw"))
    out = h.render(64, 128, True, rays=rec, mma=mma, want=("z_vals", "raw", "relu_masks"))
    assert torch.equal(plain["raw"], out["raw"]) and torch.equal(plain["rgb"], out["rgb"])   # recording changes nothing
    P = Hh * Ww * 192
    g_rgb = T((np.random.RandomState(3).randn(Hh * Ww, 3) * 1e-7).astype(np.float32))
    got = h.render_backward(rec, out["z_vals"], out["raw"], g_rgb, mma=mma, relu_masks=out["relu_masks"])
    torch.cuda.synchronize()
    m_in = out["relu_masks"].permute(0, 3, 1, 2).reshape(-1, 12, 8)[:P].contiguous()   # [tile,12,8,128] -> [P,12,8]
    m_own = torch.zeros(P, 12, 8, dtype=torch.int32, device=dev())
    vp = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    try:
        check(lib.dfb_debug_bwd_masks(None, vp(m_in), vp(m_own)))
        pinned = h.render_backward(rec, out["z_vals"], out["raw"], g_rgb, mma=mma)
        torch.cuda.synchronize()
    finally:
        check(lib.dfb_debug_bwd_masks(None, None, None))
    for a, b in zip(got, pinned):
        assert torch.equal(a, b)
    valid = torch.ones(12, 8, dtype=torch.bool, device=dev())
    valid[9:, 4:] = False
    x = ((m_own ^ m_in).to(torch.int64) & 0xFFFFFFFF)[:, valid]
    rate = sum(int(((x >> b) & 1).sum()) for b in range(32)) / (P * (9 * 256 + 3 * 128))
    assert rate < (2e-3 if mma == "f16" else 2e-2), rate
    want = h.render_backward(rec, out["z_vals"], out["raw"], g_rgb, mma="fp32")
    for a, b in zip(got, want):
        assert _cos(a, b) > (0.998 if mma == "f16" else 0.99), _cos(a, b)
    n = 777   # ragged tail
    sub = h.render(64, 128, True, rays=rec[:n], mma=mma, want=("z_vals", "raw", "relu_masks"))
    g2 = h.render_backward(rec[:n], sub["z_vals"], sub["raw"], g_rgb[:n], mma=mma, relu_masks=sub["relu_masks"])
    for a, b in zip(g2, got):
        assert torch.equal(a, b[:n])


def test_render_backward_tcgen05_matches_autograd(ops, golden):
    """End to end through rendering.render with mma="f16": forward AND backward on tcgen05, against float64 autograd."""
    from dfnet_b200 import rendering
    mods, _ = synthetic_nets(8, 256)
    dmods = to_dev(mods)
    rays = golden["e2e_c_rays"]
    hist = T(golden["hist"])
    ro = T(rays[0]).clone().requires_grad_(True)
    rd = T(rays[1]).clone().requires_grad_(True)
    kw = _render_kwargs(mods, 64, 128, True)
    kw["network_fn"], kw["network_fine"], kw["embedding_a"], kw["embedding_t"] = dmods
    rgb, _, _, _ = rendering.render(4, 6, 5.0, rays=(ro, rd), img_idx=hist, near=0.0, far=2.5, mma="f16", **kw)
    torch.manual_seed(0)
    wgt = torch.randn_like(rgb)
    (rgb * wgt).sum().backward()
    h = ops.handle_for(*dmods)
    rec = O.make_ray_records(rays[0], rays[1], 0.0, 2.5, golden["hist"])
    z = h.render(64, 128, True, rays=T(rec), mma="f16", want=("z_vals",))["z_vals"].double()
    ro64 = T(rays[0]).double().requires_grad_(True)
    rd64 = T(rays[1]).double().requires_grad_(True)
    rgb64 = _torch_render_rgb(dmods, ro64, rd64, z, hist.reshape(-1))
    (rgb64 * wgt.double()).sum().backward()
    # un-pinned: the fp16 forward flips ~0.1 % of the ReLU decisions of the float64 forward (see the test above)
    for got, want in ((ro.grad, ro64.grad), (rd.grad, rd64.grad)):
        scale = float(want.abs().max())
        assert float((got.double() - want).abs().max()) < 6e-2 * scale, (float((got.double() - want).abs().max()), scale)
        assert _cos(got, want) > 0.998, _cos(got, want)
