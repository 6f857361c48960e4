"""NeRF-Hist training step: `render` in train mode differentiable w.r.t. the NeRF-W networks and the
histogram embeddings, against the UNMODIFIED reference's step (tests/golden/make_golden_nerf_train.py: render(...,
retraw=True, **render_kwargs_train) -> NerfWLoss -> backward, CPU fp32), plus kernel-level checks of the pieces."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from helpers import nerf_train_case

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda:0")


def _kwargs(mods, Nc, Nf):
    c, f, ea, et = mods
    return dict(network_query_fn=None, perturb=0.0, N_importance=Nf, network_fine=f, N_samples=Nc, network_fn=c, use_viewdirs=True,
                white_bkgd=False, raw_noise_std=0.0, embedding_a=ea, embedding_t=et, test_time=False, ndc=False, lindisp=False)


@pytest.mark.parametrize("case", ["w128", "w256"])
def test_nerf_training_step_vs_reference_golden(case):
    from dfnet_b200 import rendering
    from dfnet_b200.losses import loss_dict
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "nerf_train_golden.npz"))
    cfg = nerf_train_case(case)
    mods = [m.to(dev()) for m in cfg["mods"]]
    for m in mods:
        for p in m.parameters():
            p.requires_grad_(True)
    kw = _kwargs(mods, cfg["Nc"], cfg["Nf"])
    rays = torch.tensor(cfg["rays"], device=dev())
    target = torch.tensor(cfg["target"], device=dev())
    rgb, disp, acc, extras = rendering.render(1, 1, 1.0, chunk=32768, rays=(rays[0], rays[1]), retraw=True, near=cfg["near"],
                                              far=cfg["far"], img_idx=torch.tensor(cfg["hist"], device=dev()), **kw)
    assert extras["raw"].shape == (rays.shape[1], cfg["Nc"] + cfg["Nf"], 9)
    results = {"rgb_fine": rgb, "rgb_coarse": extras["rgb0"], "beta": extras["beta"], "transient_sigmas": extras["transient_sigmas"]}
    loss_d = loss_dict["nerfw"](coef=1)(results, target)
    loss = sum(l for l in loss_d.values())
    loss.backward()
    # forward: fp16 operands against the fp32 reference
    for k, want in (("rgb", rgb), ("rgb0", extras["rgb0"]), ("beta", extras["beta"])):
        e = float(np.abs(want.detach().cpu().numpy() - g[f"{case}_{k}"]).max() / np.abs(g[f"{case}_{k}"]).max())
        print(case, k, "max err / max", e)
        assert e < 2e-3, (k, e)
    for k in ("c_l", "f_l", "b_l", "s_l"):
        assert abs(float(loss_d[k]) - float(g[f"{case}_{k}"])) < 2e-3 * max(abs(float(g[f"{case}_{k}"])), 1e-2), k
    assert abs(float(loss) - float(g[f"{case}_loss"])) < 1e-3 * abs(float(g[f"{case}_loss"]))
    names = bytes(g[f"{case}_names"]).decode().split("\n")
    got = {}
    for tag, m in zip(("coarse", "fine", "emb_a", "emb_t"), mods):
        for n, p in m.named_parameters():
            got[f"{tag}.{n}"] = p.grad if p.grad is not None else torch.zeros_like(p)
    assert sorted(got) == sorted(names)
    worst_cos, worst_norm = (1.0, ""), (0.0, "")
    for n in names:
        gg = got[n].flatten()
        want_norm = float(g[f"{case}_g_{n}_stats"][0])
        sub = gg[:: max(1, gg.numel() // 2048)][:2048].double().cpu()
        want = torch.from_numpy(g[f"{case}_g_{n}_sub"]).double()
        if want_norm < 1e-12:
            assert float(gg.norm()) < 1e-6, n
            continue
        cos = float(F.cosine_similarity(sub, want, dim=0)) if float(want.norm()) > 0 else 1.0
        nr = abs(float(gg.norm()) / want_norm - 1.0)
        worst_cos = min(worst_cos, (cos, n))
        worst_norm = max(worst_norm, (nr, n))
    print(case, "loss", float(loss), float(g[f"{case}_loss"]), "worst cos", worst_cos, "worst norm dev", worst_norm)
    # bf16 gradients through fp16 activations (ReLU masks of a few units flip against the fp32 run)
    assert worst_cos[0] > 0.995 and worst_norm[0] < 0.03, (worst_cos, worst_norm)    # measured 0.9990 / 0.8 %


@pytest.mark.parametrize("S,seq", [(40, "0"), (40, "1"), (33, "0"), (128, "0"), (192, "0"), (256, "0"), (300, "0")])
def test_raw2outputs_backward_vs_float64_autograd(S, seq, monkeypatch):
    """dfb_raw2outputs_bwd (fine with rgb / beta / transient_sigmas upstream, coarse with noise) against float64 autograd
    of a torch restatement of raw2outputs_NeRFW (rendering.py:132-243, train mode): the warp-per-ray kernel (prefix / suffix
    scans; sample counts that are not a multiple of 32, the 256-sample limit) and the one-thread-per-ray kernel
    (DFB_R2O_BWD_SEQ=1, and S > 256)."""
    from dfnet_b200.nerf_train import _CompositeFn
    monkeypatch.setenv("DFB_R2O_BWD_SEQ", seq)
    torch.manual_seed(4)
    N = 37
    z, _ = torch.sort(torch.rand(N, S, device=dev()) * 2.5, -1)
    for typ, Cc in (("fine", 9), ("coarse", 4)):
        raw = torch.rand(N, S, Cc, device=dev())
        raw[..., 3] = raw[..., 3] * 6 - (1.0 if typ == "coarse" else 0.0)
        if Cc == 9:
            raw[..., 7] *= 3
        noise = torch.randn(N, S, device=dev()) if typ == "coarse" else None
        std = 0.7 if typ == "coarse" else 0.0
        x = raw.clone().requires_grad_(True)
        rgb, disp, acc, w, beta, tsig = _CompositeFn.apply(x, z, typ, 0.1, noise, std)
        gr, gb, gt = torch.randn_like(rgb), torch.randn_like(beta), torch.randn(N, S, device=dev())
        L = (rgb * gr).sum() + ((beta * gb).sum() + (tsig * gt).sum() if typ == "fine" else 0.0)
        L.backward()
        xd = raw.double().requires_grad_(True)
        zd = z.double()
        delta = torch.cat([zd[:, 1:] - zd[:, :-1], 1e2 * torch.ones_like(zd[:, :1])], -1)
        if typ == "fine":
            a_s, a_t = 1 - torch.exp(-delta * xd[..., 3]), 1 - torch.exp(-delta * xd[..., 7])
            a = 1 - torch.exp(-delta * (xd[..., 3] + xd[..., 7]))
        else:
            a = 1 - torch.exp(-delta * torch.relu(xd[..., 3] + noise.double() * std))
        T = torch.cumprod(torch.cat([torch.ones_like(a[:, :1]), 1 - a], -1)[:, :-1], -1)
        if typ == "fine":
            rgb_r = ((a_s * T)[..., None] * xd[..., :3]).sum(1) + ((a_t * T)[..., None] * xd[..., 4:7]).sum(1)
            beta_r = (a_t * T * xd[..., 8]).sum(1) + 0.1
            Lr = (rgb_r * gr.double()).sum() + (beta_r * gb.double()).sum() + (xd[..., 7] * gt.double()).sum()
        else:
            rgb_r = ((a * T)[..., None] * xd[..., :3]).sum(1)
            Lr = (rgb_r * gr.double()).sum()
        Lr.backward()
        assert float((rgb.double() - rgb_r).abs().max()) < 1e-5
        err = float((x.grad.double() - xd.grad).abs().max() / xd.grad.abs().max())
        print(typ, "raw2outputs backward max err / max", err)
        assert err < 2e-5


def test_nerf_training_reduces_the_loss():
    """A few Adam steps of the reference's loop body on one ray bundle: the loss goes down, with perturb and noise on."""
    from dfnet_b200 import rendering
    from dfnet_b200.losses import loss_dict
    cfg = nerf_train_case("w128")
    mods = [m.to(dev()) for m in cfg["mods"]]
    params = [p for m in mods for p in m.parameters()]
    for p in params:
        p.requires_grad_(True)
    opt = torch.optim.Adam(params, lr=5e-4)
    kw = _kwargs(mods, cfg["Nc"], cfg["Nf"])
    kw["perturb"], kw["raw_noise_std"] = 1.0, 1.0
    rays = torch.tensor(cfg["rays"], device=dev())
    target = torch.tensor(cfg["target"], device=dev()) * 0.2 + 0.4
    torch.manual_seed(0)
    hist = []
    for it in range(40):
        rgb, _, _, ex = rendering.render(1, 1, 1.0, chunk=32768, rays=(rays[0], rays[1]), retraw=True, near=cfg["near"], far=cfg["far"],
                                         img_idx=torch.tensor(cfg["hist"], device=dev()), **kw)
        opt.zero_grad()
        ld = loss_dict["nerfw"](coef=1)({"rgb_fine": rgb, "rgb_coarse": ex["rgb0"], "beta": ex["beta"],
                                         "transient_sigmas": ex["transient_sigmas"]}, target)
        loss = sum(ld.values())
        loss.backward()
        opt.step()
        hist.append((float(loss.detach()), float(ld["c_l"].detach())))
    print("total loss", hist[0][0], "->", hist[-1][0], " coarse colour loss", hist[0][1], "->", hist[-1][1])
    assert np.isfinite(np.array(hist)).all()
    assert hist[-1][0] < hist[0][0] - 0.2 and hist[-1][1] < hist[0][1]    # NeRF-W's loss trades beta against the fine residual;
                                                                          # the total and the (unweighted) coarse term must fall
    # the inference path sees the updated weights (handles re-upload on version change)
    with torch.no_grad():
        kt = dict(kw, perturb=False, raw_noise_std=0.0, test_time=True)
        out = rendering.render(1, 1, 1.0, rays=(rays[0], rays[1]), near=cfg["near"], far=cfg["far"],
                               img_idx=torch.tensor(cfg["hist"], device=dev()), **kt)[0]
    assert torch.isfinite(out).all()


def test_train_on_batch_nerfw_loop_body():
    """The reference's loop body as one call (run_nerf.py:33-77): random pixels of an image, loss, Adam step, lr decay."""
    import types
    from dfnet_b200 import nerf_train
    from dfnet_b200.losses import loss_dict
    cfg = nerf_train_case("w128")
    mods = [m.to(dev()) for m in cfg["mods"]]
    params = [p for m in mods for p in m.parameters()]
    for p in params:
        p.requires_grad_(True)
    args = types.SimpleNamespace(chunk=32768, lrate=5e-4, lrate_decay=5)
    opt = torch.optim.Adam(params, lr=args.lrate)
    kw = _kwargs(mods, 32, 32)
    kw["perturb"] = 1.0
    H, W, focal = 60, 80, 73.0
    torch.manual_seed(1)
    np.random.seed(1)
    target = torch.rand(3, H, W) * 0.3 + 0.3
    pose = torch.tensor([[1., 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 1]])
    hist = torch.tensor(cfg["hist"])
    losses = []
    for step in range(12):
        loss, psnr = nerf_train.train_on_batch_nerfw(args, target, pose, hist, H, W, focal, 512, opt, loss_dict["nerfw"](coef=1), step, kw,
                                                     near=0.0, far=2.5)
        losses.append(float(loss))
    assert np.isfinite(losses).all() and losses[-1] < losses[0] and torch.isfinite(psnr)
    assert abs(opt.param_groups[0]["lr"] - args.lrate * 0.1 ** (11 / 5000)) < 1e-12


def test_copy2d_batch_vs_torch_copies():
    """dfb_copy2d_batch (parameters -> padded staging, one launch): column slices, row offsets, 1-row biases, > 96 items
    (two launches), an empty item; bit-exact against tensor.copy_ and nothing outside the rectangles is touched."""
    from dfnet_b200._lib import Copy2d, check, lib
    gen = torch.Generator().manual_seed(3)
    items, pairs = [], []
    for k in range(130):
        R, Cs, Cd = int(torch.randint(1, 70, (1,), generator=gen)), int(torch.randint(8, 300, (1,), generator=gen)), 0
        cols = int(torch.randint(0 if k == 7 else 1, Cs + 1, (1,), generator=gen))
        c0 = int(torch.randint(0, Cs - cols + 1, (1,), generator=gen))
        Cd = cols + int(torch.randint(0, 9, (1,), generator=gen))
        src = torch.randn(R, Cs, generator=gen).to(dev())
        dst = torch.full((R + 3, Cd), -7.0, device=dev())
        want = dst.clone()
        want[2:2 + R, :cols] = src[:, c0:c0 + cols]
        view = src[:, c0:c0 + cols]
        items.append((view.data_ptr() if cols else src.data_ptr(), dst.data_ptr() + 4 * 2 * Cd, R, cols, Cs, Cd))
        pairs.append((dst, want, src))
    tbl = (Copy2d * len(items))(*[Copy2d(*it) for it in items])
    check(lib.dfb_copy2d_batch(tbl, len(items), None))
    torch.cuda.synchronize()
    for dst, want, _ in pairs:
        assert torch.equal(dst, want)


def test_trainer_refresh_tracks_parameter_updates():
    """The executor's staged copy of the parameters follows in-place optimizer updates (version counters) and the cached
    copy table survives them: forward after `p.add_` equals a fresh executor's forward on the same values."""
    from dfnet_b200 import nerf_train, nerfw
    mods = [m.to(dev()) for m in nerfw.make_synthetic_nerf(D=8, W=128, fine=True)]
    fine = mods[1]
    tr = nerf_train.NetTrainer(fine)
    P = 64
    pe = (torch.randn(P, 64, device=dev()) * 0.5).half()
    rb_d, rb_t = torch.randn(8, 64, device=dev()) * 0.1, torch.randn(8, 64, device=dev()) * 0.1
    with torch.no_grad():
        raw0, _ = tr.forward(pe, P, rb_d, rb_t, 8)
        raw0 = raw0.clone()
        tr._live = False
        for p in fine.parameters():
            p.add_(0.01 * torch.randn_like(p))
        raw1, _ = tr.forward(pe, P, rb_d, rb_t, 8)
        raw2, _ = nerf_train.NetTrainer(fine).forward(pe, P, rb_d, rb_t, 8)
    assert (raw1 - raw0).abs().max() > 1e-4
    assert torch.equal(raw1, raw2)



def test_embed_xyz16_with_bf16_twin_vs_torch():
    """dfb_embed_xyz16_ex: [pts, sin / cos of 10 bands, zero column] as fp16 rows (models/nerfw.py:105-133 on
    pts = o + d z) and the same values as bf16; padding rows past N S stay untouched."""
    from dfnet_b200._lib import check, lib
    torch.manual_seed(0)
    N, S = 11, 13
    rays = torch.randn(N, 21, device=dev())
    z, _ = torch.sort(torch.rand(N, S, device=dev()) * 2.5, -1)
    P = N * S
    Pp = (P + 7) // 8 * 8
    out = torch.full((Pp, 64), 9.0, device=dev(), dtype=torch.float16)
    out_b = torch.full((Pp, 64), 9.0, device=dev(), dtype=torch.bfloat16)
    out_1 = torch.zeros(Pp, 64, device=dev(), dtype=torch.float16)
    check(lib.dfb_embed_xyz16_ex(rays.data_ptr(), 21, z.data_ptr(), N, S, 10, 64, out.data_ptr(), out_b.data_ptr(), None))
    check(lib.dfb_embed_xyz16(rays.data_ptr(), 21, z.data_ptr(), N, S, 10, 64, out_1.data_ptr(), None))
    pts = (rays[:, None, :3] + rays[:, None, 3:6] * z[..., None]).reshape(-1, 3)
    want = [pts]
    for l in range(10):
        want += [torch.sin(pts * 2.0 ** l), torch.cos(pts * 2.0 ** l)]
    want = torch.cat(want + [torch.zeros(P, 1, device=dev())], -1)
    assert float((out[:P].float() - want).abs().max()) < 2e-3          # fp16 rounding of values in [-4, 4]
    assert torch.equal(out[:P], out_1[:P])
    assert torch.equal(out_b[:P], out[:P].float().bfloat16())
    assert bool((out[P:] == 9.0).all()) and bool((out_b[P:] == 9.0).all())


def test_heads_backward_vs_torch():
    """dfb_nerf_heads_bwd: d raw -> d pre-activation of the Sigmoid / Softplus heads (models/nerfw.py:275-295) as bf16
    [P, 64] rows, zero past the head's width."""
    from dfnet_b200._lib import check, lib
    torch.manual_seed(1)
    for Cc in (9, 4):
        P = 1003
        pre = torch.randn(P, Cc, device=dev()) * 2
        sig = [3] + ([7, 8] if Cc == 9 else [])
        x = pre.clone().requires_grad_(True)
        raw = torch.sigmoid(x)
        raw = torch.cat([F.softplus(x[:, c:c + 1]) if c in sig else raw[:, c:c + 1] for c in range(Cc)], -1)
        g = torch.randn(P, Cc, device=dev())
        raw.backward(g)
        gs, gr, gt = (torch.full((P, 64), 5.0, device=dev(), dtype=torch.bfloat16) for _ in range(3))
        check(lib.dfb_nerf_heads_bwd(raw.detach().contiguous().data_ptr(), g.data_ptr(), P, Cc, gs.data_ptr(), gr.data_ptr(),
                                     gt.data_ptr() if Cc == 9 else None, None))
        assert torch.allclose(gs[:, 0].float(), x.grad[:, 3], rtol=8e-3, atol=1e-6) and bool((gs[:, 1:] == 0).all())
        assert torch.allclose(gr[:, :3].float(), x.grad[:, :3], rtol=8e-3, atol=1e-6) and bool((gr[:, 3:] == 0).all())
        if Cc == 9:
            assert torch.allclose(gt[:, :5].float(), x.grad[:, 4:9], rtol=8e-3, atol=1e-6) and bool((gt[:, 5:] == 0).all())


@pytest.mark.parametrize("N,S", [(1, 1), (37, 40), (1536, 128), (5000, 192)])
def test_fused_nerfw_loss_vs_tensor_expressions(N, S):
    """NerfWLoss on CUDA tensors (dfb_nerfw_loss_fwd / _bwd) against the same module's tensor expressions
    (models/losses.py:42-57) evaluated in float64 on the host: the four terms, the fine pass' mean squared error and the
    gradients w.r.t. rgb_coarse, rgb_fine, beta, transient_sigmas under unequal upstream weights."""
    from dfnet_b200.losses import NerfWLoss
    torch.manual_seed(N + S)
    lf = NerfWLoss(coef=0.7, lambda_u=0.02)
    host = dict(rgb_coarse=torch.rand(N, 3), rgb_fine=torch.rand(N, 3), beta=torch.rand(N) * 0.8 + 0.1, transient_sigmas=torch.rand(N, S) * 3)
    tg = torch.rand(N, 3)
    wts = dict(c_l=1.0, f_l=0.5, b_l=2.0, s_l=3.0)
    a = {k: v.to(dev()).requires_grad_(True) for k, v in host.items()}
    b = {k: v.double().requires_grad_(True) for k, v in host.items()}
    la, lb = lf(a, tg.to(dev())), lf(b, tg.double())
    assert lf.last_mse_fine is None                      # the float64 host call took the tensor-expression path
    la2 = lf(a, tg.to(dev()))
    assert lf.last_mse_fine is not None
    assert abs(float(lf.last_mse_fine) - float(((host["rgb_fine"] - tg) ** 2).mean())) < 1e-6
    for k in wts:
        assert abs(float(la[k]) - float(lb[k])) < 2e-6 * max(1.0, abs(float(lb[k]))), k
        assert float(la2[k]) == float(la[k])             # deterministic
    sum(wts[k] * la[k] for k in wts).backward()
    sum(wts[k] * lb[k] for k in wts).backward()
    for k in host:
        ga, gb = a[k].grad.cpu().double(), b[k].grad
        assert float((ga - gb).abs().max()) <= 2e-6 * float(gb.abs().max()) + 1e-12, k
    # only some inputs require gradient
    c = {k: v.to(dev()).requires_grad_(k == "beta") for k, v in host.items()}
    sum(lf(c, tg.to(dev())).values()).backward()
    assert c["rgb_fine"].grad is None and c["beta"].grad is not None
