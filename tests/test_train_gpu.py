"""GPU parity of the training-step kernels (row a15: reference feature/direct_feature_matching.py:322-390).

Checker: torch autograd in fp32/fp64 on the same tensors (a floating-point kernel family, so a torch reference is the
oracle here), plus the golden gradients generated from the reference's own `train_on_batch` on the CPU
(tests/golden/make_golden_train.py).  Tolerances: the gradient path is bf16 with fp32 accumulation (BASELINE config[3]
names bf16 for the training loop), so tensor-core gradients are compared at 3e-2 of the tensor's max and a cosine
similarity above 0.999; fp32 kernels (losses, resampling adjoints) at 1e-4.
"""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from helpers import synthetic_dfnet

pytestmark = pytest.mark.gpu
torch.backends.cudnn.allow_tf32 = False  # the checker runs in true fp32
torch.backends.cuda.matmul.allow_tf32 = False


def dev():
    return torch.device("cuda:0")


def _ops():
    from dfnet_b200 import ops
    return ops


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def close_grad(got, want, tol=3e-2, cos_min=0.999):
    got, want = got.double().flatten(), want.double().flatten()
    scale = float(want.abs().max()) + 1e-30
    err = float((got - want).abs().max()) / scale
    cos = float(F.cosine_similarity(got, want, dim=0))
    assert err < tol and cos > cos_min, (err, cos)
    return err, cos


@pytest.mark.parametrize("cin,cout,k,B,H,W", [(64, 128, 3, 2, 37, 45), (3, 64, 3, 1, 40, 56), (256, 512, 3, 1, 20, 24),
                                               (128, 64, 3, 3, 16, 8), (64, 128, 5, 1, 33, 20), (256, 64, 1, 1, 24, 24),
                                               # 1x1 layers: the pixel-major 128-byte-swizzled tiles (64-channel multiples), ragged
                                               # images, a NeRF layer ([P/8, 8] "image"), and a width that stays on the panel layout
                                               (128, 128, 1, 1, 1000, 8), (64, 128, 1, 2, 19, 13), (128, 64, 1, 1, 5, 3),
                                               (192, 256, 1, 1, 33, 20), (48, 64, 1, 1, 24, 24)])
def test_conv_wgrad_vs_torch(cin, cout, k, B, H, W):
    ops = _ops()
    torch.manual_seed(1)
    cpad = (cin + 7) // 8 * 8
    x = torch.randn(B, cin, H, W, device=dev()).bfloat16()
    go = (torch.randn(B, cout, H, W, device=dev()) * 0.1).bfloat16()
    x_nhwc = torch.zeros(B, H, W, cpad, device=dev(), dtype=torch.bfloat16)
    x_nhwc[..., :cin] = x.permute(0, 2, 3, 1)
    go_nhwc = go.permute(0, 2, 3, 1).contiguous()
    dW = torch.full((cout, cin, k, k), 7.0, device=dev())
    dB = torch.full((cout,), 7.0, device=dev())
    ops.check(ops.lib.dfb_conv_wgrad(_p(go_nhwc), _p(x_nhwc), B, H, W, cin, cpad, cout, k, 1, _p(dW), _p(dB), None))
    torch.cuda.synchronize()
    want = torch.nn.grad.conv2d_weight(x.double(), (cout, cin, k, k), go.double(), padding=k // 2)
    assert float((dW.double() - want).abs().max()) < 2e-4 * float(want.abs().max()) + 1e-5
    wb = go.double().sum((0, 2, 3))
    assert float((dB.double() - wb).abs().max()) < 1e-4 * float(wb.abs().max()) + 1e-5


@pytest.mark.parametrize("cin,cout,k,B,H,W", [(64, 128, 3, 2, 37, 45), (3, 64, 3, 1, 40, 56), (512, 512, 3, 1, 20, 24),
                                               (64, 128, 5, 1, 33, 20), (256, 64, 1, 1, 24, 24)])
def test_conv_dgrad_vs_torch(cin, cout, k, B, H, W):
    """Data-gradient convolution (transposed + flipped filter through the forward kernel) with the ReLU mask epilogue."""
    ops = _ops()
    torch.manual_seed(2)
    w = torch.randn(cout, cin, k, k, device=dev()) * 0.05
    sc = torch.rand(cout, device=dev()) + 0.5
    go = (torch.randn(B, cout, H, W, device=dev()) * 0.1).bfloat16()
    act = torch.relu(torch.randn(B, cin, H, W, device=dev())).half()
    h = C.c_void_p()
    ops.check(ops.lib.dfb_conv_create_ex(cin, cout, k, k, _p(w), None, _p(sc), None, 1, 1, C.byref(h)))
    try:
        go_nhwc = go.permute(0, 2, 3, 1).contiguous()
        cpo = (cin + 63) // 64 * 64
        want = torch.nn.grad.conv2d_input((B, cin, H, W), (w * sc.view(-1, 1, 1, 1)).bfloat16().double(), go.double(), padding=k // 2)
        if cin >= 64:
            mask = act.permute(0, 2, 3, 1).contiguous()
            out = torch.empty(B, H, W, cpo, device=dev(), dtype=torch.bfloat16)
            ops.check(ops.lib.dfb_conv_fwd_ex(h, _p(go_nhwc), B, H, W, 0, _p(out), None, None, _p(mask), None, None))
            torch.cuda.synchronize()
            want = want * (act > 0)
            got = out.permute(0, 3, 1, 2).double()
            assert float((got - want).abs().max()) < 1e-2 * float(want.abs().max())
        else:  # conv1_1: fp32 NCHW image gradient, 3 real channels
            out = torch.empty(B, cin, H, W, device=dev())
            ops.check(ops.lib.dfb_conv_fwd_ex(h, _p(go_nhwc), B, H, W, 0, None, None, _p(out), None, None, None))
            torch.cuda.synchronize()
            assert float((out.double() - want).abs().max()) < 2e-4 * float(want.abs().max())
    finally:
        ops.lib.dfb_conv_destroy(h)


def test_loss_and_resample_backward_vs_torch():
    from dfnet_b200 import dfnet as D, misc
    torch.manual_seed(3)
    # cosine feature loss (per_channel=False: cosine over the pixels of each channel)
    fr = torch.randn(128, 60 * 80, device=dev(), requires_grad=True)
    ft = torch.randn(128, 60 * 80, device=dev())
    D.feature_loss(fr, ft).backward()
    fr2 = fr.detach().clone().requires_grad_(True)
    (1 - torch.nn.CosineSimilarity(dim=1, eps=1e-6)(fr2, ft).mean()).backward()
    assert float((fr.grad - fr2.grad).abs().max()) < 1e-4 * float(fr2.grad.abs().max())
    # the ABI's two ways of getting the row statistics: recomputed (flag 0) and left in the workspace by the forward (flag 2,
    # what the autograd function above uses) give the same gradient
    from dfnet_b200._lib import check, lib
    a_, b_ = fr.detach().contiguous(), ft.contiguous()
    g1, g2, one = torch.empty_like(a_), torch.empty_like(a_), torch.ones((), device=dev())
    ws1, ws2, loss = torch.empty(128 * 64 * 3, device=dev()), torch.empty(128 * 64 * 3, device=dev()), torch.empty((), device=dev())
    check(lib.dfb_cosine_loss_bwd(a_.data_ptr(), b_.data_ptr(), 128, 4800, 0, 1e-6, one.data_ptr(), g1.data_ptr(), ws1.data_ptr(), ws1.numel() * 4, None))
    check(lib.dfb_cosine_loss(a_.data_ptr(), b_.data_ptr(), 128, 4800, 0, 1e-6, loss.data_ptr(), ws2.data_ptr(), ws2.numel() * 4, None))
    check(lib.dfb_cosine_loss_bwd(a_.data_ptr(), b_.data_ptr(), 128, 4800, 2, 1e-6, one.data_ptr(), g2.data_ptr(), ws2.data_ptr(), ws2.numel() * 4, None))
    assert torch.equal(g1, g2) and torch.equal(g1, fr.grad)
    # MSE
    a = torch.rand(1, 3, 48, 64, device=dev(), requires_grad=True)
    b = torch.rand(1, 3, 48, 64, device=dev())
    (misc.img2mse(a, b) * 3.0).backward()
    a2 = a.detach().clone().requires_grad_(True)
    (F.mse_loss(a2, b) * 3.0).backward()
    assert float((a.grad - a2.grad).abs().max()) < 1e-5 * float(a2.grad.abs().max())
    # bicubic x4 (train_on_batch half_res) and bilinear align_corners
    for fn, ref in ((misc.upsample_bicubic, lambda t, s: F.interpolate(t, size=s, mode="bicubic", align_corners=False)),
                    (misc.upsample_bilinear_ac, lambda t, s: F.interpolate(t, size=s, mode="bilinear", align_corners=True))):
        x = torch.rand(1, 3, 15, 20, device=dev(), requires_grad=True)
        w = torch.randn(1, 3, 60, 83, device=dev())
        (fn(x, (60, 83)) * w).sum().backward()
        x2 = x.detach().clone().requires_grad_(True)
        (ref(x2, (60, 83)) * w).sum().backward()
        assert float((x.grad - x2.grad).abs().max()) < 1e-4 * float(x2.grad.abs().max())


def _ste(t, dt):
    """Round to the kernel's storage type in the forward, identity in the backward."""
    return t if dt is None else t + (t.to(dt).float() - t).detach()


def _pin(t, acts, key):
    """Substitute the value the GPU forward stored (identity in the backward): the torch graph is then differentiated at
    exactly the kernels' forward state (same ReLU masks, same max-pool arg-maxima)."""
    return t if acts is None or key not in acts else t + (acts[key] - t).detach()


def torch_dfnet_forward(net, x, return_feature, single, return_pose, upH, upW, quant=None, acts=None):
    """torch restatement of reference feature/dfnet.py:106-172 on the mirror's own nn layers (eval BatchNorm).
    quant=torch.float16 / torch.bfloat16 additionally rounds weights and stored activations exactly where the kernels
    do (straight-through), so that ReLU masks and max-pool arg-maxima are those of the function the GPU path actually
    evaluates; quant=None is the plain fp32 reference."""
    mean = torch.tensor(net.mean, device=x.device).view(1, 3, 1, 1)
    std = torch.tensor(net.std, device=x.device).view(1, 3, 1, 1)
    h = _pin(_ste((x - mean) / std, quant), acts, "in")
    taps = {2: 0, 14: 1, 28: 2}
    feats = []
    ci = 0
    for i, m in enumerate(net.encoder):
        if isinstance(m, torch.nn.ReLU):
            h = _pin(_ste(F.relu(h), quant), acts, f"act{ci}")
            ci += 1
        elif isinstance(m, torch.nn.Conv2d):
            h = F.conv2d(h, _ste(m.weight, quant), m.bias, padding=1)
        else:
            h = m(h)
        if i in taps and taps[i] < len(net.hypercolumn_layers):
            feats.append(h)
    pose = net.fc_pose(h.mean((2, 3))) if return_pose else None
    if not return_feature:
        return None, pose
    outs = []
    for l, f in enumerate(feats):
        seq = getattr(net.adaptation_layers, f"adapt_layer_{l}")
        a = F.conv2d(_pin(_ste(f, quant), acts, f"tap{l}"), _ste(seq[0].weight, quant), seq[0].bias)
        a = _pin(_ste(F.relu(a), quant), acts, f"mid{l}")
        a = seq[3](F.conv2d(a, _ste(seq[2].weight, quant), seq[2].bias, padding=2))
        outs.append(F.interpolate(a, size=(upH, upW), mode="bilinear", align_corners=True))
    st = torch.stack(outs)
    if single:
        return [st], pose
    B = x.shape[0] // 2
    return [st[:, :B], st[:, B:]], pose


def _randomise_bn(net, seed):
    g = torch.Generator().manual_seed(seed)
    for l in range(len(net.hypercolumn_layers)):
        bn = getattr(net.adaptation_layers, f"adapt_layer_{l}")[3]
        bn.running_mean.copy_(torch.randn(128, generator=g) * 0.1)
        bn.running_var.copy_(torch.rand(128, generator=g) + 0.5)
        bn.weight.data.copy_(torch.rand(128, generator=g) + 0.5)
        bn.bias.data.copy_(torch.randn(128, generator=g) * 0.1)


@pytest.mark.parametrize("cls,levels,H,W", [("DFNet", [0, 1, 2], 64, 96), ("DFNet", [0], 48, 80), ("DFNet_s", [0], 50, 70),
                                             ("DFNet", [1, 2], 64, 64)])
def test_feature_loss_gradient_wrt_rendered_image(cls, levels, H, W):
    """d feature_loss / d rgb through the frozen feature net (siamese forward on cat([data, rgb]), only the rendered
    stream differentiated): the gradient train_on_batch sends back into the renderer."""
    from dfnet_b200 import dfnet as D
    net = synthetic_dfnet(cls, seed=4).to(dev()).eval()
    _randomise_bn(net, 5)
    for p in net.parameters():
        p.requires_grad_(False)
    net.grad_levels = levels
    torch.manual_seed(6)
    data = torch.rand(1, 3, H, W, device=dev())
    rgb0 = (data + 0.1 * torch.randn_like(data)).clamp(0, 1)

    def loss_of(fwd, rgb):
        feats, _ = fwd(torch.cat([data, rgb]))
        idx = torch.tensor(levels, device=dev())
        fr = D.preprocess_features_for_loss(torch.index_select(feats[1], 0, idx))
        ft = D.preprocess_features_for_loss(torch.index_select(feats[0], 0, idx))
        return fr, ft

    rgb = rgb0.clone().requires_grad_(True)
    fr, ft = loss_of(lambda x: net(x, return_feature=True, isSingleStream=False, return_pose=False, upsampleH=H, upsampleW=W), rgb)
    loss = D.feature_loss(fr[0], ft[0])
    loss.backward()
    # (a) differentiated at the kernels' own forward state (stored activations pinned): tight, this is the kernel check;
    # (b) fp16-rounded and (c) plain fp32 torch forwards: ReLU-mask / arg-max flips of the independently rounded
    #     forwards accumulate with depth, so these are sanity bounds on the end-to-end gradient
    acts = net._handle.tape_activations()
    for quant, pin, tol, cos_min in ((torch.float16, acts, 3e-2, 0.9998), (torch.float16, None, 0.5, 0.995), (None, None, 0.5, 0.99)):
        rgb_t = rgb0.clone().requires_grad_(True)
        fr_t, ft_t = loss_of(lambda x: torch_dfnet_forward(net, x, True, False, False, H, W, quant, pin), rgb_t)
        loss_t = 1 - torch.nn.CosineSimilarity(dim=1, eps=1e-6)(fr_t[0].reshape(fr_t.shape[1], -1),
                                                                ft_t[0].reshape(ft_t.shape[1], -1)).mean()
        loss_t.backward()
        assert abs(float(loss.detach()) - float(loss_t.detach())) < 2e-3 * abs(float(loss_t.detach())) + 1e-5
        print("feature-loss grad err/cos vs", quant, "pinned" if pin else "free", close_grad(rgb.grad, rgb_t.grad, tol=tol, cos_min=cos_min))


@pytest.mark.parametrize("B,H,W,dtype", [(1, 64, 96, "f16"), (2, 60, 80, "f16"), (1, 64, 96, "bf16"), (2, 60, 80, "bf16")])
def test_pose_regressor_parameter_gradients(B, H, W, dtype):
    """d PoseLoss / d (encoder, fc_pose parameters) of the pose regressor: weight, bias and data-gradient kernels
    chained through all 13 convolutions and 5 poolings."""
    from dfnet_b200 import misc
    net = synthetic_dfnet("DFNet", seed=7).to(dev())
    net.train()
    net.train_dtype = dtype
    qdt = torch.float16 if dtype == "f16" else torch.bfloat16
    for m in net.modules():  # freeze_bn_layer_train (reference feature/direct_feature_matching.py:52-61)
        if isinstance(m, torch.nn.BatchNorm2d):
            m.eval()
            m.weight.requires_grad_(False), m.bias.requires_grad_(False)
    torch.manual_seed(8)
    x = torch.rand(B, 3, H, W, device=dev())
    target = torch.randn(B, 12, device=dev())
    _, pose = net(x, return_feature=False, isSingleStream=True, return_pose=True, upsampleH=H, upsampleW=W)
    misc.mse(pose, target).backward()
    got = {n: p.grad.clone() for n, p in net.named_parameters() if p.grad is not None}
    # (a) differentiated at the kernels' own forward state (stored bf16 activations pinned): tight, the kernel check;
    # (b) plain fp32: arg-max / mask flips of the bf16 forward accumulate with depth, so only a sanity bound
    acts = net._handle.tape_activations()
    for quant, pin, tol, cos_min in ((qdt, acts, 5e-2, 0.9995), (None, None, 1.0, 0.85 if dtype == "bf16" else 0.97)):
        net.zero_grad()
        _, pose_t = torch_dfnet_forward(net, x, False, True, True, H, W, quant, pin)
        F.mse_loss(pose_t, target).backward()
        assert float((pose - pose_t).abs().max()) < 2e-2 * float(pose_t.abs().max())
        names = [n for n, p in net.named_parameters() if p.grad is not None]
        assert set(names) == set(got), (set(names) ^ set(got))
        worst = (0.0, 1.0, "")
        for n, p in net.named_parameters():
            if p.grad is None:
                continue
            err, cos = close_grad(got[n], p.grad, tol=tol, cos_min=cos_min)
            if 1 - cos > 1 - worst[1]:
                worst = (err, cos, n)
        print("pose-regressor parameter gradients vs", quant, "worst (err, cos, name):", worst)


class _Recorder:
    """Stands in for the optimizer (as in tests/golden/make_golden_train.py): keeps the gradients of the step."""

    def __init__(self, model):
        self.model, self.grads = model, None

    def step(self):
        self.grads = {n: p.grad.detach().clone() for n, p in self.model.named_parameters() if p.grad is not None}

    def zero_grad(self):
        self.model.zero_grad()


@pytest.mark.parametrize("case,mma", [("lvl0", "fp32"), ("lvl012", "fp32"), ("lvl0", "f16"), ("lvl012", "f16")])
def test_train_on_batch_vs_reference_golden(case, mma):
    """One full step (pose regressor -> SVD -> render at H//4 -> bicubic x4 -> feature net -> losses -> backward) against
    the gradients the reference's own train_on_batch produced on the CPU in fp32 (tests/golden/make_golden_train.py).
    Loss / PSNR: 1e-3.  Gradients: the GPU step evaluates the networks with fp16 storage, so ReLU masks and max-pool
    arg-maxima differ from the fp32 run on a small fraction of elements and the difference grows with depth (see
    test_pose_regressor_parameter_gradients for the kernel-level check at a pinned forward state); here the norm of every
    parameter gradient must agree within 10 % and its direction (cosine over the stored subsample) within 0.95."""
    import os
    from helpers import pose_head_init_, sd_checksum, synthetic_nets, train_case
    from dfnet_b200 import direct_feature_matching as dfm
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "train_golden.npz"))
    cfg = train_case(case)
    Fnet = pose_head_init_(synthetic_dfnet("DFNet", seed=0))
    Gnet = synthetic_dfnet("DFNet", seed=1)
    assert sd_checksum(Fnet.state_dict()) == bytes(g[f"{case}_F_sha"]).decode()
    assert sd_checksum(Gnet.state_dict()) == bytes(g[f"{case}_G_sha"]).decode()
    Fnet, Gnet = Fnet.to(dev()), Gnet.to(dev()).eval()
    for p in Gnet.parameters():
        p.requires_grad_(False)
    Fnet.train()
    for m in Fnet.modules():  # freeze_bn_layer_train
        if isinstance(m, torch.nn.BatchNorm2d):
            m.eval()
            m.weight.requires_grad_(False), m.bias.requires_grad_(False)
    mods, _ = synthetic_nets(cfg["D"], cfg["W"])
    c, f, ea, et = [m.to(dev()) for m in mods]
    for m in (c, f, ea, et):
        for p in m.parameters():
            p.requires_grad_(False)
    kw = dict(network_query_fn=None, perturb=0.0, N_importance=cfg["Nf"], network_fine=f, N_samples=cfg["Nc"], network_fn=c,
              use_viewdirs=True, white_bkgd=False, raw_noise_std=0.0, embedding_a=ea, embedding_t=et, test_time=True,
              ndc=False, lindisp=False, near=cfg["near"], far=cfg["far"], mma=mma)   # f16: the product default (tcgen05
    rec = _Recorder(Fnet)                                                              # forward with saved masks + dfb_render_bwd_saved)
    loss, psnr = dfm.train_on_batch(cfg["args"], cfg["data"], Fnet, Gnet, cfg["pose"], cfg["hist"], cfg["hwf"], rec, True, dev(),
                                    cfg["world"], **kw)
    want_loss, want_psnr = float(g[f"{case}_loss"].reshape(-1)[0]), float(g[f"{case}_psnr"].reshape(-1)[0])
    assert abs(float(loss.reshape(-1)[0]) - want_loss) < 1e-3 * abs(want_loss), (loss, want_loss)
    assert abs(float(np.asarray(psnr).reshape(-1)[0]) - want_psnr) < 1e-3 * abs(want_psnr), (psnr, want_psnr)
    names = bytes(g[f"{case}_grad_names"]).decode().split("\n")
    assert sorted(rec.grads) == names
    worst_cos, worst_norm = (1.0, ""), (0.0, "")
    for n in names:
        gg = rec.grads[n].flatten()
        sub = gg[:: max(1, gg.numel() // 4096)][:4096].double().cpu()
        want = torch.from_numpy(g[f"{case}_g_{n}_sub"]).double()
        cos = float(F.cosine_similarity(sub, want, dim=0))
        nr = abs(float(gg.norm()) / float(g[f"{case}_g_{n}_stats"][0]) - 1.0)
        worst_cos = min(worst_cos, (cos, n))
        worst_norm = max(worst_norm, (nr, n))
    print(case, mma, "loss", float(loss.reshape(-1)[0]), want_loss, "worst cos", worst_cos, "worst norm dev", worst_norm)
    assert worst_cos[0] > 0.95 and worst_norm[0] < 0.10, (worst_cos, worst_norm)


@pytest.mark.parametrize("bn_mode", ["train", "eval"])
def test_dfnet_head_and_encoder_training_gradients(bn_mode):
    """run_feature.py training: ONE siamese forward returns features and pose, the loss reads both, every
    parameter is trained - encoder, adaptation heads (1x1, 5x5, BatchNorm affine), fc_pose.  dfb_dfnet_bwd with the head
    tape (BatchNorm backward with batch statistics or running statistics, head weight gradients, combined feature + pose
    gradient through the encoder) against torch autograd pinned to the kernels' forward state, and against plain fp32."""
    from dfnet_b200 import misc
    net = synthetic_dfnet("DFNet", seed=5).to(dev())
    with torch.no_grad():
        _randomise_bn(net, 9)
    net.train()
    if bn_mode == "eval":
        for m in net.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.eval()
    B, H, W = 2, 48, 64
    torch.manual_seed(12)
    x = torch.rand(2 * B, 3, H, W, device=dev())
    wt = torch.randn(3, B, 128, H, W, device=dev()) / (3 * B * 128 * H * W) ** 0.5
    wr = torch.randn(3, B, 128, H, W, device=dev()) / (3 * B * 128 * H * W) ** 0.5
    target = torch.randn(2 * B, 12, device=dev())

    def loss_of(feats, pose, mse):
        return (feats[0] * wt).sum() + (feats[1] * wr).sum() + 0.5 * mse(pose, target)
    feats, pose = net(x, return_feature=True, isSingleStream=False, return_pose=True, upsampleH=H, upsampleW=W)
    loss_of(feats, pose, misc.mse).backward()
    got = {n: p.grad.clone() for n, p in net.named_parameters() if p.grad is not None}
    assert len(got) == len(list(net.parameters()))
    acts = net._handle.tape_activations()
    for quant, pin, tol, cos_min in ((torch.float16, acts, 5e-2, 0.999), (None, None, 1.0, 0.97)):
        net.zero_grad()
        f_t, pose_t = torch_dfnet_forward(net, x, True, False, True, H, W, quant, pin)
        loss_of(f_t, pose_t, F.mse_loss).backward()
        worst = (0.0, 1.0, "")
        for n, p in net.named_parameters():
            if bn_mode == "train" and "adapt_layer" in n and n.endswith(".2.bias"):
                # batch statistics remove the mean: d loss / d (5x5 bias) is exactly 0 (torch: ~1e-9 of rounding); the
                # kernels' bf16 gradient operands leave a residue that must be negligible against the layer's weight gradient
                assert float(got[n].abs().max()) < 2e-3 * float(got[n.replace(".bias", ".weight")].abs().max()), n
                continue
            err, cos = close_grad(got[n], p.grad, tol=tol, cos_min=cos_min)
            if 1 - cos > 1 - worst[1]:
                worst = (err, cos, n)
        print(bn_mode, "head + encoder parameter gradients vs", quant, "worst (err, cos, name):", worst)


def test_batched_pose_render_and_virtual_views():
    """8f-2: dfb_render_poses_fwd renders a batch of poses in one call, bit-identical to per-pose renders;
    render_virtual_imgs (feature/misc.py:249-289) on top of it with the tiny-image + bicubic path."""
    import types
    from helpers import synthetic_nets
    from dfnet_b200 import misc, ops, rendering
    mods, _ = synthetic_nets(8, 256)
    c, f, ea, et = [m.to(dev()) for m in mods]
    h = ops.NerfHandle(c, f, ea, et)
    rng = np.random.RandomState(5)
    n, H, W, focal = 5, 24, 40, 35.0
    poses = np.tile(np.eye(4, dtype=np.float32)[None, :3], (n, 1, 1))
    poses[:, :, 3] = rng.randn(n, 3) * 0.1 + np.array([0, 0, 1.0])
    hists = np.stack([np.roll(np.array([5., 10, 20, 30, 15, 10, 5, 3, 1, 1], np.float32), i) for i in range(n)])
    o = h.render_poses(16, 32, torch.tensor(poses, device=dev()), torch.tensor(hists, device=dev()), H, W, focal, 0.0, 2.5, mma="f16")
    for i in range(n):
        one = h.render(16, 32, True, c2w=torch.tensor(poses[i], device=dev()), H=H, W=W, focal=focal, near=0.0, far=2.5,
                       hist=torch.tensor(hists[i], device=dev()), mma="f16")
        assert torch.equal(o["rgb"][i].reshape(-1, 3), one["rgb"]) and torch.equal(o["disp"][i].reshape(-1), one["disp"])
    kw = dict(network_query_fn=None, perturb=False, N_importance=32, network_fine=f, N_samples=16, network_fn=c, use_viewdirs=True,
              white_bkgd=False, raw_noise_std=0.0, embedding_a=ea, embedding_t=et, test_time=True, ndc=False, lindisp=False,
              near=0.0, far=2.5)
    args = types.SimpleNamespace(tinyimg=True, tinyscale=4.0, chunk=32768)
    world = dict(pose_scale=1.0, pose_scale2=1.0, move_all_cam_vec=[0.0, 0.0, 0.0])
    Hf, Wf = 96, 160
    rgbs = misc.render_virtual_imgs(args, torch.tensor(poses), torch.tensor(hists)[:, None], (Hf, Wf, focal * 4), dev(), kw, world)
    assert rgbs.shape == (n, Hf, Wf, 3) and not rgbs.is_cuda
    ref, _, _, _ = rendering.render(Hf // 4, Wf // 4, focal, c2w=torch.tensor(poses[2], device=dev()),
                                    img_idx=torch.tensor(hists[2], device=dev()), **kw)
    want = torch.nn.Upsample(size=(Hf, Wf), mode="bicubic")(ref[None].permute(0, 3, 1, 2))[0].permute(1, 2, 0).cpu()
    assert float((rgbs[2] - want).abs().max()) < 2e-5


def test_dfnet_feature_training_step_reduces_the_loss():
    """run_feature.py's step (pose loss + triplet loss on a siamese forward, RVS pose loss), a few Adam steps."""
    import types
    from dfnet_b200 import feature_train
    net = synthetic_dfnet("DFNet", seed=2).to(dev())
    net.train()
    opt = torch.optim.Adam(net.parameters(), lr=1e-4)
    args = types.SimpleNamespace(tripletloss=True, triplet_margin=1.0, combine_loss_w=[1.0, 1.0, 1.0], featurenet_batch_size=2,
                                 freezeBN=False)
    torch.manual_seed(3)
    n, H, W = 6, 48, 64
    targets, rgbs, virt = torch.rand(n, H, W, 3), torch.rand(n, H, W, 3), torch.rand(n, H, W, 3)
    rgbs = 0.7 * targets + 0.3 * rgbs
    poses = torch.randn(n, 3, 4) * 0.3
    poses_p = poses + 0.05 * torch.randn(n, 3, 4)
    np.random.seed(0)
    hist = [feature_train.train_on_batch_with_random_view_synthesis(args, targets, rgbs, poses, virt, poses_p, net, n, None, opt, (H, W, 60.0))
            for _ in range(6)]
    print("DFNet training epoch losses", hist)
    assert np.isfinite(hist).all() and hist[-1] < hist[0]
    plain = feature_train.train_on_batch(args, targets, rgbs, poses, net, n, None, opt, (H, W, 60.0))
    assert np.isfinite(plain)
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            assert int(m.num_batches_tracked) > 0          # train-mode BatchNorm kept its running statistics


def test_polar_orthogonalize_vs_torch_svd():
    """svd_reg of the pose regressor (reference feature/direct_feature_matching.py:81-86): the device kernel against
    u @ v^T of torch.svd, forward and the gradient through it (float64 autograd), on near-rotations (what the regressor
    predicts), scaled / sheared matrices and a reflection."""
    from dfnet_b200.misc import polar_orthogonalize
    torch.manual_seed(0)
    n = 64
    q, _ = torch.linalg.qr(torch.randn(n, 3, 3, dtype=torch.float64))
    A = q + 0.2 * torch.randn(n, 3, 3, dtype=torch.float64)
    A[:8] *= torch.linspace(0.05, 20.0, 8, dtype=torch.float64)[:, None, None]
    A[8] = torch.diag(torch.tensor([1.0, 1.0, -1.0], dtype=torch.float64)) + 0.05 * torch.randn(3, 3, dtype=torch.float64)
    G = torch.randn(n, 3, 3, dtype=torch.float64)
    a64 = A.clone().requires_grad_(True)
    u, s, v = torch.svd(a64)
    want = u @ v.transpose(-2, -1)
    (want * G).sum().backward()
    a32 = A.float().to(dev()).requires_grad_(True)
    got = polar_orthogonalize(a32)
    (got * G.float().to(dev())).sum().backward()
    torch.cuda.synchronize()
    assert float((got.detach().cpu().double() - want.detach()).abs().max()) < 2e-6
    eye = got.detach().cpu().double() @ got.detach().cpu().double().transpose(-2, -1)
    assert float((eye - torch.eye(3, dtype=torch.float64)).abs().max()) < 2e-6
    gw = a64.grad
    err = float((a32.grad.cpu().double() - gw).abs().max()) / float(gw.abs().max())
    assert err < 1e-5, err


@pytest.mark.parametrize("H,W,shape", [(120, 160, (3, 4)), (37, 53, (4, 4)), (1, 1, (3, 4))])
def test_pose_rays_and_adjoint_vs_torch(H, W, shape):
    """dfb_pose_rays_fwd / _bwd (get_rays of a pose that carries gradient + view-direction normalisation, ray_utils.py:5-15,
    rendering.py:366-370) against the tensor expressions and torch autograd."""
    from dfnet_b200 import rendering
    torch.manual_seed(5)
    focal = 73.5
    base = torch.eye(4, device=dev())[:shape[0]].clone()
    base[:3, :3] += 0.1 * torch.randn(3, 3, device=dev())
    base[:3, 3] = torch.tensor([0.3, -0.2, 1.5], device=dev())
    a, b = base.clone().requires_grad_(True), base.clone().requires_grad_(True)
    o1, d1, v1 = rendering._PoseRaysFn.apply(a, H, W, focal)
    o2, d2 = rendering._get_rays_torch(H, W, focal, b)
    o2, d2 = o2.reshape(-1, 3), d2.reshape(-1, 3)
    v2 = d2 / torch.norm(d2, dim=-1, keepdim=True)
    for x, y in ((o1, o2), (d1, d2), (v1, v2)):
        assert float((x - y).abs().max()) <= 2e-6 * float(y.abs().max())
    wo, wd, wv = torch.randn_like(o1), torch.randn_like(d1), torch.randn_like(v1)
    ((o1 * wo).sum() + (d1 * wd).sum() + (v1 * wv).sum()).backward()
    ((o2 * wo).sum() + (d2 * wd).sum() + (v2 * wv).sum()).backward()
    assert a.grad.shape == b.grad.shape
    assert float((a.grad - b.grad).abs().max()) <= 2e-5 * float(b.grad.abs().max()) + 1e-6
    # only the view directions carry gradient
    c = base.clone().requires_grad_(True)
    _, _, v3 = rendering._PoseRaysFn.apply(c, H, W, focal)
    e = base.clone().requires_grad_(True)
    _, d4 = rendering._get_rays_torch(H, W, focal, e)
    d4 = d4.reshape(-1, 3)
    (v3 * wv).sum().backward()
    ((d4 / torch.norm(d4, dim=-1, keepdim=True)) * wv).sum().backward()
    assert float((c.grad - e.grad).abs().max()) <= 2e-5 * float(e.grad.abs().max()) + 1e-6
