"""Shared test helpers: synthetic networks as numpy parameter dicts for the oracle."""
import hashlib

import numpy as np
import torch

from dfnet_b200 import nerfw as my_nerfw


def sd_checksum(sd):
    h = hashlib.sha256()
    for k in sd:
        h.update(k.encode())
        h.update(sd[k].detach().cpu().numpy().tobytes())
    return h.hexdigest()


def np_params(mod):
    return {k: v.detach().cpu().numpy() for k, v in mod.state_dict().items()}


_CACHE = {}


def synthetic_nets(D, W, fine=True):
    """(torch modules, oracle dict) for the seeded synthetic networks of SURVEY §8(d)."""
    key = (D, W, fine)
    if key not in _CACHE:
        state = torch.get_rng_state()
        c, f, ea, et = my_nerfw.make_synthetic_nerf(D=D, W=W, fine=fine)
        torch.set_rng_state(state)
        nets = dict(coarse=np_params(c), fine=np_params(f) if f is not None else None,
                    emb_a=ea.weight.detach().numpy(), emb_t=et.weight.detach().numpy(),
                    D=D, skips=(4,), beta_min=0.1)
        _CACHE[key] = ((c, f, ea, et), nets)
    return _CACHE[key]


def rel_err(a, b, floor=1e-3):
    """max |a-b| / max(|b|, floor): the relative-error measure used for the 1e-3 bar."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor))) if a.size else 0.0


def synthetic_dfnet(cls_name="DFNet", seed=0):
    """dfnet_b200 DFNet/DFNet_s with the weights the reference constructor produces under `seed`
    when torchvision's vgg16 is built without pretrained weights (test infrastructure: pretrained
    ImageNet weights are not available offline)."""
    import torchvision
    from dfnet_b200 import dfnet as my_dfnet
    state = torch.get_rng_state()
    torch.manual_seed(seed)
    vgg = torchvision.models.vgg16(weights=None)
    net = getattr(my_dfnet, cls_name)()          # builds encoder / heads / fc in the reference's order
    # the reference takes the encoder from torchvision and only then creates heads and fc_pose:
    # rebuild those so that they consume the RNG stream exactly like the reference constructor
    torch.manual_seed(seed)
    vgg = torchvision.models.vgg16(weights=None)
    net.encoder.load_state_dict(vgg.features.state_dict())
    net.adaptation_layers = my_dfnet.AdaptLayers(net.hypercolumn_layers, 128)
    net.fc_pose = torch.nn.Linear(512, 12)
    torch.set_rng_state(state)
    return net.eval()


def pose_head_init_(net):
    """Deterministic fc_pose so that a randomly initialised pose regressor predicts a camera that looks at the
    synthetic scene: small weights, bias = a slightly non-orthogonal [R | t] (svd_reg has something to fix)."""
    g = torch.Generator().manual_seed(123)
    with torch.no_grad():
        net.fc_pose.weight.copy_(torch.randn(12, 512, generator=g) * 2e-3)
        net.fc_pose.bias.copy_(torch.tensor([0.98, 0.02, 0.15, 0.2, -0.03, 1.01, 0.04, -0.1, -0.16, -0.02, 0.97, 1.9]))
    return net


def train_case(case):
    """Inputs of the train_on_batch golden cases (shared by tests/golden/make_golden_train.py and the GPU test)."""
    import types
    rng = np.random.RandomState(21)
    H, W = 64, 96
    args = types.SimpleNamespace(
        DFNet=True, preprocess_ImgNet=False, svd_reg=True, combine_loss=True, per_channel=False, chunk=32768, batch_size=1,
        combine_loss_w=[0.3, 0.2, 1.0] if case == "lvl012" else [0.0, 0.0, 1.0],
        feature_matching_lvl=[0, 1, 2] if case == "lvl012" else [0])
    data = torch.from_numpy(rng.rand(1, 3, H, W).astype(np.float32))
    pose = torch.from_numpy(np.array([[1, 0, 0, 0.1, 0, 1, 0, -0.05, 0, 0, 1, 2.0]], np.float32))
    hist = torch.from_numpy(np.array([[5, 10, 20, 30, 15, 10, 5, 3, 1, 1]], np.float32))
    world = dict(pose_scale=0.5, pose_scale2=1.0, move_all_cam_vec=[0.0, 0.0, 0.05])
    return dict(args=args, data=data, pose=pose, hist=hist, hwf=(H, W, 80.0), world=world, D=8, W=64, Nc=16, Nf=24,
                near=0.0, far=2.5)


def pad_nerfw_state_dict(P, W, W2=256, in_xyz=63):
    """Python restatement of the library's `tc_pad_params` (csrc/mlp_tc.cu): the state dict of the 8-layer network of
    width W2 that computes exactly the same function as the width-W network P (extra hidden units have zero input
    weights and zero bias, so they stay at relu(0) = 0 and feed nothing).  numpy arrays, reference parameter names."""
    H, H2 = W // 2, W2 // 2
    Q = {}

    def pad(w, rows2, cols2, col_shift_from=None, shift=0):
        out = np.zeros((rows2, cols2), np.float32)
        r, c = w.shape
        if col_shift_from is None:
            out[:r, :c] = w
        else:   # columns >= col_shift_from (the ray-constant inputs) move behind the widened hidden block
            out[:r, :col_shift_from] = w[:, :col_shift_from]
            out[:r, col_shift_from + shift:c + shift] = w[:, col_shift_from:]
        return out

    def pad1(b, n2):
        out = np.zeros(n2, np.float32)
        out[:b.shape[0]] = b
        return out

    for i in range(8):
        w, b = P[f"xyz_encoding_{i+1}.0.weight"], P[f"xyz_encoding_{i+1}.0.bias"]
        cols2 = in_xyz if i == 0 else (in_xyz + W2 if i == 4 else W2)
        Q[f"xyz_encoding_{i+1}.0.weight"], Q[f"xyz_encoding_{i+1}.0.bias"] = pad(w, W2, cols2), pad1(b, W2)
    Q["xyz_encoding_final.weight"], Q["xyz_encoding_final.bias"] = pad(P["xyz_encoding_final.weight"], W2, W2), pad1(P["xyz_encoding_final.bias"], W2)
    wd = P["dir_encoding.0.weight"]
    Q["dir_encoding.0.weight"] = pad(wd, H2, W2 + wd.shape[1] - W, col_shift_from=W, shift=W2 - W)
    Q["dir_encoding.0.bias"] = pad1(P["dir_encoding.0.bias"], H2)
    Q["static_sigma.0.weight"], Q["static_sigma.0.bias"] = pad(P["static_sigma.0.weight"], 1, W2), P["static_sigma.0.bias"]
    Q["static_rgb.0.weight"], Q["static_rgb.0.bias"] = pad(P["static_rgb.0.weight"], 3, H2), P["static_rgb.0.bias"]
    if "transient_encoding.0.weight" in P:
        wt = P["transient_encoding.0.weight"]
        Q["transient_encoding.0.weight"] = pad(wt, H2, W2 + wt.shape[1] - W, col_shift_from=W, shift=W2 - W)
        Q["transient_encoding.0.bias"] = pad1(P["transient_encoding.0.bias"], H2)
        for k in (2, 4, 6):
            Q[f"transient_encoding.{k}.weight"] = pad(P[f"transient_encoding.{k}.weight"], H2, H2)
            Q[f"transient_encoding.{k}.bias"] = pad1(P[f"transient_encoding.{k}.bias"], H2)
        for nm, r in (("transient_sigma", 1), ("transient_rgb", 3), ("transient_beta", 1)):
            Q[f"{nm}.0.weight"], Q[f"{nm}.0.bias"] = pad(P[f"{nm}.0.weight"], r, H2), P[f"{nm}.0.bias"]
    return Q


def torch_nerfw_forward(m, x, mode):
    """Plain torch restatement of NeRFW.forward (reference models/nerfw.py:297-354) on the module's own layers; test
    infrastructure for fitting a field (dfnet_b200's NeRFW.forward itself runs on the CUDA kernels and is not
    differentiable w.r.t. the weights).  mode: "sigma" | "static" | "full"."""
    ixyz = x[:, :m.in_channels_xyz]
    h = ixyz
    for i in range(m.D):
        if i in m.skips:
            h = torch.cat([ixyz, h], 1)
        h = getattr(m, f"xyz_encoding_{i + 1}")(h)
    sigma = m.static_sigma(h)
    if mode == "sigma":
        return sigma
    final = m.xyz_encoding_final(h)
    nd = m.in_channels_dir + m.in_channels_a
    de = m.dir_encoding(torch.cat([final, x[:, m.in_channels_xyz:m.in_channels_xyz + nd]], 1))
    static = torch.cat([m.static_rgb(de), sigma], 1)
    if mode == "static":
        return static
    t = m.transient_encoding(torch.cat([final, x[:, m.in_channels_xyz + nd:]], 1))
    return torch.cat([static, m.transient_rgb(t), m.transient_sigma(t), m.transient_beta(t)], 1)


def _embed_t(x, L):
    out = [x]
    for l in range(L):
        out += [torch.sin(x * 2.0 ** l), torch.cos(x * 2.0 ** l)]
    return torch.cat(out, -1)


def fit_synthetic_scene(D=8, W=256, steps=400, batch=8192, device="cpu", seed=0):
    """A "trained-like" NeRF-Hist field: the seeded default-initialised networks regressed (Adam, fp32, plain torch) onto
    an analytic scene in front of the benchmark camera - three solid objects with sharp density steps (sigma 0 / 40) and a
    position-dependent colour with high-frequency stripes, a faint transient fog.  Returns (coarse, fine, emb_a, emb_t)
    like make_synthetic_nerf.  Used by the sharper-field parity tests: the oracle is evaluated on the SAME weights, so
    the fit does not need to be reproducible across machines."""
    state = torch.get_rng_state()
    c, f, ea, et = my_nerfw.make_synthetic_nerf(D=D, W=W, gain=1.0, sigma_bias=0.0)
    c, f, ea, et = c.to(device), f.to(device), ea.to(device), et.to(device)
    g = torch.Generator(device="cpu").manual_seed(seed)
    opt = torch.optim.Adam(list(c.parameters()) + list(f.parameters()), lr=1e-3)
    hist = torch.tensor([5, 10, 20, 30, 15, 10, 5, 3, 1, 1], device=device)

    def scene(p):
        s1 = ((p - torch.tensor([0.15, 0.1, -0.4], device=device)).norm(dim=-1) < 0.35)
        s2 = ((p - torch.tensor([-0.35, -0.15, -0.9], device=device)).abs().amax(-1) < 0.25)
        wall = p[:, 2] < -1.2
        inside = (s1 | s2 | wall).float()
        sigma = 40.0 * inside
        stripes = 0.5 + 0.5 * torch.sin(18.0 * p[:, :1] + 11.0 * p[:, 1:2])
        rgb = torch.cat([stripes, 0.5 + 0.4 * torch.sin(7.0 * p[:, 1:2]), 0.3 + 0.6 * s1.float()[:, None]], 1)
        return sigma, rgb

    with torch.enable_grad():
        for it in range(steps):
            p = (torch.rand(batch, 3, generator=g) * torch.tensor([1.6, 1.2, 2.5]) + torch.tensor([-0.8, -0.6, -1.5])).to(device)
            d = torch.nn.functional.normalize(torch.randn(batch, 3, generator=g), dim=-1).to(device)
            sig, rgb = scene(p)
            a = ea(hist).reshape(1, -1).expand(batch, -1)
            t = et(hist).reshape(1, -1).expand(batch, -1)
            x = torch.cat([_embed_t(p, 10), _embed_t(d, 4), a, t], -1)
            out_f = torch_nerfw_forward(f, x, "full")
            out_c = torch_nerfw_forward(c, x[:, :63], "sigma")
            w = 1.0 / 40.0
            loss = (((out_f[:, 3] - sig) * w) ** 2).mean() + (((out_c[:, 0] - sig) * w) ** 2).mean() + \
                ((out_f[:, :3] - rgb) ** 2 * (0.05 + (sig > 0).float()[:, None])).mean() + \
                (out_f[:, 7] ** 2).mean() * 0.1 + ((out_f[:, 8] - 0.05) ** 2).mean() * 0.1
            opt.zero_grad()
            loss.backward()
            opt.step()
    for m in (c, f, ea, et):
        for q in m.parameters():
            q.requires_grad_(False)
    torch.set_rng_state(state)
    return c, f, ea, et


def full_size_pair(H=480, W=640, seed=6):
    """The 640x480 target / render pair of BASELINE config[2] (tests/golden/make_golden_dfnet_full.py): a smooth
    low-frequency image plus noise, and a slightly shifted / re-lit copy of it (a render resembles its target)."""
    rng = np.random.RandomState(seed)
    yy, xx = np.meshgrid(np.linspace(0, 1, H, dtype=np.float32), np.linspace(0, 1, W, dtype=np.float32), indexing="ij")
    base = np.stack([0.5 + 0.4 * np.sin(9 * xx + 3 * yy), 0.5 + 0.4 * np.cos(7 * yy - 2 * xx), 0.3 + 0.5 * xx * yy]).astype(np.float32)
    a = np.clip(base + 0.08 * rng.randn(3, H, W).astype(np.float32), 0, 1)
    b = np.clip(0.95 * np.roll(base, (3, -2), (1, 2)) + 0.02 + 0.08 * rng.randn(3, H, W).astype(np.float32), 0, 1)
    return np.stack([a, b]).astype(np.float32)


def nerf_train_case(case):
    """Inputs of the NeRF-Hist training-step golden (tests/golden/make_golden_nerf_train.py and the GPU test): seeded
    networks of the reference's default width (128) and of the benchmark width (256), a bundle of rays looking at the
    synthetic field, random target colours."""
    W = 128 if case == "w128" else 256
    state = torch.get_rng_state()
    mods = my_nerfw.make_synthetic_nerf(D=8, W=W)
    torch.set_rng_state(state)
    rng = np.random.RandomState(31 + W)
    n = 96
    o = np.tile(np.array([[0.0, 0.0, 1.0]], np.float32), (n, 1)) + 0.05 * rng.randn(n, 3).astype(np.float32)
    d = (rng.randn(n, 3) * 0.25 + np.array([0, 0, -1.0])).astype(np.float32)
    return dict(mods=mods, rays=np.stack([o, d]).astype(np.float32), target=rng.rand(n, 3).astype(np.float32),
                hist=np.array([[5, 10, 20, 30, 15, 10, 5, 3, 1, 1]], np.float32), Nc=16, Nf=24, near=0.0, far=2.5)
