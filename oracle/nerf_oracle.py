"""CPU oracle for the NeRF-Hist render path (TEST INFRASTRUCTURE, not product code).

A numpy restatement of the reference's render hot path, used ONLY as the checker
in tests/, in __graft_entry__.smoke() and as bench.py's cpu_baseline / reference arm.
Nothing under dfnet_b200/ may import this module.

Parity status: the reference ships no tests or golden vectors, so the
oracle is pinned against outputs of the reference itself, generated in the build
container by tests/golden/make_golden.py (imports /root/reference) and committed as
tests/golden/*.npz.  tests/test_oracle_golden.py checks every function below against
those vectors.

Each function cites the reference file:line it follows (paths relative to
/root/reference/script).  Arithmetic is float32 with one rounding per operation, like
eager ATen-CPU; where ATen-CPU accumulates differently (float64 scans, vectorised
float32 sums, FMA linspace) the same order is reproduced so that sample indices are
bit-exact against the reference.
"""
import numpy as np

f32 = np.float32


# --------------------------------------------------------------------------------------
# ATen-CPU arithmetic that index parity depends on
# --------------------------------------------------------------------------------------
def linspace_f32(start, end, steps):
    """torch.linspace(start, end, steps) on CPU float32 (rendering.py:32,269).

    ATen computes step=(end-start)/(steps-1) in float32 and evaluates the lower half as
    fma(step, i, start) and the upper half as fma(-step, steps-1-i, end); pinned
    empirically against torch 2.11 (tests/test_oracle_golden.py::test_linspace).
    """
    start = f32(start)
    end = f32(end)
    if steps == 1:
        return np.array([start], f32)
    step = f32((end - start) / f32(steps - 1))
    i = np.arange(steps)
    lo = (np.float64(step) * i.astype(np.float64) + np.float64(start)).astype(f32)
    hi = (np.float64(-step) * (steps - 1 - i).astype(np.float64) + np.float64(end)).astype(f32)
    return np.where(i < steps // 2, lo, hi).astype(f32)


def _ceil_log2(x):
    return 0 if x <= 1 else int(np.ceil(np.log2(x)))


def aten_sum_lastdim(x):
    """torch.sum(x, -1) for a contiguous float32 [rows, n] tensor on CPU
    (used by rendering.py:27 `torch.sum(weights, -1, keepdim=True)`).

    ATen's cascade_sum reduces each row with 8-lane vectors, 4 interleaved vector
    accumulators (ilp), a 4-level cascade every 16 steps, then adds the scalar tail and
    finally the 8 lanes sequentially.  The order is pinned empirically (it decides
    whether cdf[-1] rounds above 1.0, i.e. the last sample index).
    """
    x = np.ascontiguousarray(x, dtype=f32)
    rows, n = x.shape
    V, ILP, LEVELS = 8, 4, 4
    vec_size = n // V
    vecs = x[:, : vec_size * V].reshape(rows, vec_size, V)
    size_ilp = vec_size // ILP
    acc = np.zeros((LEVELS, rows, ILP, V), f32)
    if size_ilp > 0:
        blocks = vecs[:, : size_ilp * ILP].reshape(rows, size_ilp, ILP, V)
        level_power = max(4, _ceil_log2(size_ilp) // LEVELS)
        level_step = 1 << level_power
        level_mask = level_step - 1
        i = 0
        while i + level_step <= size_ilp:
            for _ in range(level_step):
                acc[0] = acc[0] + blocks[:, i]
                i += 1
            for j in range(1, LEVELS):
                acc[j] = acc[j] + acc[j - 1]
                acc[j - 1] = 0
                if (i & (level_mask << (j * level_power))) != 0:
                    break
        while i < size_ilp:
            acc[0] = acc[0] + blocks[:, i]
            i += 1
        for j in range(1, LEVELS):
            acc[0] = acc[0] + acc[j]
    ps = acc[0]  # [rows, ILP, V]
    p0 = ps[:, 0]
    for i in range(size_ilp * ILP, vec_size):
        p0 = p0 + vecs[:, i]
    for k in range(1, ILP):
        p0 = p0 + ps[:, k]
    fin = np.zeros(rows, f32)
    for k in range(vec_size * V, n):
        fin = fin + x[:, k]
    for k in range(V):
        fin = fin + p0[:, k]
    return fin.astype(f32)


def cumsum_f64acc(x):
    """torch.cumsum on CPU float32: accumulates in float64, rounds each element
    (rendering.py:28)."""
    return np.cumsum(x.astype(np.float64), -1).astype(f32)


def cumprod_f64acc(x):
    """torch.cumprod on CPU float32: float64 running product (rendering.py:178)."""
    return np.cumprod(x.astype(np.float64), -1).astype(f32)


# --------------------------------------------------------------------------------------
# a1/a2: rays
# --------------------------------------------------------------------------------------
def get_rays(H, W, focal, c2w):
    """models/ray_utils.py:5-15.  No half-pixel offset; rays_d[k] = (dx*R[k,0] +
    dy*R[k,1]) + dz*R[k,2] with every product rounded (materialised product then a
    3-term sequential sum)."""
    c2w = np.asarray(c2w, f32)
    i = np.broadcast_to(np.arange(W, dtype=f32)[None, :], (H, W))
    j = np.broadcast_to(np.arange(H, dtype=f32)[:, None], (H, W))
    dx = (i - f32(W * 0.5)) / f32(focal)
    dy = -(j - f32(H * 0.5)) / f32(focal)
    dz = -np.ones_like(dx)
    R = c2w[:3, :3]
    d = np.empty((H, W, 3), f32)
    for k in range(3):
        d[..., k] = (dx * R[k, 0] + dy * R[k, 1]) + dz * R[k, 2]
    o = np.broadcast_to(c2w[:3, 3], d.shape).copy()
    return o, d


def make_ray_records(rays_o, rays_d, near, far, hist):
    """rendering.py:366-389 with use_viewdirs=True, ndc=False: [o3,d3,near,far,vd3,hist]."""
    o = rays_o.reshape(-1, 3).astype(f32)
    d = rays_d.reshape(-1, 3).astype(f32)
    nrm = np.sqrt((d * d).sum(-1, keepdims=True, dtype=f32)).astype(f32)
    vd = (d / nrm).astype(f32)
    n = o.shape[0]
    hist = np.asarray(hist, f32).reshape(-1, np.asarray(hist).shape[-1])
    if hist.shape[0] != n:
        hist = np.broadcast_to(hist[:1], (n, hist.shape[1]))
    nf = np.ones((n, 1), f32)
    return np.concatenate([o, d, f32(near) * nf, f32(far) * nf, vd, hist], -1).astype(f32)


# --------------------------------------------------------------------------------------
# a6: positional encoding
# --------------------------------------------------------------------------------------
def embed(x, L):
    """models/nerfw.py:105-133 with get_embedder defaults (:198-207): include_input,
    bands 2**linspace(0,L-1,L) = 1,2,4..., sin before cos per band."""
    x = np.asarray(x, f32)
    out = [x]
    for l in range(L):
        fr = f32(2.0 ** l)
        xf = (x * fr).astype(f32)
        out.append(np.sin(xf).astype(f32))
        out.append(np.cos(xf).astype(f32))
    return np.concatenate(out, -1)


# --------------------------------------------------------------------------------------
# a7: NeRFW MLP
# --------------------------------------------------------------------------------------
_LINEAR_BACKEND = "numpy"


def set_linear_backend(name):
    """"numpy" (default, used by every parity test) or "torch": run the Linear layers through
    torch-CPU addmm (MKL/oneDNN), the library the reference itself executes them with.  Only
    bench.py's CPU-baseline timing selects "torch", so that the baseline is not handicapped by
    numpy's BLAS; all other arithmetic is unchanged."""
    global _LINEAR_BACKEND
    assert name in ("numpy", "torch")
    _LINEAR_BACKEND = name


def _linear(x, w, b):
    if _LINEAR_BACKEND == "torch":
        import torch
        return torch.addmm(torch.from_numpy(b), torch.from_numpy(np.ascontiguousarray(x)),
                           torch.from_numpy(w).t()).numpy()
    return (x @ w.T + b).astype(f32)


def _relu(x):
    if _LINEAR_BACKEND == "torch":
        import torch
        torch.relu_(torch.from_numpy(x))  # in place, multi-threaded, as nn.ReLU(True) in the reference
        return x
    return np.maximum(x, f32(0), out=x if x.flags.writeable and x.dtype == f32 else None)


def _softplus(x):
    # torch.nn.Softplus(beta=1, threshold=20)
    with np.errstate(over="ignore"):
        sp = np.log1p(np.exp(x.astype(f32))).astype(f32)
    return np.where(x > f32(20), x, sp).astype(f32)


def _sigmoid(x):
    with np.errstate(over="ignore"):
        return (f32(1) / (f32(1) + np.exp(-x.astype(f32)))).astype(f32)


def nerfw_forward(P, x, D, skips=(4,), sigma_only=False, output_transient=True,
                  in_xyz=63, in_dir=27, in_a=0, in_t=20):
    """models/nerfw.py:297-354.  P: dict name->ndarray with the reference state_dict
    names.  Output order [rgb3, sigma, t_rgb3, t_sigma, t_beta] (:340,351-354)."""
    x = np.asarray(x, f32)
    if sigma_only:
        input_xyz = x
    elif output_transient:
        input_xyz = x[:, :in_xyz]
        input_dir_a = x[:, in_xyz:in_xyz + in_dir + in_a]
        input_t = x[:, in_xyz + in_dir + in_a:in_xyz + in_dir + in_a + in_t]
    else:
        input_xyz = x[:, :in_xyz]
        input_dir_a = x[:, in_xyz:in_xyz + in_dir + in_a]
    h = input_xyz
    for i in range(D):
        if i in skips:
            h = np.concatenate([input_xyz, h], 1)
        h = _relu(_linear(h, P[f"xyz_encoding_{i+1}.0.weight"], P[f"xyz_encoding_{i+1}.0.bias"]))
    sigma = _softplus(_linear(h, P["static_sigma.0.weight"], P["static_sigma.0.bias"]))
    if sigma_only:
        return sigma
    final = _linear(h, P["xyz_encoding_final.weight"], P["xyz_encoding_final.bias"])
    de = _relu(_linear(np.concatenate([final, input_dir_a], 1),
                       P["dir_encoding.0.weight"], P["dir_encoding.0.bias"]))
    rgb = _sigmoid(_linear(de, P["static_rgb.0.weight"], P["static_rgb.0.bias"]))
    static = np.concatenate([rgb, sigma], 1)
    if not output_transient:
        return static
    t = np.concatenate([final, input_t], 1)
    for k in (0, 2, 4, 6):
        t = _relu(_linear(t, P[f"transient_encoding.{k}.weight"], P[f"transient_encoding.{k}.bias"]))
    t_sigma = _softplus(_linear(t, P["transient_sigma.0.weight"], P["transient_sigma.0.bias"]))
    t_rgb = _sigmoid(_linear(t, P["transient_rgb.0.weight"], P["transient_rgb.0.bias"]))
    t_beta = _softplus(_linear(t, P["transient_beta.0.weight"], P["transient_beta.0.bias"]))
    return np.concatenate([static, t_rgb, t_sigma, t_beta], 1).astype(f32)


def run_network(P, pts, viewdirs, hist, emb_a, emb_t, typ, test_time, D, skips=(4,),
                L_xyz=10, L_dir=4, netchunk=65536):
    """models/nerfw.py:15-95 (three modes).  pts [N,S,3], viewdirs [N,3], hist [N,hb]
    float percentages truncated like `.long()` (:69-72) and used as embedding rows."""
    N, S, _ = pts.shape
    flat = pts.reshape(-1, 3)
    outs = []
    if typ == "coarse" and test_time:
        for i in range(0, flat.shape[0], netchunk):
            outs.append(nerfw_forward(P, embed(flat[i:i + netchunk], L_xyz), D, skips, sigma_only=True))
        return np.concatenate(outs, 0).reshape(N, S, -1)
    dirs = np.broadcast_to(viewdirs[:, None, :], pts.shape).reshape(-1, 3)
    if typ == "coarse":
        for i in range(0, flat.shape[0], netchunk):
            e = np.concatenate([embed(flat[i:i + netchunk], L_xyz), embed(dirs[i:i + netchunk], L_dir)], 1)
            outs.append(nerfw_forward(P, e, D, skips, output_transient=False, in_a=0))
        return np.concatenate(outs, 0).reshape(N, S, -1)
    idx = np.asarray(hist).astype(np.int64)  # .long() truncation
    a = emb_a[idx].reshape(N, -1).astype(f32)
    t = emb_t[idx].reshape(N, -1).astype(f32)
    a_ = np.repeat(a, S, 0)
    t_ = np.repeat(t, S, 0)
    for i in range(0, flat.shape[0], netchunk):
        e = np.concatenate([embed(flat[i:i + netchunk], L_xyz), embed(dirs[i:i + netchunk], L_dir),
                            a_[i:i + netchunk], t_[i:i + netchunk]], 1)
        outs.append(nerfw_forward(P, e, D, skips, output_transient=True,
                                  in_a=a.shape[1], in_t=t.shape[1]))
    return np.concatenate(outs, 0).reshape(N, S, -1)


# --------------------------------------------------------------------------------------
# a8: compositing
# --------------------------------------------------------------------------------------
def raw2outputs_nerfw(raw, z_vals, raw_noise_std=0.0, output_transient=False, beta_min=0.1,
                      white_bkgd=False, test_time=False, static_only=True, typ="coarse", noise=None):
    """rendering.py:132-243.  Returns dict(rgb, disp, acc, weights, depth,
    transient_sigmas, beta).  Last delta is 1e2, no ||d|| scaling, transmittance is a
    float64-accumulated exclusive cumprod without epsilon.  `noise` replaces
    randn_like(sigma) (the reference draws it even when std == 0)."""
    raw = np.asarray(raw, f32)
    z = np.asarray(z_vals, f32)
    if typ == "coarse" and test_time:
        s_sig = raw[..., 0]
        t_sig = None
    else:
        ch = raw.shape[-1]
        ch_rgbs = (ch - 3) // 2 if output_transient else ch - 1
        s_rgb = raw[..., :ch_rgbs]
        s_sig = raw[..., ch_rgbs]
        if output_transient:
            t_rgb = raw[..., ch_rgbs + 1:2 * ch_rgbs + 1]
            t_sig = raw[..., 2 * ch_rgbs + 1]
            t_beta = raw[..., 2 * ch_rgbs + 2]
        else:
            t_sig = None
    deltas = np.concatenate([z[:, 1:] - z[:, :-1], np.full_like(z[:, :1], f32(1e2))], -1).astype(f32)
    one = f32(1)
    with np.errstate(over="ignore"):
        if output_transient:
            s_alpha = one - np.exp(-deltas * s_sig)
            t_alpha = one - np.exp(-deltas * t_sig)
            alphas = one - np.exp(-deltas * (s_sig + t_sig))
        else:
            nz = np.zeros_like(s_sig) if noise is None else np.asarray(noise, f32)
            alphas = one - np.exp(-deltas * np.maximum(s_sig + nz * f32(raw_noise_std), f32(0)))
    alphas = alphas.astype(f32)
    shifted = np.concatenate([np.ones_like(alphas[:, :1]), one - alphas], -1)
    trans = cumprod_f64acc(shifted[:, :-1])
    weights = (alphas * trans).astype(f32)
    wsum = weights.sum(-1, dtype=f32)
    out = dict(rgb=None, disp=None, acc=wsum, weights=weights, depth=None, transient_sigmas=t_sig, beta=None)
    if typ == "coarse" and test_time:
        return out
    if output_transient:
        s_w = (s_alpha.astype(f32) * trans).astype(f32)
        t_w = (t_alpha.astype(f32) * trans).astype(f32)
        s_map = (s_w[..., None] * s_rgb).sum(1, dtype=f32)
        if white_bkgd:
            s_map = s_map + (one - wsum[:, None])
        t_map = (t_w[..., None] * t_rgb).sum(1, dtype=f32)
        beta = (t_w * t_beta).sum(1, dtype=f32) + f32(beta_min)
        out["beta"] = beta.astype(f32)
        out["rgb"] = (s_map + t_map).astype(f32)
        if test_time and static_only:
            # :214-230 — rgb stays the static+transient composite; depth/disp use the
            # static-only transmittance.
            s_shift = np.concatenate([np.ones_like(alphas[:, :1]), one - s_alpha.astype(f32)], -1)
            s_trans = cumprod_f64acc(s_shift[:, :-1])
            s_w_ = (s_alpha.astype(f32) * s_trans).astype(f32)
            depth = (s_w_ * z).sum(-1, dtype=f32)
            out["depth"] = depth
            out["disp"] = (one / np.maximum(f32(1e-10), depth / wsum)).astype(f32)
            return out
    else:
        rgb = (weights[..., None] * s_rgb).sum(1, dtype=f32)
        if white_bkgd:
            rgb = rgb + (one - wsum[:, None])
        out["rgb"] = rgb.astype(f32)
        out["beta"] = np.zeros_like(wsum)
    depth = (weights * z).sum(-1, dtype=f32)
    out["depth"] = depth
    with np.errstate(divide="ignore", invalid="ignore"):
        out["disp"] = (one / np.maximum(f32(1e-10), depth / wsum)).astype(f32)
    return out


# --------------------------------------------------------------------------------------
# a9: hierarchical sampling
# --------------------------------------------------------------------------------------
def sample_pdf(bins, weights, n_samples, det=True, u=None):
    """rendering.py:24-65.  Returns (samples [N,n_samples], inds int64 [N,n_samples]).
    `u` supplies the random draws when det is False (the reference calls torch.rand)."""
    bins = np.asarray(bins, f32)
    w = (np.asarray(weights, f32) + f32(1e-5)).astype(f32)
    pdf = (w / aten_sum_lastdim(w)[:, None]).astype(f32)
    cdf = np.concatenate([np.zeros_like(pdf[:, :1]), cumsum_f64acc(pdf)], -1)
    N = cdf.shape[0]
    if det:
        u = np.broadcast_to(linspace_f32(0.0, 1.0, n_samples), (N, n_samples))
    else:
        u = np.asarray(u, f32)
    # searchsorted(right=True): number of cdf entries <= u
    inds = (cdf[:, None, :] <= u[:, :, None]).sum(-1).astype(np.int64)
    below = np.maximum(0, inds - 1)
    above = np.minimum(cdf.shape[-1] - 1, inds)
    c0 = np.take_along_axis(cdf, below, 1)
    c1 = np.take_along_axis(cdf, above, 1)
    b0 = np.take_along_axis(bins, below, 1)
    b1 = np.take_along_axis(bins, above, 1)
    denom = (c1 - c0).astype(f32)
    denom = np.where(denom < f32(1e-5), f32(1), denom)
    t = ((u - c0) / denom).astype(f32)
    samples = (b0 + (t * (b1 - b0)).astype(f32)).astype(f32)
    return samples, inds


# --------------------------------------------------------------------------------------
# a4: render_rays / a2-a3: render
# --------------------------------------------------------------------------------------
def render_rays(rays, nets, N_samples, N_importance, perturb=0.0, raw_noise_std=0.0,
                test_time=False, lindisp=False, t_rand=None, u=None, retraw=False,
                return_internals=False):
    """rendering.py:245-337.  rays [N, 11+hb]; nets = dict(coarse=P, fine=P, emb_a,
    emb_t, D, skips, beta_min).  white_bkgd is never forwarded correctly by the
    reference (positional-argument slip at :295) and is therefore not modelled."""
    rays = np.asarray(rays, f32)
    N = rays.shape[0]
    o, d = rays[:, 0:3], rays[:, 3:6]
    near, far = rays[:, 6:7], rays[:, 7:8]
    vd = rays[:, 8:11]
    hist = rays[:, 11:]
    t_vals = linspace_f32(0.0, 1.0, N_samples)[None, :]
    one = f32(1)
    if not lindisp:
        z = (near * (one - t_vals) + far * t_vals).astype(f32)
    else:
        z = (one / (one / near * (one - t_vals) + one / far * t_vals)).astype(f32)
    z = np.broadcast_to(z, (N, N_samples)).astype(f32)
    if perturb > 0.0:
        mids = (f32(0.5) * (z[:, 1:] + z[:, :-1])).astype(f32)
        upper = np.concatenate([mids, z[:, -1:]], -1)
        lower = np.concatenate([z[:, :1], mids], -1)
        z = (lower + (upper - lower) * np.asarray(t_rand, f32)).astype(f32)
    pts = (o[:, None, :] + (d[:, None, :] * z[:, :, None]).astype(f32)).astype(f32)
    D, skips = nets["D"], tuple(nets.get("skips", (4,)))
    raw = run_network(nets["coarse"], pts, vd, None, None, None, "coarse", test_time, D, skips)
    c = raw2outputs_nerfw(raw, z, raw_noise_std, False, test_time=test_time, typ="coarse")
    ret = {"rgb_map": c["rgb"], "disp_map": c["disp"], "acc_map": c["acc"]}
    internals = {"z_coarse": z, "weights_coarse": c["weights"]}
    if N_importance > 0:
        z_mid = (f32(0.5) * (z[:, 1:] + z[:, :-1])).astype(f32)
        z_samples, inds = sample_pdf(z_mid, c["weights"][:, 1:-1], N_importance, det=(perturb == 0.0), u=u)
        z_all = np.sort(np.concatenate([z, z_samples], -1), -1)
        pts = (o[:, None, :] + (d[:, None, :] * z_all[:, :, None]).astype(f32)).astype(f32)
        raw = run_network(nets["fine"], pts, vd, hist, nets["emb_a"], nets["emb_t"], "fine", test_time, D, skips)
        f = raw2outputs_nerfw(raw, z_all, raw_noise_std, True, nets.get("beta_min", 0.1),
                              test_time=test_time, typ="fine")
        ret = {"rgb_map": f["rgb"], "disp_map": f["disp"], "acc_map": f["acc"]}
        internals.update(z_samples=z_samples, inds=inds, z_vals=z_all, weights_fine=f["weights"])
        if not test_time:
            ret["rgb0"], ret["disp0"], ret["acc0"] = c["rgb"], c["disp"], c["acc"]
            ret["z_std"] = z_samples.std(-1).astype(f32)
            ret["transient_sigmas"] = f["transient_sigmas"]
            ret["beta"] = f["beta"]
    if retraw:
        ret["raw"] = raw
    if return_internals:
        ret["_internals"] = internals
    return ret


def render(H, W, focal, nets, N_samples, N_importance, near, far, c2w=None, rays=None,
           hist=None, chunk=32768, **kw):
    """rendering.py:353-400 with use_viewdirs=True, ndc=False."""
    if c2w is not None:
        o, d = get_rays(H, W, focal, c2w)
    else:
        o, d = rays
    sh = d.shape
    rec = make_ray_records(o, d, near, far, hist)
    outs = {}
    for i in range(0, rec.shape[0], chunk):
        r = render_rays(rec[i:i + chunk], nets, N_samples, N_importance, **kw)
        for k, v in r.items():
            if k == "_internals":
                for kk, vv in v.items():
                    outs.setdefault("_" + kk, []).append(vv)
            elif v is not None:
                outs.setdefault(k, []).append(v)
    res = {}
    for k, v in outs.items():
        a = np.concatenate(v, 0)
        res[k] = a.reshape(list(sh[:-1]) + list(a.shape[1:]))
    return res
