"""Tensor-level wrappers over the C ABI: torch supplies device memory and streams only."""
import ctypes as C
import weakref

import torch

from . import _lib
from ._lib import lib, check, raw_stream


def _require_cuda(t, name):
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise _lib.DfbError(f"{name} must be a CUDA tensor: the dfnet_b200 hot path has no CPU fallback")


def _f32c(t):
    return t.detach().to(torch.float32).contiguous()


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream():
    return raw_stream()


def linspace(start, end, steps):
    """torch.linspace as ATen-CPU evaluates it, from the library's host helper."""
    buf = (C.c_float * steps)()
    check(lib.dfb_linspace_f32(start, end, steps, buf))
    return torch.tensor(list(buf), dtype=torch.float32)


class NerfHandle:
    """Owns a DfbNerf*: the repacked coarse/fine NeRFW weights and the histogram embeddings."""

    def __init__(self, network_fn, network_fine=None, embedding_a=None, embedding_t=None):
        if not torch.cuda.is_available():
            raise _lib.DfbError("no CUDA device: the dfnet_b200 hot path has no CPU fallback")
        net = network_fn
        a_dim = network_fine.in_channels_a if network_fine is not None else 50
        t_dim = network_fine.in_channels_t if network_fine is not None else 20
        hist_bin = a_dim // 5
        n_vocab = embedding_a.weight.shape[0] if embedding_a is not None else 1
        skips = list(net.skips)
        d = _lib.NerfDesc(D=net.D, W=net.W, skip=skips[0] if skips else -1, L_xyz=(net.in_channels_xyz - 3) // 6,
                          L_dir=(net.in_channels_dir - 3) // 6, a_dim=a_dim, t_dim=t_dim, hist_bin=hist_bin,
                          n_vocab=n_vocab, beta_min=float(getattr(network_fine or net, "beta_min", 0.1)),
                          has_fine=1 if network_fine is not None else 0)
        self.desc = d
        h = C.c_void_p()
        check(lib.dfb_nerf_create(C.byref(d), C.byref(h)))
        self._h = h
        self._fin = weakref.finalize(self, lib.dfb_nerf_destroy, h)
        self._mods = (network_fn, network_fine, embedding_a, embedding_t)
        # shapes the tcgen05 kernels cover (forward with saved ReLU masks + mask-driven backward)
        # (narrower networks run on the same kernels, embedded in 8x256 with zero weights)
        self.tc_train = bool(network_fine is not None and 32 <= net.W <= 256 and net.W % 16 == 0 and net.D == 8 and
                             skips == [4] and net.in_channels_xyz == 63 and net.in_channels_dir == 27 and
                             network_fine.W == net.W and network_fine.D == 8)
        self._versions = None
        self._ws = None
        self.refresh(force=True)

    def __reduce__(self):
        # handles are caches of device state: pickling / deep-copying a module that carries one (torch.save(model),
        # copy.deepcopy) drops it, and the copy builds its own on first use
        return (_no_handle, ())

    def invalidate(self):
        """Force a re-upload on the next render.  Change detection compares (data_ptr, _version) of every tensor;
        writes through `.data` (p.data.copy_(...), common in older training code) do not bump `_version`, so call this
        (or refresh(force=True)) after such writes."""
        self._versions = None

    def _param_versions(self):
        # polled on every render: iterate the tensors directly (state_dict() renames and detaches every one of them)
        v = []
        for m in self._mods:
            if m is not None:
                v += [(p.data_ptr(), p._version) for p in m.parameters()]
                v += [(p.data_ptr(), p._version) for p in m.buffers()]
        return v

    def refresh(self, force=False):
        """Re-upload parameters if any module tensor changed since the last call."""
        v = self._param_versions()
        if not force and v == self._versions:
            return
        fn, fine, ea, et = self._mods
        for which, m in ((0, fn), (1, fine)):
            if m is None:
                continue
            ts = [_f32c(t) for t in m.state_dict().values()]
            ptrs = (C.c_void_p * len(ts))(*[t.data_ptr() for t in ts])
            numel = (C.c_int64 * len(ts))(*[t.numel() for t in ts])
            check(lib.dfb_nerf_load(self._h, which, ptrs, numel, len(ts)))
        if ea is not None and et is not None:
            wa, wt = _f32c(ea.weight), _f32c(et.weight)
            check(lib.dfb_nerf_set_embeddings(self._h, _ptr(wa), _ptr(wt)))
        self._versions = v

    def workspace(self, cfg, n_rays, device, extra_bytes=0):
        need = C.c_size_t()
        check(lib.dfb_render_workspace_bytes(self._h, C.byref(cfg), n_rays, C.byref(need)))
        total = need.value + extra_bytes
        if self._ws is None or self._ws.numel() < total or self._ws.device != device:
            self._ws = torch.empty(total, dtype=torch.uint8, device=device)
        return self._ws, need.value

    # ------------------------------------------------------------------------------
    def render(self, N_samples, N_importance, test_time, rays=None, c2w=None, H=0, W=0, focal=1.0, near=0.0,
               far=1.0, hist=None, perturb=False, t_rand=None, u=None, mma="f16", lindisp=False, raw_noise_std=0.0,
               noise=None, ert_eps=0.0, want=()):
        """dfb_render_fwd.  Returns dict(rgb, disp, acc, + requested extras)."""
        cfg = _lib.RenderCfg(N_samples=N_samples, N_importance=N_importance, test_time=int(bool(test_time)),
                             perturb=int(bool(perturb)), mma_kind=_lib.MMA_KINDS[mma], lindisp=int(bool(lindisp)),
                             raw_noise_std=float(raw_noise_std), ert_eps=float(ert_eps))
        hb = self.desc.hist_bin
        if rays is not None:
            _require_cuda(rays, "rays")
            if rays.dim() != 2 or rays.shape[1] != 11 + hb:
                raise _lib.DfbError(f"ray records must be [N, 11 + hist_bin = {11 + hb}] ([o3,d3,near,far,viewdir3,hist], reference "
                                    f"rendering.py:382-389), got {tuple(rays.shape)}")
            rays = _f32c(rays)
            N, dev = rays.shape[0], rays.device
            cfg.ray_stride = rays.shape[1]
        else:
            _require_cuda(c2w, "c2w")
            c2w = _f32c(c2w[:3, :4])
            if hist is None or hist.numel() != hb:
                raise _lib.DfbError(f"img_idx / hist must hold hist_bin = {hb} values, got "
                                    f"{0 if hist is None else hist.numel()}")
            hist = _f32c(hist.reshape(-1)).to(c2w.device)
            N, dev = H * W, c2w.device
            cfg.hist_len = hist.numel()
        S = N_samples + N_importance
        out = {"rgb": torch.empty(N, 3, device=dev), "disp": torch.empty(N, device=dev),
               "acc": torch.empty(N, device=dev)}
        shapes = {"rgb0": (N, 3), "disp0": (N,), "acc0": (N,), "z_std": (N,), "beta": (N,),
                  "transient_sigmas": (N, S), "raw": (N, S, 9) if N_importance > 0 else (N, N_samples, 4),
                  "weights_coarse": (N, N_samples), "z_vals": (N, S), "z_samples": (N, max(N_importance, 1)),
                  "inds": (N, max(N_importance, 1)), "depth": (N,), "relu_masks": ((N * S + 127) // 128, 12, 8, 128),
                  "n_live": (N,)}
        ex = _lib.RenderExtras()
        for k in want:
            out[k] = torch.empty(shapes[k], device=dev, dtype=torch.int32 if k in ("inds", "relu_masks", "n_live") else torch.float32)
            setattr(ex, k, out[k].data_ptr())
        ws, ws_bytes = self.workspace(cfg, N, dev)
        if t_rand is not None:
            t_rand = _f32c(t_rand).to(dev)
        if u is not None:
            u = _f32c(u).to(dev)
        if noise is not None:
            noise = _f32c(noise).to(dev)
        check(lib.dfb_render_fwd(self._h, C.byref(cfg), _ptr(rays), _ptr(c2w), H, W, focal, near, far, _ptr(hist), N,
                                 _ptr(t_rand), _ptr(u), _ptr(noise), _ptr(out["rgb"]), _ptr(out["disp"]), _ptr(out["acc"]),
                                 C.byref(ex), _ptr(ws), ws_bytes, _stream()))
        return out

    def render_poses(self, N_samples, N_importance, c2w, hist, H, W, focal, near, far, mma="f16", lindisp=False):
        """dfb_render_poses_fwd: test-time render of n poses [n,3,4] (histograms [n,hist_bin]) in one call
        -> dict(rgb [n,H,W,3], disp [n,H,W], acc [n,H,W])."""
        _require_cuda(c2w, "c2w")
        c2w = _f32c(c2w[:, :3, :4])
        n, dev = c2w.shape[0], c2w.device
        hist = _f32c(hist.reshape(n, -1)).to(dev)
        if hist.shape[1] != self.desc.hist_bin:
            raise _lib.DfbError(f"hist must be [n, hist_bin = {self.desc.hist_bin}], got {tuple(hist.shape)}")
        cfg = _lib.RenderCfg(N_samples=N_samples, N_importance=N_importance, test_time=1, perturb=0, mma_kind=_lib.MMA_KINDS[mma],
                             lindisp=int(bool(lindisp)), raw_noise_std=0.0, hist_len=hist.shape[1])
        N = n * H * W
        out = {"rgb": torch.empty(n, H, W, 3, device=dev), "disp": torch.empty(n, H, W, device=dev), "acc": torch.empty(n, H, W, device=dev)}
        ws, ws_bytes = self.workspace(cfg, N, dev)
        check(lib.dfb_render_poses_fwd(self._h, C.byref(cfg), _ptr(c2w), n, H, W, float(focal), float(near), float(far), _ptr(hist),
                                       _ptr(out["rgb"]), _ptr(out["disp"]), _ptr(out["acc"]), _ptr(ws), ws_bytes, _stream()))
        return out

    def render_image_host(self, cfg, c2w_host, H, W, focal, near, far, hist_host, rgb_host, disp_host, acc_host,
                          device):
        """dfb_render_image_host: pinned host pose/hist in, pinned host image out, all on the current stream."""
        stage = 256 * 2 + (H * W * 5 * 4 + 255) // 256 * 256
        if hist_host.numel() != self.desc.hist_bin:
            raise _lib.DfbError(f"hist must hold hist_bin = {self.desc.hist_bin} values, got {hist_host.numel()}")
        cfg.hist_len = hist_host.numel()
        ws, ws_bytes = self.workspace(cfg, H * W, device, extra_bytes=stage)
        check(lib.dfb_render_image_host(self._h, C.byref(cfg), _ptr(c2w_host), H, W, focal, near, far, _ptr(hist_host),
                                        _ptr(rgb_host), _ptr(disp_host), _ptr(acc_host), _ptr(ws), ws.numel(),
                                        _stream()))

    def render_backward(self, rays, z_vals, raw, g_rgb, mma="fp32", relu_masks=None):
        """dfb_render_bwd_mma: gradients of a test-time render w.r.t. rays_o, rays_d and viewdirs, each [N,3].
        mma = the kind the forward ran with; "f16"/"bf16" run the 8x256 fine network's backward on tcgen05."""
        _require_cuda(rays, "rays")
        if rays.dim() != 2 or rays.shape[1] != 11 + self.desc.hist_bin:
            raise _lib.DfbError(f"ray records must be [N, 11 + hist_bin = {11 + self.desc.hist_bin}], got {tuple(rays.shape)}")
        rays, z_vals, raw, g_rgb = _f32c(rays), _f32c(z_vals), _f32c(raw), _f32c(g_rgb)
        N, S = z_vals.shape
        dev = rays.device
        need = C.c_size_t()
        check(lib.dfb_render_bwd_workspace_bytes(self._h, N, S, C.byref(need)))
        ws = torch.empty(need.value, dtype=torch.uint8, device=dev)
        g_o, g_d, g_vd = (torch.empty(N, 3, device=dev) for _ in range(3))
        # relu_masks: the forward's saved ReLU masks (render(want=("relu_masks",)), tcgen05 path) -> no forward recompute
        check(lib.dfb_render_bwd_saved(self._h, _lib.MMA_KINDS[mma], _ptr(rays), rays.shape[1], N, S, _ptr(z_vals), _ptr(raw), _ptr(relu_masks),
                                       _ptr(g_rgb), _ptr(g_o), _ptr(g_d), _ptr(g_vd), _ptr(ws), need.value, _stream()))
        return g_o, g_d, g_vd

    def nerfw_forward(self, which, mode, x):
        _require_cuda(x, "x")
        x = _f32c(x)
        C_out = {0: 1, 1: 4, 2: 9}[mode]
        out = torch.empty(x.shape[0], C_out, device=x.device)
        check(lib.dfb_nerfw_forward(self._h, which, mode, _ptr(x), x.shape[0], _ptr(out), _stream()))
        return out


def _no_handle():
    return None


def handle_for(network_fn, network_fine=None, embedding_a=None, embedding_t=None):
    """One handle per (coarse, fine, emb_a, emb_t) module set; weights re-uploaded when they change.
    The handle is cached ON the coarse module (a module -> handle -> module cycle the garbage collector can free), so
    it lives exactly as long as the networks do and its device memory is released with them."""
    cache = network_fn.__dict__.setdefault("_dfb_handles", {})
    key = tuple(id(m) for m in (network_fine, embedding_a, embedding_t))
    h = cache.get(key)
    if h is None or h._mods[1:] != (network_fine, embedding_a, embedding_t):
        h = NerfHandle(network_fn, network_fine, embedding_a, embedding_t)
        cache[key] = h
    else:
        h.refresh()
    return h


def nerfw_forward(module, x, sigma_only=False, output_transient=True):
    """NeRFW.forward seam (reference nerfw.py:297-354) through dfb_nerfw_forward."""
    fine = module.typ == "fine"
    h = module.__dict__.get("_dfb_single_handle")
    if h is None:
        if fine:
            # a fine net needs a coarse slot: a proxy with the same trunk sizes (never evaluated)
            from .nerfw import NeRFW
            with torch.random.fork_rng(devices=[]):   # NeRFW.__init__ reseeds the global RNG (nerfw.py:245): keep that
                coarse = NeRFW("coarse", D=module.D, W=module.W, skips=module.skips)  # side effect out of a forward call
            h = NerfHandle(coarse, module, None, None)
        else:
            h = NerfHandle(module, None, None, None)
        module.__dict__["_dfb_single_handle"] = h
    else:
        h.refresh()
    mode = 0 if sigma_only else (2 if (output_transient and fine) else 1)
    return h.nerfw_forward(1 if fine else 0, mode, x)


def sample_pdf(bins, weights, N_samples, det=False, u=None):
    """sample_pdf seam (reference rendering.py:24-65) -> (samples, inds int32)."""
    _require_cuda(bins, "bins")
    bins, weights = _f32c(bins), _f32c(weights)
    N, nb = bins.shape
    if not det and u is None:
        u = torch.rand(N, N_samples, device=bins.device)
    if u is not None:
        u = _f32c(u).to(bins.device)
    samples = torch.empty(N, N_samples, device=bins.device)
    inds = torch.empty(N, N_samples, device=bins.device, dtype=torch.int32)
    check(lib.dfb_sample_pdf(_ptr(bins), _ptr(weights), _ptr(None if det else u), N, nb, N_samples, _ptr(samples),
                             _ptr(inds), _stream()))
    return samples, inds


def raw2outputs(raw, z_vals, typ, test_time, beta_min=0.1, noise=None, raw_noise_std=0.0):
    """raw2outputs_NeRFW seam (reference rendering.py:132-243)."""
    _require_cuda(raw, "raw")
    raw, z = _f32c(raw), _f32c(z_vals)
    N, S, Cc = raw.shape
    dev = raw.device
    o = {k: torch.empty(N, device=dev) for k in ("disp", "acc", "depth", "beta")}
    o["rgb"] = torch.empty(N, 3, device=dev)
    o["weights"] = torch.empty(N, S, device=dev)
    o["transient_sigmas"] = torch.empty(N, S, device=dev) if Cc == 9 else None
    check(lib.dfb_raw2outputs(_ptr(raw), _ptr(z), N, S, Cc, 1 if typ == "fine" else 0, int(bool(test_time)),
                              beta_min, _ptr(o["rgb"]), _ptr(o["disp"]), _ptr(o["acc"]), _ptr(o["weights"]),
                              _ptr(o["depth"]), _ptr(o["transient_sigmas"]), _ptr(o["beta"]),
                              _ptr(_f32c(noise) if noise is not None else None), float(raw_noise_std), _stream()))
    return o


def get_rays(H, W, focal, c2w):
    """get_rays seam (reference ray_utils.py:5-15)."""
    _require_cuda(c2w, "c2w")
    c = _f32c(c2w)
    o = torch.empty(H * W, 3, device=c.device)
    d = torch.empty(H * W, 3, device=c.device)
    check(lib.dfb_get_rays(_ptr(c), c.stride(0), H, W, float(focal), _ptr(o), _ptr(d), _stream()))
    return o.view(H, W, 3), d.view(H, W, 3)
