"""Host mirror of the reference's models/rendering.py call surface.

Same names, arguments and return structure as the reference (render, batchify_rays,
render_rays, raw2outputs_NeRFW, sample_pdf); the arithmetic runs in libdfnet_b200 through
dfnet_b200.ops.  Options no shipped config reaches (ndc, c2w_staticcam, white_bkgd)
raise instead of silently diverging.
"""
import os
import struct
import time
import zlib
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

from . import _lib, ops
from .nerfw import to8b
from .ray_utils import get_rays  # noqa: F401  (re-exported like `from models.ray_utils import *`)

DEFAULT_MMA = "f16"


def sample_pdf(bins, weights, N_samples, det=False, pytest=False):
    """Hierarchical inverse-CDF sampling (reference rendering.py:24-65)."""
    u = None
    if pytest and not det:
        import numpy as np
        np.random.seed(0)
        u = torch.tensor(np.random.rand(bins.shape[0], N_samples), dtype=torch.float32)
    samples, _ = ops.sample_pdf(bins, weights, N_samples, det=det, u=u)
    return samples


def raw2outputs_NeRFW(raw, z_vals, rays_d=None, raw_noise_std=0, output_transient=False, beta_min=0.1,
                      white_bkgd=False, test_time=False, static_only=True, typ="coarse"):
    """Volumetric compositing (reference rendering.py:132-243) ->
    (rgb_map, disp_map, acc_map, weights, depth_map, transient_sigmas, beta)."""
    if white_bkgd:
        raise NotImplementedError("white_bkgd is not on the B200 hot path")
    if not static_only:
        raise NotImplementedError("static_only=False is never used by the reference")
    if typ == "coarse" and test_time and raw.shape[-1] != 1:
        raw = raw[..., :1]
    noise = None
    if not output_transient:
        # the reference draws randn_like(static_sigmas) on this branch even when raw_noise_std == 0 (:173)
        noise = torch.randn(raw.shape[:-1], device=raw.device)
    o = ops.raw2outputs(raw, z_vals, typ, test_time, beta_min, noise=noise if raw_noise_std else None,
                        raw_noise_std=raw_noise_std if not output_transient else 0.0)
    if typ == "coarse" and test_time:
        return None, None, o["acc"], o["weights"], None, None, None
    return o["rgb"], o["disp"], o["acc"], o["weights"], o["depth"], o["transient_sigmas"], o["beta"]


class _RenderRaysFn(torch.autograd.Function):
    """Test-time render of explicit rays with a hand-written backward w.r.t. rays_o, rays_d and viewdirs
    (dfb_render_bwd).  Only rgb_map carries gradient (what train.py back-propagates, reference
    feature/direct_feature_matching.py:342-378); disp_map and acc_map are marked non-differentiable."""

    @staticmethod
    def forward(ctx, rays_o, rays_d, viewdirs, near_far_hist, handle, N_samples, N_importance, mma):
        rec = torch.cat([rays_o, rays_d, near_far_hist[:, :2], viewdirs, near_far_hist[:, 2:]], -1)
        # tcgen05 path: the forward also saves the ReLU masks of the fine network (one bit per activation), so that
        # the backward kernel does not have to recompute the forward
        saved = mma in ("f16", "bf16", "f16s") and handle.tc_train
        o = handle.render(N_samples, N_importance, True, rays=rec, mma=mma,
                          want=("z_vals", "raw", "relu_masks") if saved else ("z_vals", "raw"))
        ctx.handle, ctx.mma, ctx.saved = handle, mma, saved
        ctx.save_for_backward(rec, o["z_vals"], o["raw"], *([o["relu_masks"]] if saved else []))
        ctx.mark_non_differentiable(o["disp"], o["acc"])
        return o["rgb"], o["disp"], o["acc"]

    @staticmethod
    def backward(ctx, g_rgb, g_disp, g_acc):
        rec, z_vals, raw = ctx.saved_tensors[:3]
        masks = ctx.saved_tensors[3] if ctx.saved else None
        g_o, g_d, g_vd = ctx.handle.render_backward(rec, z_vals, raw, g_rgb, mma=ctx.mma, relu_masks=masks)
        return g_o, g_d, g_vd, None, None, None, None, None


def _get_rays_torch(H, W, focal, c2w):
    """Differentiable get_rays (reference models/ray_utils.py:5-15) for the training path, where the
    pose carries gradient; a handful of tiny torch ops."""
    dev = c2w.device
    i, j = torch.meshgrid(torch.linspace(0, W - 1, W, device=dev), torch.linspace(0, H - 1, H, device=dev), indexing="ij")
    i, j = i.t(), j.t()
    dirs = torch.stack([(i - W * .5) / focal, -(j - H * .5) / focal, -torch.ones_like(i)], -1)
    rays_d = torch.sum(dirs[..., None, :] * c2w[:3, :3], -1)
    rays_o = c2w[:3, -1].expand(rays_d.shape)
    return rays_o, rays_d


class _PoseRaysFn(torch.autograd.Function):
    """(rays_o, rays_d, viewdirs) [H*W,3] of a pose that carries gradient, with the adjoint onto c2w[:3,:4]
    (dfb_pose_rays_fwd / _bwd): what `_get_rays_torch` + the view-direction normalisation do in ~20 launches forward and
    ~25 backward, in 1 + 2."""

    @staticmethod
    def forward(ctx, c2w, H, W, focal):
        c = c2w.detach().float()
        if c.stride(-1) != 1:
            c = c.contiguous()
        dev = c.device
        o, d, v = (torch.empty(H * W, 3, device=dev) for _ in range(3))
        ops.check(ops.lib.dfb_pose_rays_fwd(c.data_ptr(), c.stride(0), H, W, float(focal), o.data_ptr(), d.data_ptr(), v.data_ptr(),
                                            _lib.raw_stream()))
        ctx.save_for_backward(c)
        ctx.geom, ctx.in_shape = (H, W, float(focal)), tuple(c2w.shape)
        return o, d, v

    @staticmethod
    def backward(ctx, g_o, g_d, g_v):
        (c,) = ctx.saved_tensors
        H, W, focal = ctx.geom

        def p(t):
            return None if t is None else t.float().contiguous()
        g_o, g_d, g_v = p(g_o), p(g_d), p(g_v)
        ws = torch.empty(ops.lib.dfb_pose_rays_workspace_bytes() // 8, device=c.device, dtype=torch.float64)
        out = torch.zeros(ctx.in_shape, device=c.device) if ctx.in_shape != (3, 4) else torch.empty(3, 4, device=c.device)
        g12 = out if ctx.in_shape == (3, 4) else torch.empty(3, 4, device=c.device)
        ops.check(ops.lib.dfb_pose_rays_bwd(c.data_ptr(), c.stride(0), H, W, focal, None if g_o is None else g_o.data_ptr(),
                                            None if g_d is None else g_d.data_ptr(), None if g_v is None else g_v.data_ptr(),
                                            ws.data_ptr(), g12.data_ptr(), _lib.raw_stream()))
        if g12 is not out:
            out[:3, :4] = g12
        return out, None, None, None


def _handle(kw):
    return ops.handle_for(kw["network_fn"], kw.get("network_fine"), kw.get("embedding_a"), kw.get("embedding_t"))


def _check_kwargs(kw):
    if kw.get("white_bkgd"):
        raise NotImplementedError("white_bkgd is dropped by the reference itself (rendering.py:295) and unsupported here")


def render_rays(ray_batch, network_fn, network_query_fn=None, N_samples=64, retraw=False, lindisp=False, perturb=0.,
                N_importance=0, network_fine=None, white_bkgd=False, raw_noise_std=0., verbose=False, pytest=False,
                i_epoch=-1, embedding_a=None, embedding_t=None, test_time=False, mma=None, ert_eps=0.0):
    """Render a batch of ray records [N, 11+hist_bin] (reference rendering.py:245-337)."""
    kw = dict(network_fn=network_fn, network_fine=network_fine, embedding_a=embedding_a, embedding_t=embedding_t,
              white_bkgd=white_bkgd, raw_noise_std=raw_noise_std)
    _check_kwargs(kw)
    if not test_time and torch.is_grad_enabled() and any(
            p.requires_grad for m in (network_fn, network_fine, embedding_a, embedding_t) if m is not None for p in m.parameters()):
        # NeRF-Hist training (run_nerf.py:51): differentiable w.r.t. the networks and the histogram embeddings.  Other
        # network shapes render through the forward kernels below; their outputs carry no gradient to the parameters
        # (freshly constructed modules require grad by default, so this is also the path of a plain train-mode render).
        from . import nerf_train
        if all(nerf_train.trainable_shape(m) for m in (network_fn, network_fine) if m is not None):
            return nerf_train.render_rays_train(ray_batch, network_fn, network_fine, embedding_a, embedding_t, N_samples, N_importance,
                                                perturb=perturb, raw_noise_std=raw_noise_std, lindisp=lindisp, retraw=retraw,
                                                pytest=pytest)
    h = _handle(kw)
    N = ray_batch.shape[0]
    t_rand = u = noise = None
    if perturb > 0. or raw_noise_std > 0.:
        # same draws, shapes and order as the reference: t_rand (:282), then the coarse pass' randn_like(static_sigmas)
        # (:173 - drawn even when raw_noise_std == 0, it advances the generator between t_rand and u), then u (:36)
        if perturb > 0.:
            t_rand = torch.rand(N, N_samples, device=ray_batch.device)
        noise = torch.randn(N, N_samples, device=ray_batch.device)
        if N_importance > 0 and perturb > 0.:
            if pytest:
                import numpy as np
                np.random.seed(0)
                u = torch.tensor(np.random.rand(N, N_importance), dtype=torch.float32)
            else:
                u = torch.rand(N, N_importance, device=ray_batch.device)
    want = []
    if N_importance > 0 and not test_time:
        want += ["rgb0", "disp0", "acc0", "z_std", "transient_sigmas", "beta"]
    if retraw:
        want.append("raw")
    o = h.render(N_samples, N_importance, test_time, rays=ray_batch, perturb=perturb > 0., t_rand=t_rand, u=u,
                 mma=mma or DEFAULT_MMA, lindisp=lindisp, raw_noise_std=raw_noise_std,
                 noise=noise if raw_noise_std > 0. else None, ert_eps=ert_eps, want=want)
    ret = {"rgb_map": o["rgb"], "disp_map": o["disp"], "acc_map": o["acc"]}
    for k in want:
        ret[k] = o[k]
    return ret


def batchify_rays(rays_flat, chunk=1024 * 32, **kwargs):
    """Reference rendering.py:339-351.  The kernels bound their own workspace, so `chunk`
    only controls how the torch.rand draws are grouped when perturb > 0."""
    if kwargs.get("perturb", 0.) > 0. or kwargs.get("raw_noise_std", 0.) > 0. or (
            torch.is_grad_enabled() and not kwargs.get("test_time", False)):
        outs = {}
        for i in range(0, rays_flat.shape[0], chunk):
            r = render_rays(rays_flat[i:i + chunk], **kwargs)
            for k, v in r.items():
                outs.setdefault(k, []).append(v)
        return {k: torch.cat(v, 0) for k, v in outs.items()}
    return render_rays(rays_flat, **kwargs)


def render(H, W, focal, chunk=1024 * 32, rays=None, c2w=None, ndc=True, near=0., far=1., use_viewdirs=False,
           c2w_staticcam=None, img_idx=torch.Tensor(0), **kwargs):
    """Reference rendering.py:353-400 -> [rgb_map, disp_map, acc_map, extras]."""
    if ndc:
        raise NotImplementedError("ndc=True: create_nerf forces ndc=False for every shipped dataset (nerfw.py:492-495)")
    if not use_viewdirs or c2w_staticcam is not None:
        raise NotImplementedError("NeRF-Hist always renders with use_viewdirs=True and no static camera")
    kwargs.pop("network_query_fn", None)
    _check_kwargs(kwargs)
    mma = kwargs.pop("mma", None) or DEFAULT_MMA
    ert_eps = float(kwargs.pop("ert_eps", 0.0) or 0.0)   # opt-in early ray termination (not a reference option)
    test_time = kwargs.get("test_time", False)
    Nc, Nf = kwargs["N_samples"], kwargs.get("N_importance", 0)
    perturb = kwargs.get("perturb", 0.)
    needs_grad = torch.is_grad_enabled() and ((c2w is not None and c2w.requires_grad) or
                                              (rays is not None and any(t.requires_grad for t in rays)))
    if needs_grad:
        # training path of train.py: gradient of rgb_map w.r.t. the pose / rays through the fine network
        if not test_time or perturb or Nf == 0:
            raise NotImplementedError("the differentiable render covers the test-time configuration train.py uses "
                                      "(render_kwargs_test with N_importance > 0)")
        h = _handle(kwargs)
        if c2w is not None and c2w.is_cuda and c2w.dim() == 2 and os.environ.get("DFB_POSE_RAYS_TORCH") != "1":
            rays_o, rays_d, viewdirs = _PoseRaysFn.apply(c2w, int(H), int(W), float(focal))
            sh = [int(H), int(W)]
        else:
            rays_o, rays_d = _get_rays_torch(int(H), int(W), float(focal), c2w) if c2w is not None else rays
            sh = list(rays_d.shape[:-1])
            rays_o, rays_d = rays_o.reshape(-1, 3).float(), rays_d.reshape(-1, 3).float()
            viewdirs = rays_d / torch.norm(rays_d, dim=-1, keepdim=True)
        n = rays_d.shape[0]
        hist = img_idx.to(rays_d.device).float().reshape(-1, img_idx.shape[-1])
        if hist.shape[0] != n:
            hist = hist[:1].expand(n, -1)
        nfh = torch.cat([torch.full((n, 1), float(near), device=rays_d.device),
                         torch.full((n, 1), float(far), device=rays_d.device), hist], -1)
        rgb, disp, acc = _RenderRaysFn.apply(rays_o, rays_d, viewdirs, nfh, h, Nc, Nf, mma)
        return [rgb.reshape(sh + [3]), disp.reshape(sh), acc.reshape(sh), {}]
    if c2w is not None and not perturb and not kwargs.get("raw_noise_std", 0.):
        # whole image from one pose: rays are generated in-kernel
        h = _handle(kwargs)
        want = []
        if Nf > 0 and not test_time:
            want += ["rgb0", "disp0", "acc0", "z_std", "transient_sigmas", "beta"]
        if kwargs.get("retraw"):
            want.append("raw")
        o = h.render(Nc, Nf, test_time, c2w=c2w, H=int(H), W=int(W), focal=float(focal), near=float(near),
                     far=float(far), hist=img_idx.to(c2w.device), mma=mma, lindisp=kwargs.get("lindisp", False),
                     ert_eps=ert_eps, want=want)
        sh = [int(H), int(W)]
        all_ret = {"rgb_map": o["rgb"], "disp_map": o["disp"], "acc_map": o["acc"], **{k: o[k] for k in want}}
    else:
        if c2w is not None:
            rays_o, rays_d = get_rays(H, W, focal, c2w)
        else:
            rays_o, rays_d = rays
        viewdirs = rays_d / torch.norm(rays_d, dim=-1, keepdim=True)
        viewdirs = torch.reshape(viewdirs, [-1, 3]).float()
        sh = list(rays_d.shape[:-1])
        rays_o = torch.reshape(rays_o, [-1, 3]).float()
        rays_d = torch.reshape(rays_d, [-1, 3]).float()
        nf = torch.ones_like(rays_d[..., :1])
        img_idx = img_idx.to(rays_d.device).float()
        if img_idx.shape[0] != rays_d.shape[0]:
            img_idx = img_idx.reshape(1, -1).repeat(rays_d.shape[0], 1)
        rec = torch.cat([rays_o, rays_d, near * nf, far * nf, viewdirs, img_idx], -1)
        kwargs.pop("use_viewdirs", None), kwargs.pop("ndc", None)
        all_ret = batchify_rays(rec, chunk, mma=mma, ert_eps=ert_eps, **{k: v for k, v in kwargs.items() if k not in ("ndc",)})
    for k in all_ret:
        all_ret[k] = torch.reshape(all_ret[k], sh + list(all_ret[k].shape[1:]))
    k_extract = ["rgb_map", "disp_map", "acc_map"]
    return [all_ret[k] for k in k_extract] + [{k: v for k, v in all_ret.items() if k not in k_extract}]


def write_png(path, img8):
    """8-bit RGB / greyscale PNG with the standard library only (the reference calls imageio.imwrite,
    rendering.py:441-452; imageio is used instead when it is installed)."""
    try:
        import imageio
        imageio.imwrite(path, img8)
        return
    except ImportError:
        pass
    a = np.ascontiguousarray(img8, dtype=np.uint8)
    if a.ndim == 2:
        a = a[..., None]
    h, w, c = a.shape
    ctype = {1: 0, 3: 2, 4: 6}[c]
    rows = np.concatenate([np.zeros((h, 1), np.uint8), a.reshape(h, w * c)], 1).tobytes()  # filter type 0 per scanline

    def chunk(tag, data):
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xffffffff)
    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, ctype, 0, 0, 0)) +
                chunk(b"IDAT", zlib.compress(rows, 3)) + chunk(b"IEND", b""))


def render_path(args, render_poses, hwf, chunk, render_kwargs, gt_imgs=None, savedir=None, render_factor=0,
                single_gt_img=False, img_ids=torch.Tensor(0), mma=None):
    """Validation render loop (reference rendering.py:403-458) -> (rgbs [n,H,W,3], disps [n,H,W]) numpy.

    Per image the reference runs render(...) and `.cpu().numpy()`; here one dfb_render_image_host call uploads the pose
    and histogram, renders and downloads rgb / disp into pinned host memory on the current stream.  Two sets of pinned
    buffers alternate, so image i+1 renders while image i is turned into PSNR and PNG files; the PNG encoder runs on a
    worker thread (off the stream's critical path).  Files and printed PSNR are the reference's: {i:03d}.png,
    {i:03d}_GT.png, {i:03d}_disp.png."""
    H, W, focal = hwf
    if render_factor != 0:
        H, W, focal = int(H // render_factor), int(W // render_factor), focal / render_factor
    H, W = int(H), int(W)
    kw = dict(render_kwargs)
    kw.pop("network_query_fn", None)
    mma = mma or kw.pop("mma", None) or DEFAULT_MMA
    near, far = float(kw.get("near", 0.)), float(kw.get("far", 1.))
    n = len(render_poses)
    fast = bool(kw.get("test_time")) and kw.get("N_importance", 0) > 0 and not kw.get("perturb") and not kw.get("raw_noise_std")
    rgbs, disps, psnr = [], [], []
    pool = ThreadPoolExecutor(max_workers=2) if savedir is not None else None
    jobs = []
    t0 = time.time()

    def consume(i, rgb_np, disp_np):
        rgbs.append(rgb_np), disps.append(disp_np)
        if i == 0:
            print(rgb_np.shape, disp_np.shape)
        if gt_imgs is not None:
            gt = gt_imgs if single_gt_img else gt_imgs[i]
            psnr.append(-10. * np.log10(np.mean(np.square(rgb_np - np.asarray(gt)))))
        if savedir is not None:
            jobs.append(pool.submit(write_png, os.path.join(savedir, "{:03d}.png".format(i)), to8b(rgb_np)))
            if gt_imgs is not None:
                jobs.append(pool.submit(write_png, os.path.join(savedir, "{:03d}_GT.png".format(i)), to8b(np.asarray(gt_imgs[i]))))
            jobs.append(pool.submit(write_png, os.path.join(savedir, "{:03d}_disp.png".format(i)), to8b(disp_np / np.max(disp_np))))

    if fast and n > 0:
        h = _handle(kw)
        dev = next(kw["network_fn"].parameters()).device
        cfg = _lib.RenderCfg(N_samples=kw["N_samples"], N_importance=kw["N_importance"], test_time=1, perturb=0,
                             mma_kind=_lib.MMA_KINDS[mma], lindisp=int(bool(kw.get("lindisp", False))), raw_noise_std=0.0)
        poses_h = torch.stack([torch.as_tensor(p)[:3, :4] for p in render_poses]).float().cpu().contiguous().pin_memory()
        hists_h = torch.stack([torch.as_tensor(img_ids[i]).reshape(-1) for i in range(n)]).float().cpu().contiguous().pin_memory()
        bufs = [(torch.empty(H * W, 3).pin_memory(), torch.empty(H * W).pin_memory(), torch.empty(H * W).pin_memory(),
                 torch.cuda.Event()) for _ in range(2)]
        stream = torch.cuda.current_stream(dev)

        def launch(i):
            rgb_h, disp_h, acc_h, ev = bufs[i & 1]
            h.render_image_host(cfg, poses_h[i], H, W, float(focal), near, far, hists_h[i], rgb_h, disp_h, acc_h, dev)
            ev.record(stream)
        launch(0)
        for i in range(n):
            rgb_h, disp_h, _, ev = bufs[i & 1]
            ev.synchronize()
            rgb_np, disp_np = rgb_h.numpy().reshape(H, W, 3).copy(), disp_h.numpy().reshape(H, W).copy()
            if i + 1 < n:
                launch(i + 1)   # the next image renders while this one is scored and encoded
            consume(i, rgb_np, disp_np)
    else:
        for i, c2w in enumerate(render_poses):
            rgb, disp, acc, _ = render(H, W, focal, chunk=chunk, c2w=c2w[:3, :4], img_idx=img_ids[i], mma=mma, **kw)
            consume(i, rgb.cpu().numpy(), disp.cpu().numpy())
    for j in jobs:
        j.result()
    if pool is not None:
        pool.shutdown()
    rgbs = np.stack(rgbs, 0) if rgbs else np.zeros((0, H, W, 3), np.float32)
    disps = np.stack(disps, 0) if disps else np.zeros((0, H, W), np.float32)
    print("Mean PSNR of this run is:", np.mean(psnr, 0) if psnr else float("nan"), "(%.2f s)" % (time.time() - t0))
    return rgbs, disps


def render_test(args, train_dl, val_dl, hwf, start, render_kwargs_test, decoder_coarse=None, decoder_fine=None):
    """Reference rendering.py:460-530: renders the training and validation views of the two loaders into
    basedir/expname/evaluate_{train,val}_{test|path}_{start:06d}/ with render_path.  Loader batches are the reference's
    (img [1,3,H,W], pose [1,12], hist [1,hist_bin]).  Videos (--render_video_*) need imageio and are written when it is
    installed."""
    dev = next(render_kwargs_test["network_fn"].parameters()).device
    for tag, dl, video, vname in (("train", train_dl, getattr(args, "render_video_train", False), "trainset"),
                                  ("val", val_dl, getattr(args, "render_video_test", False), "test")):
        savedir = os.path.join(args.basedir, args.expname,
                               "evaluate_{}_{}_{:06d}".format(tag, "test" if getattr(args, "render_test", False) else "path", start))
        os.makedirs(savedir, exist_ok=True)
        images, poses, index = [], [], []
        for img, pose, img_idx in dl:
            images.append(img.permute(0, 2, 3, 1))
            p = torch.zeros(1, 4, 4)
            p[0, :3, :4] = pose.reshape(3, 4)[:3, :4]
            p[0, 3, 3] = 1.
            poses.append(p), index.append(img_idx)
        images = torch.cat(images, 0).numpy()
        poses = torch.cat(poses, 0).to(dev)
        index = torch.cat(index, 0).to(dev)
        print(("train" if tag == "train" else "test") + " poses shape", poses.shape)
        with torch.no_grad():
            rgbs, disps = render_path(args, poses, hwf, args.chunk, render_kwargs_test, gt_imgs=images, savedir=savedir,
                                      img_ids=index)
        print("Saved {} set".format("train" if tag == "train" else "test"))
        if video:
            try:
                import imageio
            except ImportError:
                print("imageio is not installed: skipping the mp4 of the", tag, "set")
                continue
            base = os.path.join(args.basedir, args.expname, "{}_{}_{:06d}_".format(args.expname, vname, start))
            imageio.mimwrite(base + ("train" if tag == "train" else "test") + "_rgb.mp4", to8b(rgbs), fps=15, quality=8)
            imageio.mimwrite(base + ("train" if tag == "train" else "test") + "_disp.mp4", to8b(disps / np.max(disps)), fps=15, quality=8)
    return
