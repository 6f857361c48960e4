"""Host mirrors of the loss helpers on the feature path (reference feature/misc.py,
feature/direct_feature_matching.py, models/nerfw.py)."""
import ctypes as C
import math

import torch

from . import _lib
from ._lib import check, lib, raw_stream


def _p(t):
    return C.c_void_p(t.data_ptr())


def _stream():
    return raw_stream()


def _triplet_fwd(a, b, margin):
    L, B, Cc, H, W = a.shape
    loss = torch.empty((), device=a.device)
    case = torch.empty((), device=a.device, dtype=torch.int32)
    ws = torch.empty(8192, device=a.device)
    check(lib.dfb_triplet_loss(_p(a), _p(b), L, B, Cc, H, W, float(margin), _p(loss), _p(case), _p(ws), ws.numel() * 4, _stream()))
    return loss, case


class _TripletFn(torch.autograd.Function):
    """dfb_triplet_loss with dfb_triplet_loss_bwd (the mined case is a constant of the backward, as in the reference,
    where the four distances are computed under torch.no_grad())."""

    @staticmethod
    def forward(ctx, f1, f2, margin):
        a, b = f1.detach().float().contiguous(), f2.detach().float().contiguous()
        loss, case = _triplet_fwd(a, b, margin)
        ctx.save_for_backward(a, b, case)
        ctx.margin = float(margin)
        triplet_loss_hard_negative_mining_plus.last_case = case
        return loss

    @staticmethod
    def backward(ctx, g):
        a, b, case = ctx.saved_tensors
        L, B, Cc, H, W = a.shape
        g = g.float().contiguous()
        ga, gb = torch.empty_like(a), torch.empty_like(b)
        check(lib.dfb_triplet_loss_bwd(_p(a), _p(b), L, B, Cc, H, W, ctx.margin, _p(case), _p(g), _p(ga), _p(gb), _stream()))
        return (ga if ctx.needs_input_grad[0] else None, gb if ctx.needs_input_grad[1] else None, None)


def triplet_loss_hard_negative_mining_plus(f1, f2, margin=1.):
    """Reference feature/misc.py:399-435.  f1, f2: [lvl, B, C, H, W] -> scalar loss (differentiable w.r.t. f1, f2).
    The mined case of the last call is available as `triplet_loss_hard_negative_mining_plus.last_case`."""
    if not f1.is_cuda:
        raise _lib.DfbError("triplet loss inputs must be CUDA tensors")
    if torch.is_grad_enabled() and (f1.requires_grad or f2.requires_grad):
        return _TripletFn.apply(f1, f2, margin)
    loss, case = _triplet_fwd(f1.detach().float().contiguous(), f2.detach().float().contiguous(), margin)
    triplet_loss_hard_negative_mining_plus.last_case = case
    return loss


class _MseFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y):
        a, b = x.detach().float().contiguous(), y.detach().float().contiguous()
        out = torch.empty((), device=a.device)
        ws = torch.empty(1024, device=a.device)
        check(lib.dfb_mse(_p(a), _p(b), a.numel(), _p(out), _p(ws), ws.numel() * 4, _stream()))
        ctx.save_for_backward(a, b)
        ctx.shapes = (x.shape, y.shape)
        return out

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        ga = torch.empty_like(a)
        g = g.float().contiguous()
        check(lib.dfb_mse_bwd(_p(a), _p(b), a.numel(), _p(g), _p(ga), _stream()))
        return (ga.view(ctx.shapes[0]) if ctx.needs_input_grad[0] else None,
                (-ga).view(ctx.shapes[1]) if ctx.needs_input_grad[1] else None)


class _ResizeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, size, kind):
        a = x.detach().float().contiguous()
        B, Cc, h, w = a.shape
        out = torch.empty(B, Cc, size[0], size[1], device=a.device)
        fwd = lib.dfb_resize_bicubic if kind == "bicubic" else lib.dfb_resize_bilinear_ac
        check(fwd(_p(a), B * Cc, h, w, size[0], size[1], _p(out), _stream()))
        ctx.kind, ctx.dims = kind, (B, Cc, h, w, size[0], size[1])
        return out

    @staticmethod
    def backward(ctx, g):
        B, Cc, h, w, Ho, Wo = ctx.dims
        g = g.float().contiguous()
        gx = torch.empty(B, Cc, h, w, device=g.device)
        bwd = lib.dfb_resize_bicubic_bwd if ctx.kind == "bicubic" else lib.dfb_resize_bilinear_ac_bwd
        check(bwd(_p(g), B * Cc, h, w, Ho, Wo, _p(gx), _stream()))
        return gx, None, None


def mse(x, y):
    """nn.MSELoss()(x, y) / img2mse (reference models/nerfw.py:11)."""
    if not x.is_cuda:
        raise _lib.DfbError("mse inputs must be CUDA tensors")
    if torch.is_grad_enabled() and (x.requires_grad or y.requires_grad):
        return _MseFn.apply(x, y)
    a, b = x.detach().float().contiguous(), y.detach().float().contiguous()
    out = torch.empty((), device=a.device)
    ws = torch.empty(1024, device=a.device)
    check(lib.dfb_mse(_p(a), _p(b), a.numel(), _p(out), _p(ws), ws.numel() * 4, _stream()))
    return out


img2mse = mse


def mse2psnr(x):
    """-10 * ln(x) / ln(10) (reference models/nerfw.py:12)."""
    return -10. * torch.log(x) / math.log(10.)


def PoseLoss(args, pose_, pose, device=None):
    """Reference feature/direct_feature_matching.py:138-142."""
    return mse(pose_.reshape(args.batch_size, 12), pose)


def upsample_bicubic(x, size):
    """torch.nn.Upsample(size=size, mode='bicubic')(x) for x [B,C,h,w] (reference
    feature/direct_feature_matching.py:346): align_corners=False, output not clamped."""
    if not x.is_cuda:
        raise _lib.DfbError("upsample input must be a CUDA tensor")
    if torch.is_grad_enabled() and x.requires_grad:
        return _ResizeFn.apply(x, (int(size[0]), int(size[1])), "bicubic")
    a = x.detach().float().contiguous()
    B, Cc, h, w = a.shape
    out = torch.empty(B, Cc, size[0], size[1], device=a.device)
    check(lib.dfb_resize_bicubic(_p(a), B * Cc, h, w, size[0], size[1], _p(out), _stream()))
    return out


def upsample_bilinear_ac(x, size):
    """torch.nn.UpsamplingBilinear2d(size=size)(x) (align_corners=True; reference feature/dfnet.py:145)."""
    if not x.is_cuda:
        raise _lib.DfbError("upsample input must be a CUDA tensor")
    if torch.is_grad_enabled() and x.requires_grad:
        return _ResizeFn.apply(x, (int(size[0]), int(size[1])), "bilinear_ac")
    a = x.detach().float().contiguous()
    B, Cc, h, w = a.shape
    out = torch.empty(B, Cc, size[0], size[1], device=a.device)
    check(lib.dfb_resize_bilinear_ac(_p(a), B * Cc, h, w, size[0], size[1], _p(out), _stream()))
    return out


class _PolarFn(torch.autograd.Function):
    """R [B,3,3] -> U V^T of its SVD (what `u, s, v = torch.svd(R); u @ v.transpose(-2, -1)` computes, reference
    feature/direct_feature_matching.py:81-86) with the adjoint, both as one tiny kernel and without torch.svd's host-side
    status check, which drains the GPU in the middle of every training step."""

    @staticmethod
    def forward(ctx, R):
        A = R.detach().float().contiguous()
        n = A.shape[0]
        Q = torch.empty_like(A)
        aux = torch.empty(n, 21, dtype=torch.float64, device=A.device)
        check(lib.dfb_polar3x3_fwd(_p(A), n, _p(Q), _p(aux), _stream()))
        ctx.save_for_backward(aux)
        return Q

    @staticmethod
    def backward(ctx, g):
        aux, = ctx.saved_tensors
        G = g.float().contiguous()
        dA = torch.empty_like(G)
        check(lib.dfb_polar3x3_bwd(_p(aux), _p(G), G.shape[0], _p(dA), _stream()))
        return dA


def polar_orthogonalize(R):
    """Nearest rotation-like matrix U V^T of every 3x3 matrix of R [B,3,3] (CUDA), differentiable."""
    if not R.is_cuda:
        raise _lib.DfbError("polar_orthogonalize input must be a CUDA tensor: the dfnet_b200 path has no CPU fallback")
    return _PolarFn.apply(R)


def pose_errors(pred, gt, use_svd=True, return_fixed=False):
    """dfb_pose_error: pred, gt [n,12] (or [n,3,4]) CUDA tensors -> [n,2] = (translation error, rotation error in
    degrees); with return_fixed also the predicted poses with U V^T rotations."""
    if not pred.is_cuda:
        raise _lib.DfbError("pose_errors inputs must be CUDA tensors")
    p = pred.detach().float().reshape(-1, 12).contiguous()
    g = gt.detach().float().reshape(-1, 12).contiguous().to(p.device)
    n = p.shape[0]
    out = torch.empty(n, 2, device=p.device)
    fixed = torch.empty(n, 12, device=p.device) if return_fixed else None
    check(lib.dfb_pose_error(_p(p), _p(g), n, int(bool(use_svd)), _p(out), _p(fixed) if fixed is not None else None, _stream()))
    return (out, fixed) if return_fixed else out


def compute_error_in_q(args, dl, model, device, results, batch_size=1):
    """Reference feature/misc.py:49-107 with the same arguments and return value (results [n,2] filled with
    (metres, degrees); vis_info dict).  The reference pulls every prediction to the host, runs torch.svd and the
    quaternion conversion per image; here the predictions stay on the device and ONE dfb_pose_error launch scores the
    whole loader, followed by a single device-to-host copy."""
    import numpy as np
    preds, gts = [], []
    with torch.no_grad():
        for batch in dl:
            data, pose = batch[0], batch[1]
            _, predict_pose = model(data.to(device))
            preds.append(predict_pose.reshape(-1, 12))
            gts.append(pose.reshape(-1, 12))
    if not preds:
        return results, {"pose": np.zeros((0, 3)), "pose_gt": np.zeros((0, 3)), "theta": np.zeros((0,))}
    pred, gt = torch.cat(preds, 0), torch.cat(gts, 0).to(device)
    err, fixed = pose_errors(pred, gt, use_svd=True, return_fixed=True)
    err_h = err.cpu().numpy()
    n = err_h.shape[0]
    results[:n, :] = err_h
    vis_info_ret = {"pose": fixed.reshape(n, 3, 4)[:, :, 3].cpu().numpy(), "pose_gt": gt.reshape(n, 3, 4)[:, :, 3].cpu().numpy(),
                    "theta": err_h[:, 1].copy()}
    return results, vis_info_ret


def get_error_in_q(args, dl, model, sample_size, device, batch_size=1):
    """Reference feature/misc.py:110-124: median / mean translation and rotation error over the loader (printed in the
    reference's format; also returned)."""
    import numpy as np
    model.eval()
    results = np.zeros((sample_size, 2))
    results, vis_info = compute_error_in_q(args, dl, model, device, results, batch_size)
    median_result = np.median(results, axis=0)
    mean_result = np.mean(results, axis=0)
    print('Median error {}m and {} degrees.'.format(median_result[0], median_result[1]))
    print('Mean error {}m and {} degrees.'.format(mean_result[0], mean_result[1]))
    return median_result, mean_result


# ---------------------------------------------------------------------------------------------------------------------
# Random view synthesis for DFNet training (reference feature/misc.py:203-289): NeRF renders of the training poses and of
# perturbed ("virtual") poses.  The reference renders one view per render() call and pulls every image to the host; here a
# whole batch of poses is rendered by ONE dfb_render_poses_fwd call and upsampled in one batched bicubic launch.
# ---------------------------------------------------------------------------------------------------------------------
def _render_pose_batch(args, poses_nerf, img_idxs, hwf, render_kwargs_test, poses_per_call=64):
    from . import ops
    from .rendering import DEFAULT_MMA
    H, W, focal = hwf
    H, W = int(H), int(W)
    kw = render_kwargs_test
    h = ops.handle_for(kw["network_fn"], kw.get("network_fine"), kw.get("embedding_a"), kw.get("embedding_t"))
    tiny = bool(getattr(args, "tinyimg", False))
    s = float(args.tinyscale) if tiny else 1.0
    rh, rw, rf = int(H // s), int(W // s), focal / s
    outs = []
    for i in range(0, poses_nerf.shape[0], poses_per_call):
        o = h.render_poses(kw["N_samples"], kw["N_importance"], poses_nerf[i:i + poses_per_call], img_idxs[i:i + poses_per_call], rh, rw,
                           rf, kw.get("near", 0.), kw.get("far", 1.), mma=kw.get("mma") or DEFAULT_MMA, lindisp=kw.get("lindisp", False))
        rgb = o["rgb"]
        if tiny:
            rgb = upsample_bicubic(rgb.permute(0, 3, 1, 2), (H, W)).permute(0, 2, 3, 1)
        outs.append(rgb)
    return torch.cat(outs, 0)


def _fix_coord_supp(args, pose, world_setup_dict):
    import numpy as np
    pose = pose.clone()
    pose[:, :3, 3] *= world_setup_dict["pose_scale"]
    pose[:, :3, 3] += torch.as_tensor(np.asarray(world_setup_dict["move_all_cam_vec"], np.float32), device=pose.device)
    pose[:, :3, 3] *= world_setup_dict["pose_scale2"]
    return pose


def render_virtual_imgs(args, pose_perturb, img_idxs, hwf, device, render_kwargs_test, world_setup_dict):
    """Reference feature/misc.py:249-289 -> rgbs [n,H,W,3] (host tensor, like the reference).  pose_perturb [n,3,4] in the
    dataset frame (rescaled to the NeRF frame like fix_coord_supp), img_idxs [n,1,hist_bin] / [n,hist_bin]."""
    poses = _fix_coord_supp(args, torch.as_tensor(pose_perturb, dtype=torch.float32).reshape(-1, 3, 4).to(device), world_setup_dict)
    hists = torch.as_tensor(img_idxs, dtype=torch.float32).reshape(poses.shape[0], -1).to(device)
    with torch.no_grad():
        return _render_pose_batch(args, poses, hists, hwf, render_kwargs_test).cpu()


def render_nerfw_imgs(args, dl, hwf, device, render_kwargs_test, world_setup_dict):
    """Reference feature/misc.py:203-247 -> (targets [n,H,W,3], rgbs [n,H,W,3], poses [n,3,4], img_idxs [n,1,hist_bin])
    host tensors: the NeRF render of every training view next to its photograph."""
    targets, poses, idxs = [], [], []
    for target, pose, img_idx in dl:
        targets.append(target[0].permute(1, 2, 0))
        poses.append(pose.reshape(3, 4))
        idxs.append(img_idx)
    poses_t = torch.stack(poses)
    idx_t = torch.stack(idxs)
    with torch.no_grad():
        rgbs = _render_pose_batch(args, _fix_coord_supp(args, poses_t.to(device).float(), world_setup_dict),
                                  idx_t.reshape(len(poses), -1).to(device).float(), hwf, render_kwargs_test).cpu()
    return torch.stack(targets).detach().cpu(), rgbs, poses_t.cpu(), idx_t.cpu()
