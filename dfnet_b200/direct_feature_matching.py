"""Host-side mirror of the training step of the reference's `feature/direct_feature_matching.py`
(`train_on_batch` :322-390, `inference_pose_regression` :63-93, `rgb_loss` :95-100, `PoseLoss` :138-142) and
`dm/direct_pose_model.py:147-167` (`fix_coord_supp`).

Everything heavy runs on the sm_100a kernels with hand-written backward passes: the pose regressor (tcgen05 forward,
data- and weight-gradient convolutions), the NeRF-Hist render at H//4 x W//4 (gradient w.r.t. the pose through
dfb_render_bwd), the bicubic x4 upsampling and its adjoint, the frozen feature net (data-gradient chain) and the
cosine / MSE losses.  torch is the host runtime: autograd graph edges, the 3x3 SVD, the [1,3,4] pose arithmetic and
the optimizer.
"""
import os

import numpy as np
import torch
import torch.distributed as dist

from . import parallel
from .dfnet import feature_loss, preprocess_features_for_loss  # noqa: F401  (same import surface as the reference)
from .misc import PoseLoss, img2mse, mse2psnr, polar_orthogonalize, upsample_bicubic
from .rendering import render

_MEAN = (0.485, 0.456, 0.406)
_STD = (0.229, 0.224, 0.225)


def preprocess_data(inputs, device=None):
    """Reference dm/pose_model.py:18-23 (only taken with --preprocess_ImgNet; DFNet normalises on its own)."""
    mean = torch.tensor(_MEAN, device=inputs.device)
    std = torch.tensor(_STD, device=inputs.device)
    return (inputs - mean[None, :, None, None]) / std[None, :, None, None]


def inference_pose_regression(args, data, device, model, retFeature=False, isSingleStream=True, return_pose=True):
    """Reference feature/direct_feature_matching.py:63-93 -> (features, pose [B,3,4]) for the DFNet family."""
    inputs = data.to(device)
    _, _, H, W = data.size()
    if getattr(args, "preprocess_ImgNet", False):
        inputs = preprocess_data(inputs, device)
    if not getattr(args, "DFNet", True):
        raise NotImplementedError("only the DFNet / DFNet_s pose regressors are on the B200 path (reference --DFNet)")
    features, predict_pose = model(inputs, return_feature=retFeature, isSingleStream=isSingleStream, return_pose=return_pose,
                                   upsampleH=H, upsampleW=W)
    if not return_pose:
        return features, predict_pose
    pose = predict_pose.reshape(inputs.shape[0], 3, 4)
    if getattr(args, "svd_reg", False):
        pose[:, :3, :3] = polar_orthogonalize(pose[:, :3, :3].clone())   # u v^T of torch.svd, without its host sync
    return features, pose


_MOVE_VEC_CACHE = {}


def fix_coord_supp(args, pose, world_setup_dict, device=None):
    """Reference dm/direct_pose_model.py:147-167: predicted pose -> NeRF world scale (in place, like the reference)."""
    sc = world_setup_dict["pose_scale"]
    # the offset is a constant of the run: one upload per (value, device) instead of a pageable host-to-device copy per step
    key = (tuple(float(x) for x in world_setup_dict["move_all_cam_vec"]), str(pose.device))
    move_all_cam_vec = _MOVE_VEC_CACHE.get(key)
    if move_all_cam_vec is None:
        if len(_MOVE_VEC_CACHE) > 64:
            _MOVE_VEC_CACHE.clear()
        move_all_cam_vec = torch.as_tensor(np.asarray(world_setup_dict["move_all_cam_vec"], np.float32), device=pose.device)
        _MOVE_VEC_CACHE[key] = move_all_cam_vec
    sc2 = world_setup_dict["pose_scale2"]
    pose[:, :3, 3] *= sc
    pose[:, :3, 3] += move_all_cam_vec
    pose[:, :3, 3] *= sc2
    return pose


def rgb_loss(rgb, target, extras=None):
    """Reference feature/direct_feature_matching.py:95-100."""
    return img2mse(rgb, target)


def predict_pose(args, data, model, world_setup_dict, device):
    """Stage 1: pose regression -> (pose_ [B,3,4] in the regressor's frame, pose_nerf in NeRF world scale).
    Reference feature/direct_feature_matching.py:329-335."""
    _, pose_ = inference_pose_regression(args, data, device, model, retFeature=False)
    return pose_, fix_coord_supp(args, pose_.clone(), world_setup_dict, device=device)


def render_prediction(args, c2w, img_idx, hwf, half_res, render_kwargs):
    """Stage 2: the NeRF-Hist view at the predicted pose as a [1,3,H,W] image; at quarter resolution followed by the
    x4 bicubic upsampling when half_res (reference :341-352).  Differentiable w.r.t. c2w."""
    H, W, focal = hwf
    H, W = int(H), int(W)
    s = 4 if half_res else 1
    rgb, _, _, extras = render(H // s, W // s, focal / s, chunk=args.chunk, c2w=c2w, img_idx=img_idx, **render_kwargs)
    rgb = rgb[None, ...].permute(0, 3, 1, 2)
    return (upsample_bicubic(rgb, (H, W)) if half_res else rgb), extras


def matching_terms(args, data, rgb, feat_model, device):
    """Stage 3: photometric MSE and the cosine feature-metric loss between the target image and the rendered view
    through the frozen feature net (reference :354-370).  Only the rendered stream carries gradient."""
    # hints for the feature net, scoped to this call: only feature_matching_lvl is read afterwards, so the other levels are
    # neither computed (want_levels; with level 0 alone the encoder stops after conv1_2) nor differentiated (grad_levels)
    prev = (getattr(feat_model, "grad_levels", None), getattr(feat_model, "want_levels", None))
    feat_model.grad_levels = feat_model.want_levels = list(args.feature_matching_lvl)
    if os.environ.get("DFB_ALL_LEVELS") == "1":   # A/B: evaluate every level like the reference does (and discards)
        feat_model.want_levels = None
    try:
        (feature_target, feature_rgb), _ = inference_pose_regression(args, torch.cat([data, rgb]), device, feat_model, retFeature=True,
                                                                     isSingleStream=False, return_pose=False)
    finally:
        feat_model.grad_levels, feat_model.want_levels = prev
    f_rgb = preprocess_features_for_loss(_take_levels(feature_rgb, args.feature_matching_lvl, feat_model))
    # the target stream is a constant of the step (frozen feature net, image without gradient)
    f_tgt = preprocess_features_for_loss(_take_levels(feature_target.detach(), args.feature_matching_lvl, feat_model))
    return rgb_loss(rgb, data), feature_loss(f_rgb[0], f_tgt[0], per_channel=args.per_channel)


class _TakeLevels(torch.autograd.Function):
    """stack[lv0:lv0+n] as a view.  The reference's torch.index_select (:365-366) copies the selected levels (157 MB per
    level at 480x640) and its backward fills a zero stack and index_adds into it; here the forward is a view and the backward
    hands the feature net a stack in which ONLY the selected levels are written - the consumer was told through
    `grad_levels` (set by matching_terms) to read nothing else."""

    @staticmethod
    def forward(ctx, stack, lv0, n):
        ctx.shape, ctx.lv0, ctx.n = stack.shape, lv0, n
        ctx.set_materialize_grads(False)
        return stack.narrow(0, lv0, n)

    @staticmethod
    def backward(ctx, g):
        if g is None:
            return None, None, None
        full = torch.empty(ctx.shape, device=g.device, dtype=g.dtype)
        full.narrow(0, ctx.lv0, ctx.n).copy_(g)
        return full, None, None


def _take_levels(stack, levels, feat_model):
    lv = [int(l) for l in levels]
    from .dfnet import DFNet
    contiguous = lv == list(range(lv[0], lv[0] + len(lv)))
    if contiguous and not (torch.is_grad_enabled() and stack.requires_grad):
        return stack.narrow(0, lv[0], len(lv))
    if contiguous and isinstance(feat_model, DFNet):   # matching_terms told it to differentiate exactly these levels
        return _TakeLevels.apply(stack, lv[0], len(lv))
    return torch.index_select(stack, 0, torch.tensor(lv, device=stack.device))


def apply_gradients(model, optimizer):
    """Stage 4: optimizer step on the pose regressor.  Data parallel (torch.distributed initialised, world > 1): the
    gradients were averaged by the GradSync attached to the regressor's handle WHILE its backward ran (dfnet.py); here
    the compute stream only waits for the collective and the parameters' .grad are pointed at the averaged flat bucket.
    Modules without such a handle fall back to parallel.allreduce_gradients_."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        h = getattr(model, "_handle", None)
        sync = getattr(h, "grad_sync", None)
        if sync is not None and h.last_grad_views is not None:
            sync.finish()
            for p, g in zip(model._load_order_params(), h.last_grad_views):
                if p.grad is not None:
                    p.grad = g
            h.last_grad_views = None
        else:
            parallel.allreduce_gradients_(model.parameters())
    optimizer.step()
    optimizer.zero_grad()
    # (measured and dropped: re-packing the updated weights here, behind the optimizer step, instead of in front of the next
    # forward - the host is not ahead of the GPU at this point, the step got 0.2 ms slower)


def train_on_batch(args, data, model, feat_model, pose, img_idx, hwf, optimizer, half_res, device, world_setup_dict,
                   **render_kwargs_test):
    """One optimisation step with the reference's signature and return value (feature/direct_feature_matching.py:322-390:
    iter_loss [1] numpy, iter_psnr numpy), composed of the four stages above.  Differences in mechanism, not in
    arithmetic: every stage runs on the sm_100a kernels with hand-written backward passes; under torch.distributed the
    pose regressor's gradients are averaged with ONE bucketed all-reduce overlapped with its backward; loss
    and PSNR come back in a single device-to-host copy."""
    if not args.combine_loss:
        raise ValueError("train_on_batch needs --combine_loss (the reference leaves `loss` undefined without it, "
                         "feature/direct_feature_matching.py:373-378)")
    data, pose, img_idx = data.to(device), pose.to(device), img_idx.to(device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1 and getattr(model, "_handle", None) is not None \
            and model._handle.grad_sync is None:
        model._handle.grad_sync = parallel.GradSync()
    pose_, pose_nerf = predict_pose(args, data, model, world_setup_dict, device)
    rgb, _ = render_prediction(args, pose_nerf[0, :3, :4], img_idx, hwf, half_res, render_kwargs_test)
    photo_loss, feat_loss = matching_terms(args, data, rgb, feat_model, device)
    w = args.combine_loss_w
    loss = w[0] * PoseLoss(args, pose_, pose, device) + w[1] * photo_loss + w[2] * feat_loss
    loss.backward()
    apply_gradients(model, optimizer)
    with torch.no_grad():
        host = torch.stack([loss.detach().reshape(()), mse2psnr(img2mse(rgb.detach(), data)).reshape(())]).cpu().numpy()
    return np.array([host[0]]), host[1]


def prepare_batch_render(args, pose, batch_size, target_, H, W, focal, half_res=True, rand=True):
    """Reference feature/direct_feature_matching.py:144-176: a random batch of rays and their target colours from a batch
    of images -> (batch_rays [2,N,3], target_s [N,3]).  Rays come from dfb_get_rays, the half-resolution targets from
    dfb_resize_area (the reference round-trips through cv2.resize on the host), the selection is a gather on the device;
    the permutation is drawn with torch.randperm on the CPU generator, as in the reference."""
    from . import data as _data
    from . import ops
    dev = pose.device if pose.is_cuda else torch.device("cuda", torch.cuda.current_device())
    tgt = target_.to(dev).permute(0, 2, 3, 1).contiguous()   # [B,H,W,3]
    if half_res:
        N_rand = batch_size * (H // 2) * (W // 2)
        # the reference passes dims=(H//2, W//2) to cv2.resize, i.e. (width, height) = (H//2, W//2) (:149)
        tgt = torch.stack([_data.resize_area(tgt[i], (H // 2, W // 2)) for i in range(batch_size)], 0)
        rh, rw, rf = H // 2, W // 2, focal / 2
    else:
        N_rand = args.N_rand
        rh, rw, rf = H, W, focal
    rays = torch.stack([torch.stack(ops.get_rays(rh, rw, rf, pose[i].to(dev)), 0) for i in range(batch_size)], 0)  # [B,2,h,w,3]
    rays_rgb = torch.cat((rays, tgt[:, None, ...]), 1).permute(0, 2, 3, 1, 4).reshape(-1, 3, 3)
    sel = torch.randperm(rays_rgb.shape[0])[:N_rand].to(dev)
    batch = rays_rgb[sel].permute(1, 0, 2)
    return batch[:2], batch[2]


def eval_on_batch(args, data, model, feat_model, pose, img_idx, hwf, half_res, device, world_setup_dict, **render_kwargs_test):
    """One evaluation step (reference feature/direct_feature_matching.py:178-213): pose loss of the regressor and the PSNR
    of N_rand random rays rendered at the predicted pose -> (iter_loss [1], iter_psnr) numpy."""
    with torch.no_grad():
        H, W, focal = hwf
        H, W = int(H), int(W)
        data = data.to(device)
        _, pose_ = inference_pose_regression(args, data, device, model)
        pose_nerf = fix_coord_supp(args, pose_.clone(), world_setup_dict, device=device) if getattr(args, "NeRFH", True) else pose_.clone()
        batch_rays, target = prepare_batch_render(args, pose_nerf, args.batch_size, data, H, W, focal, False)
        rgb, _, _, _ = render(H, W, focal, chunk=args.chunk, rays=batch_rays, img_idx=img_idx.to(device), **render_kwargs_test)
        loss = PoseLoss(args, pose_, pose.to(device), device)
        psnr = mse2psnr(img2mse(rgb, target))
        host = torch.stack([loss.reshape(()), psnr.reshape(())]).cpu().numpy()
    return np.array([host[0]]), host[1]
