"""Host-side mirror of the training step of the reference's `feature/direct_feature_matching.py`
(`train_on_batch` :322-390, `inference_pose_regression` :63-93, `rgb_loss` :95-100, `PoseLoss` :138-142) and
`dm/direct_pose_model.py:147-167` (`fix_coord_supp`).

Everything heavy runs on the sm_100a kernels with hand-written backward passes: the pose regressor (tcgen05 forward,
data- and weight-gradient convolutions), the NeRF-Hist render at H//4 x W//4 (gradient w.r.t. the pose through
dfb_render_bwd), the bicubic x4 upsampling and its adjoint, the frozen feature net (data-gradient chain) and the
cosine / MSE losses.  torch is the host runtime: autograd graph edges, the 3x3 SVD, the [1,3,4] pose arithmetic and
the optimizer.
"""
import numpy as np
import torch
import torch.distributed as dist

from . import parallel
from .dfnet import feature_loss, preprocess_features_for_loss  # noqa: F401  (same import surface as the reference)
from .misc import PoseLoss, img2mse, mse2psnr, upsample_bicubic
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
        R_torch = pose[:, :3, :3].clone()
        u, s, v = torch.svd(R_torch)
        Rs = torch.matmul(u, v.transpose(-2, -1))
        pose[:, :3, :3] = Rs
    return features, pose


def fix_coord_supp(args, pose, world_setup_dict, device=None):
    """Reference dm/direct_pose_model.py:147-167: predicted pose -> NeRF world scale (in place, like the reference)."""
    sc = world_setup_dict["pose_scale"]
    move_all_cam_vec = torch.as_tensor(np.asarray(world_setup_dict["move_all_cam_vec"], np.float32), device=pose.device)
    sc2 = world_setup_dict["pose_scale2"]
    pose[:, :3, 3] *= sc
    pose[:, :3, 3] += move_all_cam_vec
    pose[:, :3, 3] *= sc2
    return pose


def rgb_loss(rgb, target, extras=None):
    """Reference feature/direct_feature_matching.py:95-100."""
    return img2mse(rgb, target)


def train_on_batch(args, data, model, feat_model, pose, img_idx, hwf, optimizer, half_res, device, world_setup_dict,
                   **render_kwargs_test):
    """One optimisation step (reference feature/direct_feature_matching.py:322-390), same arguments and return value
    (iter_loss [1] numpy, iter_psnr numpy).  When torch.distributed is initialised with more than one rank the pose
    regressor's gradients are averaged with a single all-reduce before the optimizer step."""
    H, W, focal = hwf
    H, W = int(H), int(W)
    data = data.to(device)

    # pose regression module
    _, pose_ = inference_pose_regression(args, data, device, model, retFeature=False)
    pose_nerf = pose_.clone()
    # rescale the predicted pose to nerf scales
    pose_nerf = fix_coord_supp(args, pose_nerf, world_setup_dict, device=device)
    pose = pose.to(device)
    img_idx = img_idx.to(device)

    # direct matching module
    if half_res:
        rgb, disp, acc, extras = render(H // 4, W // 4, focal / 4, chunk=args.chunk, c2w=pose_nerf[0, :3, :4], img_idx=img_idx,
                                        **render_kwargs_test)
        rgb = rgb[None, ...].permute(0, 3, 1, 2)
        rgb = upsample_bicubic(rgb, (H, W))
    else:
        rgb, disp, acc, extras = render(H, W, focal, chunk=args.chunk, c2w=pose_nerf[0, :3, :4], img_idx=img_idx,
                                        **render_kwargs_test)
        rgb = rgb[None, ...].permute(0, 3, 1, 2)

    # feature metric module
    feat_model.grad_levels = list(args.feature_matching_lvl)  # levels whose gradient is non-zero after index_select
    feature_list, _ = inference_pose_regression(args, torch.cat([data, rgb]), device, feat_model, retFeature=True,
                                                isSingleStream=False, return_pose=False)
    feature_target, feature_rgb = feature_list[0], feature_list[1]

    photo_loss = rgb_loss(rgb, data, extras)
    indices = torch.tensor(args.feature_matching_lvl, device=feature_rgb.device)
    feature_rgb = torch.index_select(feature_rgb, 0, indices)
    feature_target = torch.index_select(feature_target, 0, indices)
    feature_rgb = preprocess_features_for_loss(feature_rgb)
    feature_target = preprocess_features_for_loss(feature_target)
    feat_loss = feature_loss(feature_rgb[0], feature_target[0], per_channel=args.per_channel)

    if not args.combine_loss:
        raise ValueError("train_on_batch needs --combine_loss (the reference leaves `loss` undefined without it, "
                         "feature/direct_feature_matching.py:373-378)")
    pose_loss = PoseLoss(args, pose_, pose, device)
    loss = args.combine_loss_w[0] * pose_loss + args.combine_loss_w[1] * photo_loss + args.combine_loss_w[2] * feat_loss

    loss.backward()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        parallel.allreduce_gradients_(model.parameters())
    optimizer.step()
    optimizer.zero_grad()
    psnr = mse2psnr(img2mse(rgb.detach(), data))

    iter_loss = np.array([loss.detach().cpu().numpy()])
    iter_psnr = psnr.detach().cpu().numpy()
    return iter_loss, iter_psnr
