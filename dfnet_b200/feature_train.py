"""DFNet training (reference script/run_feature.py:102-230): one optimisation step on a batch of
(photograph, NeRF render, pose) triples, optionally with a random-view-synthesis batch, and the epoch loops around it.

Every parameter of DFNet is trained - encoder, adaptation heads (1x1, 5x5, train-mode BatchNorm), fc_pose - through the
hand-written kernels of dfnet_b200.dfnet (one siamese forward returning features AND pose, dfb_dfnet_bwd with the head tape)
and dfnet_b200.misc (triplet loss with hard negative mining, MSE)."""
import numpy as np
import torch

from .misc import PoseLoss, triplet_loss_hard_negative_mining_plus
from .dfnet import feature_loss, preprocess_features_for_loss


def _feature_loss_fn(args, FeatureLoss):
    if getattr(args, "tripletloss", False):
        return lambda fr, ft: triplet_loss_hard_negative_mining_plus(fr, ft, margin=args.triplet_margin)
    if FeatureLoss is not None:
        return FeatureLoss
    return lambda fr, ft: feature_loss(preprocess_features_for_loss(fr)[0], preprocess_features_for_loss(ft)[0])


def feature_train_step(args, feat_model, target_in, rgb_in, pose, optimizer, hwf, FeatureLoss=None, rgb_perturb=None, pose_perturb=None,
                       device=None):
    """One iteration of run_feature.py's loops (:128-160 without, :198-225 with random view synthesis).
    target_in, rgb_in [B,3,H,W]; pose [B,12]; rgb_perturb [B,3,H,W] / pose_perturb [B,12] (RVS) -> loss (device scalar)."""
    H, W, _ = hwf
    H, W = int(H), int(W)
    device = device or target_in.device
    B = target_in.shape[0]
    args_b = type("A", (), {"batch_size": 2 * B})()
    pose2 = torch.cat([pose, pose]).to(device)
    features, predict_pose = feat_model(torch.cat([target_in, rgb_in]).to(device), return_feature=True, upsampleH=H, upsampleW=W)
    features_target, features_rgb = features[0], features[1]
    floss = _feature_loss_fn(args, FeatureLoss)
    if getattr(args, "poselossonly", False):
        loss = PoseLoss(args_b, predict_pose, pose2, device)
    elif getattr(args, "featurelossonly", False):
        loss = floss(features_rgb, features_target)
    else:
        loss_pose = PoseLoss(args_b, predict_pose, pose2, device)
        loss_f = floss(features_rgb, features_target)
        if rgb_perturb is None:
            loss = loss_pose + loss_f
        else:
            _, virtue_pose = feat_model(rgb_perturb.to(device), False)
            loss_pp = PoseLoss(type("A", (), {"batch_size": B})(), virtue_pose, pose_perturb.to(device), device)
            w = args.combine_loss_w
            loss = w[0] * loss_pose + w[1] * loss_f + w[2] * loss_pp
    loss.backward()
    optimizer.step()
    optimizer.zero_grad()
    return loss.detach()


def _epoch(args, feat_model, dset_size, optimizer, hwf, FeatureLoss, fetch):
    feat_model.train()
    if getattr(args, "freezeBN", False):
        for m in feat_model.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.eval()
    select_inds = np.random.choice(dset_size, size=[dset_size], replace=False)
    bs = args.featurenet_batch_size
    losses = []
    for i0 in range(0, dset_size - bs + 1, bs):
        losses.append(feature_train_step(args, feat_model, optimizer=optimizer, hwf=hwf, FeatureLoss=FeatureLoss, **fetch(select_inds[i0:i0 + bs])))
    return float(torch.stack(losses).mean()) if losses else float("nan")    # ONE host read per epoch (the reference: one per step)


def train_on_batch(args, targets, rgbs, poses, feat_model, dset_size, FeatureLoss, optimizer, hwf):
    """Reference run_feature.py:102-163 (same arguments; targets / rgbs [n,H,W,3], poses [n,3,4] host or device tensors)."""
    dev = next(feat_model.parameters()).device

    def fetch(ii):
        return dict(target_in=targets[ii].permute(0, 3, 1, 2).to(dev), rgb_in=rgbs[ii].permute(0, 3, 1, 2).to(dev),
                    pose=poses[ii].reshape(len(ii), 12).to(dev))
    return _epoch(args, feat_model, dset_size, optimizer, hwf, FeatureLoss, fetch)


def train_on_batch_with_random_view_synthesis(args, targets, rgbs, poses, virtue_view, poses_perturb, feat_model, dset_size, FeatureLoss,
                                              optimizer, hwf, img_idxs=None, render_kwargs_test=None):
    """Reference run_feature.py:165-230."""
    dev = next(feat_model.parameters()).device

    def fetch(ii):
        return dict(target_in=targets[ii].permute(0, 3, 1, 2).to(dev), rgb_in=rgbs[ii].permute(0, 3, 1, 2).to(dev),
                    pose=poses[ii].reshape(len(ii), 12).to(dev), rgb_perturb=virtue_view[ii].permute(0, 3, 1, 2).to(dev),
                    pose_perturb=poses_perturb[ii].reshape(len(ii), 12).to(dev))
    return _epoch(args, feat_model, dset_size, optimizer, hwf, FeatureLoss, fetch)
