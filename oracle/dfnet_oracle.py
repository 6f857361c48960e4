"""CPU oracle for the DFNet feature path (TEST INFRASTRUCTURE, not product code).

numpy restatement of DFNet.forward and feature_loss; each function cites the reference lines
(paths relative to /root/reference/script).  Pinned against vectors generated from the unmodified
reference by tests/golden/make_golden_dfnet.py (tests/golden/dfnet_golden.npz).  Only tests/,
__graft_entry__.smoke() and bench.py's CPU legs may import it.
"""
import numpy as np

f32 = np.float32
VGG_CFG = [64, 64, "M", 128, 128, "M", 256, 256, 256, "M", 512, 512, 512, "M", 512, 512, 512, "M"]
MEAN = np.array([0.485, 0.456, 0.406], f32)
STD = np.array([0.229, 0.224, 0.225], f32)


def conv2d(x, w, b, pad):
    """nn.Conv2d stride 1 (im2col + matmul).  x [B,C,H,W], w [O,C,kh,kw]."""
    B, Cc, H, W = x.shape
    O, _, kh, kw = w.shape
    xp = np.pad(x, ((0, 0), (0, 0), (pad, pad), (pad, pad)))
    cols = np.empty((B, Cc, kh, kw, H, W), f32)
    for i in range(kh):
        for j in range(kw):
            cols[:, :, i, j] = xp[:, :, i:i + H, j:j + W]
    out = np.einsum("ok,bkp->bop", w.reshape(O, -1), cols.reshape(B, Cc * kh * kw, H * W), optimize=True)
    return (out.reshape(B, O, H, W) + b[None, :, None, None]).astype(f32)


def maxpool2(x):
    B, Cc, H, W = x.shape
    x = x[:, :, : H // 2 * 2, : W // 2 * 2]
    return x.reshape(B, Cc, H // 2, 2, W // 2, 2).max((3, 5))


def upsample_bilinear_ac(x, Ho, Wo):
    """torch.nn.UpsamplingBilinear2d(size) == align_corners=True (feature/dfnet.py:145,156-157)."""
    B, Cc, h, w = x.shape
    sy = f32(h - 1) / f32(Ho - 1) if Ho > 1 else f32(0)
    sx = f32(w - 1) / f32(Wo - 1) if Wo > 1 else f32(0)
    fy = (sy * np.arange(Ho, dtype=f32)).astype(f32)
    fx = (sx * np.arange(Wo, dtype=f32)).astype(f32)
    y0 = fy.astype(np.int64)
    x0 = fx.astype(np.int64)
    y1 = np.minimum(y0 + 1, h - 1)
    x1 = np.minimum(x0 + 1, w - 1)
    ly = (fy - y0).astype(f32)[None, None, :, None]
    lx = (fx - x0).astype(f32)[None, None, None, :]
    g = lambda yy, xx: x[:, :, yy][:, :, :, xx]
    return ((1 - ly) * ((1 - lx) * g(y0, x0) + lx * g(y0, x1)) + ly * ((1 - lx) * g(y1, x0) + lx * g(y1, x1))).astype(f32)


def dfnet_forward(P, x, n_levels=3, return_feature=True, single=False, return_pose=True, upH=None, upW=None, bn_eps=1e-5,
                  bn_train=False, bn_momentum=0.1, bn_running=None):
    """feature/dfnet.py:106-172.  P: state_dict as numpy arrays.  bn_train=False: eval-mode BatchNorm (freezeBN /
    model.eval()); bn_train=True: batch statistics over the whole batch as torch.nn.BatchNorm2d computes them under
    model.train() (run_feature.py:133,204), and `bn_running` (a dict) receives the updated running_mean / running_var
    of every head (momentum update with the unbiased batch variance)."""
    x = ((np.asarray(x, f32) - MEAN[None, :, None, None]) / STD[None, :, None, None]).astype(f32)
    taps, idx = [], 0
    tap_idx = [2, 14, 28][:n_levels]
    for v in VGG_CFG:
        if v == "M":
            x = maxpool2(x)
            idx += 1
            continue
        x = conv2d(x, P[f"encoder.{idx}.weight"], P[f"encoder.{idx}.bias"], 1)
        if idx in tap_idx:
            taps.append(x.copy())
            if idx == tap_idx[-1] and not return_pose:
                break
        x = np.maximum(x, 0)
        idx += 2
    feats = None
    if return_feature:
        outs = []
        for l, t in enumerate(taps):
            p = f"adaptation_layers.adapt_layer_{l}."
            h = np.maximum(conv2d(t, P[p + "0.weight"], P[p + "0.bias"], 0), 0)
            h = conv2d(h, P[p + "2.weight"], P[p + "2.bias"], 2)
            if bn_train:
                n = h.shape[0] * h.shape[2] * h.shape[3]
                mu = h.mean((0, 2, 3), dtype=np.float64)
                var = h.var((0, 2, 3), dtype=np.float64)   # biased: what normalises the batch
                if bn_running is not None:
                    bn_running[p + "3.running_mean"] = ((1 - bn_momentum) * P[p + "3.running_mean"] + bn_momentum * mu).astype(f32)
                    bn_running[p + "3.running_var"] = ((1 - bn_momentum) * P[p + "3.running_var"]
                                                       + bn_momentum * var * n / max(n - 1, 1)).astype(f32)
                mu, var = mu.astype(f32), var.astype(f32)
            else:
                mu, var = P[p + "3.running_mean"], P[p + "3.running_var"]
            sc = P[p + "3.weight"] / np.sqrt(var + f32(bn_eps))
            h = ((h - mu[None, :, None, None]) * sc[None, :, None, None]
                 + P[p + "3.bias"][None, :, None, None]).astype(f32)
            outs.append(upsample_bilinear_ac(h, upH, upW))
        stack = np.stack(outs)  # [L,B,128,H,W]
        feats = [stack] if single else [stack[:, : stack.shape[1] // 2], stack[:, stack.shape[1] // 2:]]
    pose = None
    if return_pose:
        pose = (x.mean((2, 3)) @ P["fc_pose.weight"].T + P["fc_pose.bias"]).astype(f32)
    return feats, pose


def feature_loss(fr, ft, per_channel=False, eps=1e-6):
    """feature/direct_feature_matching.py:114-136; torch >= 1.12 clamps each norm separately."""
    Cc = fr.shape[0]
    a = np.asarray(fr, np.float64).reshape(Cc, -1)
    b = np.asarray(ft, np.float64).reshape(Cc, -1)
    ax = 0 if per_channel else 1
    cos = (a * b).sum(ax) / (np.maximum(np.sqrt((a * a).sum(ax)), eps) * np.maximum(np.sqrt((b * b).sum(ax)), eps))
    return f32(1.0 - cos.mean())


def triplet_loss_hnm_plus(f1, f2, margin=1.0):
    """feature/misc.py:399-435 -> (loss, chosen_case).  TripletMarginLoss(p=2, eps=1e-6, mean):
    pairwise_distance adds eps to the difference and reduces the LAST dim (W)."""
    f1 = np.asarray(f1, np.float64)
    f2 = np.asarray(f2, np.float64)
    an, ng = np.roll(f1, 1, 1), np.roll(f2, 1, 1)
    cases = [((f1 - ng) ** 2).mean(), ((f2 - an) ** 2).mean(), ((f1 - an) ** 2).mean(), ((f2 - ng) ** 2).mean()]
    c = int(np.argmin(np.asarray(cases, np.float32)))
    a, p, n = [(f1, f2, ng), (f2, f1, an), (f1, f2, an), (f2, f1, ng)][c]
    d = lambda u, v: np.sqrt(((u - v + 1e-6) ** 2).sum(-1))
    return f32(np.maximum(d(a, p) - d(a, n) + margin, 0).mean()), c
