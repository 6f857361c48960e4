"""Host-side mirror of the reference's `feature/dfnet.py` (DFNet / DFNet_s).

Same module tree and state_dict keys as the reference (`encoder.{0..28}.*`,
`adaptation_layers.adapt_layer_{i}.{0,2,3}.*`, `fc_pose.*`, reference feature/dfnet.py:74-172), so
its checkpoints load unchanged; the constructor never touches the network (the reference
downloads ImageNet VGG-16 weights, feature/dfnet.py:90).  forward() runs on the sm_100a kernels
(implicit-GEMM tcgen05 convolutions, fused bias/ReLU, eval-mode BatchNorm folded into the 5x5
convs); there is no eager fallback.
"""
import ctypes as C
import weakref

import torch
import torch.nn as nn

from . import _lib
from ._lib import check, lib

_VGG16_CFG = [64, 64, "M", 128, 128, "M", 256, 256, 256, "M", 512, 512, 512, "M", 512, 512, 512, "M"]
_TAP_CHANNELS = {"conv1_2": 64, "conv3_3": 256, "conv5_3": 512}


def _vgg16_features():
    layers, c = [], 3
    for v in _VGG16_CFG:
        if v == "M":
            layers.append(nn.MaxPool2d(kernel_size=2, stride=2))
        else:
            layers += [nn.Conv2d(c, v, kernel_size=3, padding=1), nn.ReLU(inplace=True)]
            c = v
    return nn.Sequential(*layers)


class AdaptLayers(nn.Module):
    """Adaptation heads (reference feature/dfnet.py:42-72): Conv1x1 -> ReLU -> Conv5x5 -> BatchNorm2d."""

    def __init__(self, hypercolumn_layers, output_dim=128):
        super().__init__()
        for i, name in enumerate(hypercolumn_layers):
            self.add_module(f"adapt_layer_{i}", nn.Sequential(
                nn.Conv2d(_TAP_CHANNELS[name], 64, kernel_size=1, stride=1, padding=0), nn.ReLU(),
                nn.Conv2d(64, output_dim, kernel_size=5, stride=1, padding=2), nn.BatchNorm2d(output_dim)))


class _DfnetHandle:
    def __init__(self, module):
        h = C.c_void_p()
        check(lib.dfb_dfnet_create(len(module.hypercolumn_layers), C.byref(h)))
        self._h = h
        self._fin = weakref.finalize(self, lib.dfb_dfnet_destroy, h)
        self._versions = None
        self._ws = None
        self.n_levels = len(module.hypercolumn_layers)

    def refresh(self, module):
        sd = module.state_dict()
        v = [(t.data_ptr(), t._version) for t in sd.values()]
        if v == self._versions:
            return
        names = [f"encoder.{i}" for i, m in enumerate(module.encoder) if isinstance(m, nn.Conv2d)]
        ts = []
        for n in names:
            ts += [sd[n + ".weight"], sd[n + ".bias"]]
        for l in range(len(module.hypercolumn_layers)):
            p = f"adaptation_layers.adapt_layer_{l}."
            ts += [sd[p + "0.weight"], sd[p + "0.bias"], sd[p + "2.weight"], sd[p + "2.bias"], sd[p + "3.weight"],
                   sd[p + "3.bias"], sd[p + "3.running_mean"], sd[p + "3.running_var"]]
        ts += [sd["fc_pose.weight"], sd["fc_pose.bias"]]
        ts = [t.detach().float().contiguous() for t in ts]
        ptrs = (C.c_void_p * len(ts))(*[t.data_ptr() for t in ts])
        numel = (C.c_int64 * len(ts))(*[t.numel() for t in ts])
        eps = module.adaptation_layers.adapt_layer_0[3].eps
        check(lib.dfb_dfnet_load(self._h, ptrs, numel, len(ts), eps))
        self._versions = v

    def forward(self, x, return_feature, single, return_pose, upH, upW):
        if not x.is_cuda:
            raise _lib.DfbError("DFNet input must be a CUDA tensor: the dfnet_b200 hot path has no CPU fallback")
        x = x.detach().float().contiguous()
        B, _, H, W = x.shape
        dev = x.device
        need = C.c_size_t()
        check(lib.dfb_dfnet_workspace_bytes(self._h, B, H, W, upH, upW, C.byref(need)))
        if self._ws is None or self._ws.numel() < need.value or self._ws.device != dev:
            self._ws = torch.empty(need.value, dtype=torch.uint8, device=dev)
        flags = (1 if return_feature else 0) | (2 if single else 0) | (4 if return_pose else 0)
        ft = fr = pose = None
        if return_feature:
            Bs = B if single else B // 2
            ft = torch.empty(self.n_levels, Bs, 128, upH, upW, device=dev)
            fr = None if single else torch.empty(self.n_levels, Bs, 128, upH, upW, device=dev)
        if return_pose:
            pose = torch.empty(B, 12, device=dev)

        def p(t):
            return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)
        check(lib.dfb_dfnet_fwd(self._h, p(x), B, H, W, flags, upH, upW, p(ft), p(fr), p(pose), p(self._ws),
                                self._ws.numel(), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        return ft, fr, pose


class DFNet(nn.Module):
    """DFNet (reference feature/dfnet.py:74-172): VGG-16 encoder, three hyper-column adaptation heads
    (pre-ReLU conv1_2 / conv3_3 / conv5_3), pose head AdaptiveAvgPool2d(1) -> Linear(512, feat_dim)."""
    hypercolumn_layers = ["conv1_2", "conv3_3", "conv5_3"]
    mean = [0.485, 0.456, 0.406]
    std = [0.229, 0.224, 0.225]

    def __init__(self, feat_dim=12, places365_model_path=""):
        super().__init__()
        if feat_dim != 12:
            raise NotImplementedError("the pose head is a 3x4 matrix (feat_dim=12) everywhere in the reference")
        self.encoder = _vgg16_features()
        self.scales = [1, 4, 16][: len(self.hypercolumn_layers)]
        self.adaptation_layers = AdaptLayers(self.hypercolumn_layers, 128)
        self.avgpool = nn.AdaptiveAvgPool2d(1)
        self.fc_pose = nn.Linear(512, feat_dim)
        self._handle = None

    def forward(self, x, return_feature=False, isSingleStream=False, return_pose=True, upsampleH=240, upsampleW=427):
        """Reference feature/dfnet.py:106-172 -> (feature_maps, predict): feature_maps is None,
        [stack [L,B,128,H,W]] (single stream) or [target_stack, render_stack] (siamese)."""
        bn = self.adaptation_layers.adapt_layer_0[3]
        if return_feature and bn.training:
            raise NotImplementedError("train-mode BatchNorm (batch statistics, run_feature.py without freezeBN) "
                                      "is not on the B200 hot path yet; call .eval() / freeze_bn_layer_train")
        if self._handle is None:
            self._handle = _DfnetHandle(self)
        self._handle.refresh(self)
        ft, fr, pose = self._handle.forward(x, return_feature, isSingleStream, return_pose, int(upsampleH), int(upsampleW))
        if not return_feature:
            feature_maps = None
        elif isSingleStream:
            feature_maps = [ft]
        else:
            feature_maps = [ft, fr]
        return feature_maps, pose


class DFNet_s(DFNet):
    """DFNet_s (reference feature/dfnet.py:174-273): only the conv1_2 hyper-column."""
    hypercolumn_layers = ["conv1_2"]


def feature_loss(feature_rgb, feature_target, img_in=True, per_channel=False):
    """Cosine feature loss (reference feature/direct_feature_matching.py:114-136).  With the default
    per_channel=False the cosine runs over the pixels of each channel (the reference's naming is
    inverted w.r.t. its behaviour); per_channel=True gives the per-pixel cosine."""
    if not feature_rgb.is_cuda:
        raise _lib.DfbError("feature_loss inputs must be CUDA tensors")
    fr = feature_rgb.detach().float().contiguous()
    ft = feature_target.detach().float().contiguous()
    Cc = fr.shape[0]
    HW = fr.numel() // Cc
    loss = torch.empty((), device=fr.device)
    ws = torch.empty(max(Cc * 64 * 3, (HW + 255) // 256), device=fr.device)
    check(lib.dfb_cosine_loss(C.c_void_p(fr.data_ptr()), C.c_void_p(ft.data_ptr()), Cc, HW, int(bool(per_channel)), 1e-6,
                              C.c_void_p(loss.data_ptr()), C.c_void_p(ws.data_ptr()), ws.numel() * 4,
                              C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return loss


def preprocess_features_for_loss(feature):
    """[L,B,C,H,W] -> [B,L*C,H,W] (reference feature/direct_feature_matching.py:41-50)."""
    feature = feature.permute(1, 0, 2, 3, 4)
    B, L, Cc, H, W = feature.size()
    return feature.reshape((B, L * Cc, H, W))
