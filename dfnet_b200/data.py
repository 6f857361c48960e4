"""Data-side pieces of the loaders on the GPU: the luma histogram that indexes the NeRF-Hist
embeddings and the INTER_AREA downscale of the input images.  Same arithmetic as the reference
(dataset_loaders/seven_scenes.py:328-352, dataset_loaders/utils/color.py:29-35); on-disk formats are untouched."""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import check, lib, raw_stream


def _p(t):
    return C.c_void_p(t.data_ptr())


def _stream():
    return raw_stream()


def image_histogram(img, hist_bin=10):
    """img [3,H,W] or [B,3,H,W] float in [0,1] (CUDA) -> hist [hist_bin] / [B,hist_bin]: what the loaders return as the
    third batch element with ret_hist (seven_scenes.py:346-352): rgb_to_yuv, torch.histc of Y over [0,1], percentages,
    torch.round."""
    if not (isinstance(img, torch.Tensor) and img.is_cuda):
        raise _lib.DfbError("image_histogram input must be a CUDA tensor: the dfnet_b200 path has no CPU fallback")
    single = img.dim() == 3
    x = (img[None] if single else img).detach().float().contiguous()
    B, ch, H, W = x.shape
    if ch != 3:
        raise ValueError(f"Input size must have a shape of (*, 3, H, W). Got {tuple(img.shape)}")
    out = torch.empty(B, hist_bin, device=x.device)
    ws = torch.empty(B * hist_bin, dtype=torch.int32, device=x.device)
    check(lib.dfb_luma_hist(_p(x), B, H, W, int(hist_bin), _p(out), _p(ws), ws.numel() * 4, _stream()))
    return out[0] if single else out


def resize_area(img, dims):
    """cv2.resize(img, dims, interpolation=cv2.INTER_AREA) for an HWC float image and dims = (W, H), downscaling
    (seven_scenes.py:328-332).  Accepts a CUDA tensor (returns a CUDA tensor) or a numpy array (one upload, one
    download: the loader's contract is a numpy image)."""
    w, h = int(dims[0]), int(dims[1])
    as_np = isinstance(img, np.ndarray)
    if not torch.cuda.is_available():
        raise _lib.DfbError("no CUDA device: the dfnet_b200 path has no CPU fallback")
    x = torch.as_tensor(img, dtype=torch.float32)
    x = x.cuda() if not x.is_cuda else x
    squeeze = x.dim() == 2
    x = (x[..., None] if squeeze else x).contiguous()
    H, W, Cc = x.shape
    out = torch.empty(h, w, Cc, device=x.device)
    check(lib.dfb_resize_area(_p(x), H, W, Cc, h, w, _p(out), _stream()))
    out = out[..., 0] if squeeze else out
    return out.cpu().numpy() if as_np else out
