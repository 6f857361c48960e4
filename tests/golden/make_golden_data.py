"""Golden vectors for the data-side / evaluation rows from the reference's own statements.

    python tests/golden/make_golden_data.py   ->  tests/golden/data_golden.npz

* histogram: the three statements of dataset_loaders/seven_scenes.py:346-352 executed verbatim on top of the reference's
  rgb_to_yuv (dataset_loaders/utils/color.py);
* INTER_AREA: cv2.resize itself (the reference's call, seven_scenes.py:331), float32 HWC, integer and fractional factors;
* pose error: the reference's compute_error_in_q (script/feature/misc.py:49-107) run on a stand-in model / loader, with
  pytorch3d.transforms.matrix_to_quaternion (pytorch3d==0.3.0, not installed) supplied by its published formula."""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from baseline import ref_shims  # noqa: E402

ref_shims.activate()
import cv2  # noqa: E402
import torch  # noqa: E402

sys.path.insert(0, "/root/reference")
from dataset_loaders.utils.color import rgb_to_yuv  # noqa: E402
from oracle import data_oracle as DO  # noqa: E402

G = {}
rng = np.random.RandomState(7)
# ---- histograms -----------------------------------------------------------------------------------------------------
imgs = {
    "noise": rng.rand(3, 60, 80).astype(np.float32),
    "dark": (rng.rand(3, 48, 64) ** 3).astype(np.float32),
    "edges": np.clip(np.round(rng.rand(3, 37, 53) * 10) / 10, 0, 1).astype(np.float32),   # many values ON bin boundaries
    "ones": np.ones((3, 8, 8), np.float32),
}
for k, im in imgs.items():
    img = torch.from_numpy(im)
    yuv = rgb_to_yuv(img)
    y_img = yuv[0]
    hist = torch.histc(y_img, bins=10, min=0., max=1.)
    hist = hist / (hist.sum()) * 100
    hist = torch.round(hist)
    G[f"hist_{k}_img"], G[f"hist_{k}"] = im, hist.numpy()
    assert np.array_equal(DO.luma_hist(im, 10), hist.numpy()), k
# ---- INTER_AREA -------------------------------------------------------------------------------------------------------
for k, (H, W, h, w) in {"half": (48, 64, 24, 32), "third": (45, 63, 15, 21), "frac": (50, 70, 24, 31), "cambridge": (54, 96, 27, 48)}.items():
    im = rng.rand(H, W, 3).astype(np.float32)
    out = cv2.resize(im, (w, h), interpolation=cv2.INTER_AREA)
    G[f"area_{k}_img"], G[f"area_{k}"] = im, out
    assert np.abs(DO.resize_area(im, (w, h)) - out).max() < 2e-6, k
# ---- pose error -----------------------------------------------------------------------------------------------------
import pytorch3d.transforms as T3  # noqa: E402  (the stub module installed by ref_shims)
T3.matrix_to_quaternion = lambda m: torch.from_numpy(DO.matrix_to_quaternion(m.numpy()))
import feature.misc as RM  # noqa: E402
RM.transforms = T3
n = 12
gt = np.zeros((n, 3, 4), np.float32)
pred = np.zeros((n, 3, 4), np.float32)
for i in range(n):
    ax = rng.randn(3)
    ax /= np.linalg.norm(ax)
    ang = rng.rand() * 2.5
    K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
    R = np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K
    gt[i, :, :3], gt[i, :, 3] = R, rng.randn(3)
    pred[i, :, :3] = R @ (np.eye(3) + 0.05 * rng.randn(3, 3))      # not orthogonal: the SVD step matters
    pred[i, :, 3] = gt[i, :, 3] + 0.1 * rng.randn(3)


class _Model:
    def __init__(self):
        self.i = 0

    def __call__(self, data):
        p = torch.from_numpy(pred[self.i].reshape(1, 12))
        self.i += 1
        return None, p


args = types.SimpleNamespace(NeRFH=True)
dl = [(torch.zeros(1, 3, 4, 4), torch.from_numpy(gt[i].reshape(1, 12)), torch.zeros(1, 10)) for i in range(n)]
res, _ = RM.compute_error_in_q(args, dl, _Model(), torch.device("cpu"), np.zeros((n, 2)), batch_size=1)
G["pose_pred"], G["pose_gt"], G["pose_err"] = pred.reshape(n, 12), gt.reshape(n, 12), res.astype(np.float32)
assert np.abs(DO.pose_error(pred.reshape(n, 12), gt.reshape(n, 12)) - res).max() < 2e-3, np.abs(DO.pose_error(pred.reshape(n, 12), gt.reshape(n, 12)) - res).max()
out = os.path.join(HERE, "data_golden.npz")
np.savez_compressed(out, **G)
print("wrote", out, os.path.getsize(out) / 1e3, "KB")
