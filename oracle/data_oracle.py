"""CPU restatement (numpy) of the data-side and evaluation arithmetic around the hot path — TEST INFRASTRUCTURE ONLY
(tests/, bench cpu_baseline); the product path never imports it.

Pinned by tests/golden/data_golden.npz, generated from the reference's own statements by tests/golden/make_golden_data.py.
Paths relative to /root/reference."""
import numpy as np

f32 = np.float32


def luma_hist(img, bins=10):
    """dataset_loaders/seven_scenes.py:346-352 with dataset_loaders/utils/color.py:29-35.
    img [3,H,W] float32 in [0,1] -> [bins] float32 integer-valued percentages."""
    img = np.asarray(img, f32)
    y = (f32(0.299) * img[0] + f32(0.587) * img[1] + f32(0.114) * img[2]).astype(f32)   # rgb_to_yuv, Y only
    y = y.reshape(-1)
    y = y[(y >= 0) & (y <= 1)]                                                             # torch.histc(min=0, max=1)
    pos = ((y - f32(0)) / f32(1) * f32(bins)).astype(f32).astype(np.int64)
    pos = np.minimum(pos, bins - 1)
    hist = np.bincount(pos, minlength=bins).astype(f32)
    hist = (hist / hist.sum(dtype=f32) * f32(100)).astype(f32)
    return np.round(hist).astype(f32)                                                     # torch.round: half to even


def _area_tab(ssize, dsize):
    """cv2 computeResizeAreaTab (modules/imgproc/src/resize.cpp, opencv 4.x; a dependency of the reference, not vendored)."""
    scale = ssize / dsize
    tab = []
    for d in range(dsize):
        fsx1, fsx2 = d * scale, d * scale + scale
        cell = min(scale, ssize - fsx1)
        sx1, sx2 = int(np.ceil(fsx1)), int(np.floor(fsx2))
        sx2 = min(sx2, ssize - 1)
        sx1 = min(sx1, sx2)
        ent = []
        if sx1 - fsx1 > 1e-3:
            ent.append((sx1 - 1, f32((sx1 - fsx1) / cell)))
        for sx in range(sx1, sx2):
            ent.append((sx, f32(1.0 / cell)))
        if fsx2 - sx2 > 1e-3:
            ent.append((sx2, f32(min(min(fsx2 - sx2, 1.0), cell) / cell)))
        tab.append(ent)
    return tab


def resize_area(img, dims):
    """cv2.resize(img, dims=(w, h), interpolation=cv2.INTER_AREA), float32 HWC, downscaling
    (dataset_loaders/seven_scenes.py:328-332)."""
    img = np.asarray(img, f32)
    if img.ndim == 2:
        img = img[..., None]
    H, W, C = img.shape
    w, h = dims
    xt, yt = _area_tab(W, w), _area_tab(H, h)
    rows = np.zeros((H, w, C), f32)
    for dx, ent in enumerate(xt):
        for sx, a in ent:
            rows[:, dx] += img[:, sx] * a
    out = np.zeros((h, w, C), f32)
    for dy, ent in enumerate(yt):
        for sy, b in ent:
            out[dy] += rows[sy] * b
    return out


def matrix_to_quaternion(m):
    """pytorch3d==0.3.0 transforms.matrix_to_quaternion (requirements.txt:76; absent from /root/reference):
    w = sqrt(max(0, 1 + m00 + m11 + m22)) / 2, x / y / z alike, signs copied from the antisymmetric part."""
    m = np.asarray(m, f32)
    m00, m11, m22 = m[..., 0, 0], m[..., 1, 1], m[..., 2, 2]
    o0 = f32(0.5) * np.sqrt(np.maximum(f32(0), 1 + m00 + m11 + m22))
    x = f32(0.5) * np.sqrt(np.maximum(f32(0), 1 + m00 - m11 - m22))
    y = f32(0.5) * np.sqrt(np.maximum(f32(0), 1 - m00 + m11 - m22))
    z = f32(0.5) * np.sqrt(np.maximum(f32(0), 1 - m00 - m11 + m22))
    o1 = np.copysign(x, m[..., 2, 1] - m[..., 1, 2])
    o2 = np.copysign(y, m[..., 0, 2] - m[..., 2, 0])
    o3 = np.copysign(z, m[..., 1, 0] - m[..., 0, 1])
    return np.stack([o0, o1, o2, o3], -1).astype(f32)


def pose_error(pred, gt, use_svd=True):
    """script/feature/misc.py:49-107 (compute_error_in_q) for [n,12] row-major 3x4 poses -> [n,2] = (metres, degrees)."""
    pred = np.asarray(pred, f32).reshape(-1, 3, 4).copy()
    gt = np.asarray(gt, f32).reshape(-1, 3, 4)
    if use_svd:
        u, s, vt = np.linalg.svd(pred[:, :3, :3].astype(np.float64))
        pred[:, :3, :3] = (u @ vt).astype(f32)
    q1 = matrix_to_quaternion(gt[:, :3, :3])
    q2 = matrix_to_quaternion(pred[:, :3, :3])
    q1 = q1 / np.linalg.norm(q1, axis=-1, keepdims=True)
    q2 = q2 / np.linalg.norm(q2, axis=-1, keepdims=True)
    d = np.clip(np.abs((q1 * q2).sum(-1)), -1.0, 1.0)
    theta = 2 * np.arccos(d) * 180 / np.pi
    ex = np.linalg.norm(gt[:, :3, 3] - pred[:, :3, 3], axis=-1)
    return np.stack([ex, theta], -1).astype(f32)
