"""Data-side and evaluation rows: the oracle against the reference-generated golden (CPU), the CUDA
kernels against the golden and against cv2 itself (GPU)."""
import os
import types

import numpy as np
import pytest
import torch

from oracle import data_oracle as DO

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "data_golden.npz"))
HIST_CASES = ["noise", "dark", "edges", "ones"]
AREA_CASES = {"half": (24, 32), "third": (15, 21), "frac": (24, 31), "cambridge": (27, 48)}


@pytest.mark.parametrize("k", HIST_CASES)
def test_oracle_luma_hist_vs_reference_golden(k):
    assert np.array_equal(DO.luma_hist(G[f"hist_{k}_img"], 10), G[f"hist_{k}"])


@pytest.mark.parametrize("k", list(AREA_CASES))
def test_oracle_resize_area_vs_cv2_golden(k):
    h, w = AREA_CASES[k]
    assert np.abs(DO.resize_area(G[f"area_{k}_img"], (w, h)) - G[f"area_{k}"]).max() < 2e-6


def test_oracle_pose_error_vs_reference_golden():
    got = DO.pose_error(G["pose_pred"], G["pose_gt"])
    assert np.abs(got[:, 0] - G["pose_err"][:, 0]).max() < 1e-5
    assert np.abs(got[:, 1] - G["pose_err"][:, 1]).max() < 2e-3     # degrees; acos near 1 amplifies float32 rounding


def dev():
    return torch.device("cuda:0")


@pytest.mark.gpu
@pytest.mark.parametrize("k", HIST_CASES)
def test_luma_hist_kernel_bit_exact(k):
    from dfnet_b200.data import image_histogram
    img = torch.tensor(G[f"hist_{k}_img"], device=dev())
    got = image_histogram(img, 10)
    assert np.array_equal(got.cpu().numpy(), G[f"hist_{k}"])      # integer-valued percentages: bit-exact
    both = image_histogram(torch.stack([img, img.flip(-1)]), 10)  # batched; a mirrored image has the same histogram
    assert torch.equal(both[0], got) and torch.equal(both[1], got)


@pytest.mark.gpu
def test_luma_hist_full_size_vs_torch_statements():
    """640x480 (the loader's size): against the reference's three statements evaluated by torch on the same device
    values, and the histogram sums to ~100."""
    from dfnet_b200.data import image_histogram
    torch.manual_seed(3)
    img = torch.rand(4, 3, 480, 640, device=dev()) ** 2
    got = image_histogram(img, 10).cpu()
    for b in range(4):
        x = img[b].cpu()
        y = 0.299 * x[0] + 0.587 * x[1] + 0.114 * x[2]
        h = torch.histc(y, bins=10, min=0., max=1.)
        h = torch.round(h / h.sum() * 100)
        assert torch.equal(got[b], h), b
    assert (got.sum(-1) - 100).abs().max() <= 3


@pytest.mark.gpu
@pytest.mark.parametrize("k", list(AREA_CASES))
def test_resize_area_kernel_vs_cv2_golden(k):
    from dfnet_b200.data import resize_area
    h, w = AREA_CASES[k]
    got = resize_area(G[f"area_{k}_img"], (w, h))
    assert isinstance(got, np.ndarray) and got.shape == (h, w, 3)
    assert np.abs(got - G[f"area_{k}"]).max() < 2e-6


@pytest.mark.gpu
def test_resize_area_loader_size_vs_cv2():
    """The loader's case: 480x640 float image, df = 2 -> 240x320, against cv2 itself; and Cambridge's 480x854 -> 240x427."""
    import cv2
    from dfnet_b200.data import resize_area
    rng = np.random.RandomState(1)
    for (H, W, h, w) in ((480, 640, 240, 320), (480, 854, 240, 427), (480, 640, 120, 213)):
        im = rng.rand(H, W, 3).astype(np.float32)
        want = cv2.resize(im, (w, h), interpolation=cv2.INTER_AREA)
        got = resize_area(torch.tensor(im, device=dev()), (w, h))
        assert got.is_cuda and np.abs(got.cpu().numpy() - want).max() < 2e-6, (H, W, h, w)


@pytest.mark.gpu
def test_pose_error_kernel_and_eval_shims():
    from dfnet_b200 import misc
    pred, gt = torch.tensor(G["pose_pred"], device=dev()), torch.tensor(G["pose_gt"], device=dev())
    err, fixed = misc.pose_errors(pred, gt, return_fixed=True)
    err = err.cpu().numpy()
    assert np.abs(err[:, 0] - G["pose_err"][:, 0]).max() < 1e-5
    assert np.abs(err[:, 1] - G["pose_err"][:, 1]).max() < 2e-3
    R = fixed.reshape(-1, 3, 4)[:, :, :3].double()
    assert (R @ R.transpose(1, 2) - torch.eye(3, device=dev(), dtype=torch.float64)).abs().max() < 1e-6
    assert torch.linalg.det(R).min() > 0.99
    # identical poses -> zero error; the reference's loop interface
    z = misc.pose_errors(gt, gt).cpu().numpy()
    assert z[:, 0].max() == 0 and z[:, 1].max() < 0.1      # acos(1 - eps): float32 resolution of the angle

    class M:
        def __init__(self):
            self.i = 0

        def eval(self):
            return self

        def __call__(self, data):
            p = pred[self.i:self.i + 1]
            self.i += 1
            return None, p
    n = pred.shape[0]
    dl = [(torch.zeros(1, 3, 4, 4), gt[i:i + 1].cpu(), torch.zeros(1, 10)) for i in range(n)]
    res, vis = misc.compute_error_in_q(types.SimpleNamespace(NeRFH=True), dl, M(), dev(), np.zeros((n, 2)))
    assert np.abs(res - G["pose_err"]).max() < 2e-3 and vis["pose"].shape == (n, 3) and vis["theta"].shape == (n,)
    med, mean = misc.get_error_in_q(types.SimpleNamespace(NeRFH=True), dl, M(), n, dev())
    assert np.allclose(med, np.median(G["pose_err"], 0), atol=2e-3) and np.allclose(mean, np.mean(G["pose_err"], 0), atol=2e-3)


@pytest.mark.gpu
def test_eval_on_batch_matches_a_manual_evaluation():
    """eval_on_batch (reference feature/direct_feature_matching.py:178-213): random rays of the image rendered at the
    regressor's pose; the PSNR must equal the one computed from a full render of the same pose on the same pixels."""
    from helpers import pose_head_init_, synthetic_dfnet, synthetic_nets
    from dfnet_b200 import direct_feature_matching as dfm
    from dfnet_b200 import rendering
    H, W, focal = 48, 64, 60.0
    mods, _ = synthetic_nets(8, 64)
    c, f, ea, et = [m.to(dev()) for m in mods]
    kw = dict(network_query_fn=None, perturb=False, N_importance=24, network_fine=f, N_samples=16, network_fn=c,
              use_viewdirs=True, white_bkgd=False, raw_noise_std=0.0, embedding_a=ea, embedding_t=et, test_time=True,
              ndc=False, lindisp=False, near=0.0, far=2.5, mma="fp32")
    Fnet = pose_head_init_(synthetic_dfnet("DFNet", seed=0)).to(dev()).eval()
    args = types.SimpleNamespace(DFNet=True, preprocess_ImgNet=False, svd_reg=True, chunk=32768, batch_size=1, N_rand=500, NeRFH=True)
    rng = np.random.RandomState(2)
    data = torch.tensor(rng.rand(1, 3, H, W).astype(np.float32))
    pose = torch.tensor([[1, 0, 0, 0.1, 0, 1, 0, -0.05, 0, 0, 1, 2.0]], dtype=torch.float32)
    hist = torch.tensor([[5., 10, 20, 30, 15, 10, 5, 3, 1, 1]])
    world = dict(pose_scale=0.5, pose_scale2=1.0, move_all_cam_vec=[0.0, 0.0, 0.05])
    torch.manual_seed(11)
    loss, psnr = dfm.eval_on_batch(args, data, Fnet, None, pose, hist, (H, W, focal), False, dev(), world, **kw)
    # manual: same pose, full image, same random pixels
    with torch.no_grad():
        _, pose_ = dfm.inference_pose_regression(args, data.to(dev()), dev(), Fnet)
        pn = dfm.fix_coord_supp(args, pose_.clone(), world)
        rgb, _, _, _ = rendering.render(H, W, focal, c2w=pn[0, :3, :4], img_idx=hist.to(dev()), **kw)
        torch.manual_seed(11)
        sel = torch.randperm(H * W)[:500].to(dev())
        tgt = data.to(dev())[0].permute(1, 2, 0).reshape(-1, 3)[sel]
        mse = ((rgb.reshape(-1, 3)[sel] - tgt) ** 2).mean()
        want_psnr = float(-10 * torch.log10(mse))
        want_loss = float(((pose_.reshape(1, 12) - pose.to(dev())) ** 2).mean())
    assert abs(float(psnr) - want_psnr) < 1e-3 * abs(want_psnr) and abs(float(loss[0]) - want_loss) < 1e-5 * max(1.0, want_loss)
