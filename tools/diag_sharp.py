"""Diagnostic: where do the kernels and the oracle part ways on the trained-like (sharp) field of
tests/test_headline_gpu.py?  Compares coarse weights, sample indices, merged depths, fine raw and the composite for the
fp32 and fp16 paths, and prints the worst rays."""
import os
import sys

import numpy as np
import torch

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root)
sys.path.insert(0, os.path.join(root, "tests"))
from helpers import fit_synthetic_scene, np_params  # noqa: E402
from oracle import nerf_oracle as O  # noqa: E402  (diagnostic tool, not the product path)
from dfnet_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
HIST = np.array([5, 10, 20, 30, 15, 10, 5, 3, 1, 1], np.float32)
C2W = np.array([[0.9962, -0.0872, 0.0, 0.0], [0.0872, 0.9962, 0.0, 0.0], [0.0, 0.0, 1.0, 1.0]], np.float32)
steps = int(os.environ.get("FIT_STEPS", 400))
c, f, ea, et = fit_synthetic_scene(steps=steps, batch=16384, device=dev)
nets = dict(coarse=np_params(c), fine=np_params(f), emb_a=ea.weight.detach().cpu().numpy(), emb_t=et.weight.detach().cpu().numpy(),
            D=8, skips=(4,), beta_min=0.1)
H, W, focal, near, far, Nc, Nf = 480, 640, 585.0, 0.0, 2.5, 64, 128
n = 1500
sel = np.linspace(0, H * W - 1, n).astype(np.int64)
o, d = O.get_rays(H, W, focal, C2W)
rec_np = O.make_ray_records(o.reshape(-1, 3)[sel], d.reshape(-1, 3)[sel], near, far, HIST[None])
O.set_linear_backend("torch")
want = O.render_rays(rec_np, nets, Nc, Nf, test_time=True, retraw=True, return_internals=True)
I = want["_internals"]
h = ops.NerfHandle(c, f, ea, et)
rec = torch.tensor(rec_np, device=dev)
for mma in ("fp32", "f16s", "f16"):
    g = h.render(Nc, Nf, True, rays=rec, mma=mma, want=("z_vals", "raw", "weights_coarse", "inds", "z_samples"))
    g = {k: v.cpu().numpy() for k, v in g.items()}
    print(f"==== {mma}")
    print("coarse weights: max abs diff", np.abs(g["weights_coarse"] - I["weights_coarse"]).max())
    mism = (g["inds"] != I["inds"])
    print("inds mismatches:", mism.sum(), "of", mism.size, "rays affected", mism.any(1).sum())
    dz = np.abs(g["z_vals"] - I["z_vals"])
    print("z_vals: max abs diff", dz.max(), "rays with diff > 1e-5:", (dz.max(1) > 1e-5).sum())
    draw = np.abs(g["raw"] - want["raw"])
    print("raw max abs diff per channel", draw.reshape(-1, 9).max(0))
    r = np.abs(g["rgb"] - want["rgb_map"]) / np.maximum(np.abs(want["rgb_map"]), 1e-3)
    pr = r.max(1)
    print("rgb rel err: mean %.2e p99 %.2e max %.2e frac>1e-3 %.3f" % (pr.mean(), np.percentile(pr, 99), pr.max(), (pr > 1e-3).mean()))
    # composite of the KERNEL's own raw / z with the oracle compositor: isolates the MLP from the compositing
    comp = O.raw2outputs_nerfw(g["raw"], g["z_vals"], 0.0, True, 0.1, test_time=True, typ="fine")
    rc = np.abs(g["rgb"] - comp["rgb"]) / np.maximum(np.abs(comp["rgb"]), 1e-3)
    print("kernel composite vs oracle composite of the kernel's raw: max rel", rc.max())
    # oracle MLP on the kernel's depths: isolates the sampler
    for i in np.argsort(-pr)[:3]:
        print(" worst ray", i, "err", pr[i], "rgb", g["rgb"][i], want["rgb_map"][i], "z diff", dz[i].max(), "ind mism", mism[i].sum(),
              "raw diff", draw[i].max(), "max sigma", want["raw"][i, :, 3].max())
