"""BASELINE config[4] ("Cambridge-ShopFacade-shaped 1920x1080, 64+192 samples ... rays/sec + HBM-roofline sweep"): the render
at the 1920x1080 shape over fine-sample counts and resolutions - rays/s, the fine MLP kernel's algorithmic TFLOP/s against
the measured cuBLAS rate, and the step's HBM traffic (the per-ray buffers of DESIGN section 3, read and written once)
against the measured copy bandwidth.  Prints one JSON line per point."""
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from dfnet_b200 import _lib, nerfw, ops  # noqa: E402

lib = _lib.lib
dev = torch.device("cuda:0")
pk, _ = bench.peaks()
mods = nerfw.make_synthetic_nerf(D=8, W=256)
h = ops.NerfHandle(*[m.to(dev) for m in mods])
c2w = torch.tensor(bench.pose(0), device=dev)
hist = torch.tensor(bench.HIST, device=dev)
f_c, f_f = bench.mlp_flops(256)
for (H, W, focal, Nc, Nf) in [(1080, 1920, 1674.0, 64, 64), (1080, 1920, 1674.0, 64, 128), (1080, 1920, 1674.0, 64, 192),
                              (1080, 1920, 1674.0, 64, 256), (540, 960, 837.0, 64, 192), (2160, 3840, 3348.0, 64, 192)]:
    def step():
        return h.render(Nc, Nf, True, c2w=c2w, H=H, W=W, focal=focal, near=0.0, far=20.0, hist=hist, mma="f16")
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    lib.dfb_profile_enable(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 2
    e0.record()
    for _ in range(n):
        step()
    e1.record()
    torch.cuda.synchronize()
    lib.dfb_profile_enable(0)
    cm, fm, cl, fl = C.c_double(), C.c_double(), C.c_int64(), C.c_int64()
    lib.dfb_profile_read(C.byref(cm), C.byref(fm), C.byref(cl), C.byref(fl))
    ms = e0.elapsed_time(e1) / n
    rays = H * W
    S = Nc + Nf
    fine_tf = rays * n * S * f_f / (fm.value * 1e-3) / 1e12
    # HBM bytes per ray of the step: ray record 48, ray-constant inputs 388, coarse depths / sigma / weights
    # 3 x 4 Nc written and read, per-ray bias 512 written and read, sorted depths 4 S written and read, records 32
    # ceil(S/32)+1 written and read, outputs 20
    bytes_ray = 48 * 2 + 388 * 2 + 2 * 3 * 4 * Nc + 2 * 512 + 2 * 4 * S + 2 * 32 * ((S + 31) // 32 + 1) + 20
    print(json.dumps({"H": H, "W": W, "N_samples": Nc, "N_importance": Nf, "ms_per_image": round(ms, 2),
                      "rays_per_s": round(rays / (ms * 1e-3)), "fine_mlp_tflops": round(fine_tf, 1),
                      "frac_of_bf16_sustained": round(fine_tf / pk["bf16_tflops_sustained"], 3),
                      "whole_step_tflops": round(rays * (Nc * f_c + S * f_f) / (ms * 1e-3) / 1e12, 1),
                      "hbm_bytes_per_ray": bytes_ray, "hbm_gbs": round(rays * bytes_ray / (ms * 1e-3) / 1e9, 1),
                      "frac_of_hbm": round(rays * bytes_ray / (ms * 1e-3) / 1e9 / pk["hbm_gbs"], 4)}))
