"""Role-level cycle attribution of the tcgen05 render backward (k_mlp_tc_bwd, saved-mask program).  Needs the -DDFB_TC_PROF
build: make -C dfnet_b200/csrc prof && DFB_LIB_PATH=dfnet_b200/csrc/build_prof/libdfnet_b200_prof.so python tools/tcb_prof.py"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dfnet_b200 import nerfw, ops  # noqa: E402
from dfnet_b200.rendering import render  # noqa: E402

dev = torch.device("cuda:0")
mods = [m.to(dev) for m in nerfw.make_synthetic_nerf(D=8, W=256, fine=True)]
for m in mods:
    for p in m.parameters():
        p.requires_grad_(False)
c, f, ea, et = mods
kw = dict(network_query_fn=None, perturb=0.0, N_importance=128, network_fine=f, N_samples=64, network_fn=c, use_viewdirs=True,
          white_bkgd=False, raw_noise_std=0.0, embedding_a=ea, embedding_t=et, test_time=True, ndc=False, lindisp=False,
          near=0.0, far=2.5, mma="f16")
c2w = torch.tensor([[1., 0, 0, 0.1], [0, 1, 0, -0.05], [0, 0, 1, 2.0]], device=dev, requires_grad=True)
hist = torch.tensor([[5., 10, 20, 30, 15, 10, 5, 3, 1, 1]], device=dev)
H, W = 128, 128   # 16 384 rays = one backward launch
for _ in range(2):
    rgb = render(H, W, 146.0, chunk=32768, c2w=c2w, img_idx=hist, **kw)[0]
    (rgb * torch.rand_like(rgb)).mean().backward()
torch.cuda.synchronize()
big = np.zeros((256, 64), np.uint64)
rc = ops.lib.dfb_debug_tcb_prof(big.ctypes.data_as(C.c_void_p), 256)
if rc != 0:
    print("no cycle counters in this build (make -C dfnet_b200/csrc prof); ran one forward + backward of", H * W, "rays")
    sys.exit(0)
b = big[:148].astype(np.float64)
names = {0: ("producer", ["wait W_EMPTY"]), 4: ("issuer", ["wait W_FULL", "wait A_READY/PASS_DONE", "wait PE_READY"]),
         8: ("epi slot0", ["wait D_FULL"]), 12: ("epi slot1", ["wait D_FULL"])}
for base, (nm, labels) in names.items():
    tot = b[:, base + 3]
    act = tot > 0
    line = f"{nm:10s} total {tot[act].mean():12.0f}"
    for i, lb in enumerate(labels):
        line += f" | {lb} {b[act, base + i].mean():12.0f} ({100 * b[act, base + i].mean() / tot[act].mean():.1f}%)"
    print(line)
steps = b[:, 16:16 + 26].mean(0)
tot = b[:, 11].mean()
n_pass = (H * W * 192 / 128 / 2) / 148
print("passes per CTA %.1f; epilogue busy cycles per step and pass (slot 0):" % n_pass)
print(" ".join(f"{i}:{v / n_pass:.0f}" for i, v in enumerate(steps) if v > 0))
print("sum per pass %.0f cycles; kernel total per pass %.0f" % (steps.sum() / n_pass, tot / n_pass))
