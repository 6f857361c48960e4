"""Role-level cycle attribution of the tcgen05 MLP kernel (needs the -DDFB_TC_PROF build:
make -C dfnet_b200/csrc prof && DFB_LIB_PATH=dfnet_b200/csrc/build_prof/libdfnet_b200_prof.so python tools/tc_prof.py)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dfnet_b200 import nerfw, ops  # noqa: E402

dev = torch.device("cuda:0")
WIDTH = int(sys.argv[1]) if len(sys.argv) > 1 else 256
NF = int(sys.argv[2]) if len(sys.argv) > 2 else 128
mods = nerfw.make_synthetic_nerf(D=8, W=WIDTH)
h = ops.NerfHandle(*[m.to(dev) for m in mods])
c2w = torch.tensor([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 1.0]], device=dev)
hist = torch.tensor([5, 10, 20, 30, 15, 10, 5, 3, 1, 1.0], device=dev)
for which in (0, 1):
    os.environ["DFB_TC_PROF_WHICH"] = str(which)
    for _ in range(2):
        h.render(64, NF, True, c2w=c2w, H=256, W=256, focal=300.0, near=0.0, far=2.5, hist=hist, mma="f16")
    torch.cuda.synchronize()
    big = np.zeros((512, 16), np.uint64)
    rc = ops.lib.dfb_debug_tc_prof(big.ctypes.data_as(C.c_void_p), 512)
    buf = big[:148]
    assert rc == 0, rc
    b = buf.astype(np.float64)
    names = {0: ("producer", ["wait W_EMPTY"]), 4: ("mma", ["wait W_FULL", "wait W_FULLP", "wait A_READY/PE"]),
             8: ("epi wg0", ["wait D_FULL", "fences", "arrive"]), 12: ("epi wg1", ["wait D_FULL", "fences", "arrive"])}
    print(f"--- network {which} ({'coarse' if which == 0 else 'fine'}) : mean over CTAs, cycles")
    for base, (nm, labels) in names.items():
        tot = b[:, base + 3]
        act = tot > 0
        if not act.any():
            continue
        line = f"{nm:10s} total {tot[act].mean():12.0f}"
        for i, lb in enumerate(labels):
            line += f" | {lb} {b[act, base + i].mean():12.0f} ({100 * b[act, base + i].mean() / tot[act].mean():.1f}%)"
        print(line)
    steps = big[256:256 + 148].astype(np.float64).mean(0)
    tot = b[:, 7].mean()
    print("issuer wait for A_READY by step (% of kernel):", " ".join(f"{i}:{100 * v / tot:.1f}" for i, v in enumerate(steps) if v > 0))
