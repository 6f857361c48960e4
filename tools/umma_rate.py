"""Tensor-pipe rate probe: cycles per tcgen05.mma (M=128, K=16) for several N and grid sizes."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from dfnet_b200._lib import lib, check  # noqa: E402

torch.cuda.init()
torch.zeros(1, device="cuda")
for grid in (1, 148):
    for n in (64, 128, 256):
        v = C.c_double()
        check(lib.dfb_debug_umma_rate(2000, n, grid, C.byref(v)))
        print(f"grid={grid:4d} N={n:3d}: {v.value:7.1f} cycles/MMA  (nominal {n // 2})")
for grid in (1, 148):
    for nw in (1, 2, 4):
        v = C.c_double()
        check(lib.dfb_debug_tmem_rate(2000, nw, grid, C.byref(v)))
        print(f"grid={grid:4d} warps={nw}: {v.value:7.1f} cycles per tcgen05.ld.32x32b.x32 (4 KB) per warp -> {nw * 4096 / v.value:6.1f} B/clk/SM")
# TMEM reads while the tensor pipe is busy: 2000 x 8 loads per warp take ~450 k cycles alone; 4000 x 16 MMAs take ~8 M cycles
for nw in (1, 4):
    v = (C.c_double * 2)()
    check(lib.dfb_debug_tmem_rate_mma(20000, nw, 4000, 148, v))
    print(f"under MMA load, warps={nw}: {v[0]:7.1f} cycles per tcgen05.ld.32x32b.x32 per warp, {v[1]:6.1f} cycles per MMA (N=256)")
