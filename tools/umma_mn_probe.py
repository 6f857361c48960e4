"""Bring-up probe: MN-major tcgen05 operands, both LBO/SBO conventions, f16/bf16 and mixed operand formats."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dfnet_b200 import ops  # noqa: E402

rng = np.random.default_rng(0)
for K, N in ((64, 128), (128, 64)):
    A = rng.standard_normal((K, 128)).astype(np.float32)
    B = rng.standard_normal((K, N)).astype(np.float32)
    At, Bt = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    for fa, fb in ((0, 0), (1, 1), (0, 1), (1, 0)):
        ra = At.half().float() if fa == 0 else At.bfloat16().float()
        rb = Bt.half().float() if fb == 0 else Bt.bfloat16().float()
        want = (ra.double().t() @ rb.double()).cpu().numpy()
        for variant in (0, 1):
            D = torch.zeros(128, N, device="cuda")
            ops.check(ops.lib.dfb_debug_umma_gemm_mn(C.c_void_p(At.data_ptr()), C.c_void_p(Bt.data_ptr()), N, K, fa, fb,
                                                     variant, C.c_void_p(D.data_ptr()), None))
            torch.cuda.synchronize()
            print(f"K={K} N={N} fmt_a={fa} fmt_b={fb} variant={variant} max|err|={np.abs(D.cpu().numpy() - want).max():.3e}",
                  flush=True)
