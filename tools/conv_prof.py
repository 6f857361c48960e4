"""Role-level cycle attribution of the tcgen05 convolution kernel (dfb_debug_conv_prof) for the layer classes of DFNet at
480x640, batch 2: where does the MMA issuer wait (accumulator, input patch, weight stage), how busy is the epilogue."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dfnet_b200._lib import check, lib  # noqa: E402

dev = torch.device("cuda:0")
LAYERS = [("conv1_1", 3, 64, 3, 480, 640), ("conv1_2", 64, 64, 3, 480, 640), ("conv2_2", 128, 128, 3, 240, 320),
          ("conv3_2", 256, 256, 3, 120, 160), ("conv4_2", 512, 512, 3, 60, 80), ("conv5_2", 512, 512, 3, 30, 40),
          ("head1x1_l0", 64, 64, 1, 480, 640), ("head5x5_l0", 64, 128, 5, 480, 640)]
B = 2
print("cta_group", os.environ.get("DFB_CONV_CTA_GROUP", "2"))
for name, cin, cout, k, H, W in LAYERS:
    torch.manual_seed(0)
    w = torch.randn(cout, cin, k, k, device=dev) * 0.05
    b = torch.zeros(cout, device=dev)
    h = C.c_void_p()
    check(lib.dfb_conv_create(cin, cout, k, k, C.c_void_p(w.data_ptr()), C.c_void_p(b.data_ptr()), None, None, C.byref(h)))
    cp = (cin + 7) // 8 * 8
    x = torch.randn(B, H, W, cp, device=dev).half()
    out = torch.empty(B, H, W, cout, device=dev, dtype=torch.float16)
    for _ in range(3):
        check(lib.dfb_conv_fwd(h, C.c_void_p(x.data_ptr()), B, H, W, 1, C.c_void_p(out.data_ptr()), None, None, None))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        check(lib.dfb_conv_fwd(h, C.c_void_p(x.data_ptr()), B, H, W, 1, C.c_void_p(out.data_ptr()), None, None, None))
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    check(lib.dfb_debug_conv_prof(1, None, 0, None))
    check(lib.dfb_conv_fwd(h, C.c_void_p(x.data_ptr()), B, H, W, 1, C.c_void_p(out.data_ptr()), None, None, None))
    buf = (C.c_ulonglong * (512 * 8))()
    g = C.c_int()
    check(lib.dfb_debug_conv_prof(1, buf, 512, C.byref(g)))
    check(lib.dfb_debug_conv_prof(0, None, 0, None))
    a = np.array(buf[: g.value * 8], dtype=np.float64).reshape(g.value, 8)
    iss = a[a[:, 3] > 0]
    flops = 2.0 * k * k * cin * cout * B * H * W
    tot = iss[:, 3].mean()
    print(f"{name:11s} {us:7.1f} us {flops / us / 1e6:7.1f} TFLOP/s | issuer total {tot:9.0f} cyc: wait D_EMPTY {100 * iss[:, 0].mean() / tot:4.1f}% "
          f"A_FULL {100 * iss[:, 1].mean() / tot:4.1f}% B_FULL {100 * iss[:, 2].mean() / tot:4.1f}% | epilogue waits D_FULL "
          f"{100 * a[:, 4].mean() / max(a[:, 5].mean(), 1):4.1f}% of {a[:, 5].mean():9.0f} | loader waits A_EMPTY {a[:, 6].mean():9.0f} producer waits B_EMPTY {a[:, 7].mean():9.0f}")
    lib.dfb_conv_destroy(h)
