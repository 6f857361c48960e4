"""Host cost of one asynchronous launch through the Python -> ctypes -> C ABI path (what bounds the launch-bound steps:
NeRF-Hist training runs ~480 device operations per step): wall time per call of a small 1x1 convolution and of a weight
gradient, queue kept short so that the host, not the GPU, is measured."""
import ctypes as C
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dfnet_b200._lib import check, lib  # noqa: E402
from dfnet_b200 import nerf_train  # noqa: E402

dev = torch.device("cuda:0")
L = nerf_train._Layer(128, 128, dev)
L.load(torch.randn(128, 128, device=dev) * 0.05, torch.zeros(128, device=dev))
Pp = 8 * 64
x = torch.randn(Pp, 128, device=dev).half()
o = torch.empty(Pp, 128, device=dev, dtype=torch.float16)
o2 = torch.empty(Pp, 128, device=dev, dtype=torch.bfloat16)
g = torch.randn(Pp, 128, device=dev).bfloat16()
dW, dB = torch.zeros(128, 128, device=dev), torch.zeros(128, device=dev)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)


def bench(name, fn, n=600):
    """Host time to ENQUEUE n calls (the queue holds them all, so the GPU does not throttle the host) and the time until
    the GPU has drained them; the larger of the two bounds a launch-bound step."""
    for _ in range(50):
        fn()
    torch.cuda.synchronize()
    host, total = [], []
    for _ in range(5):
        t0 = time.perf_counter()
        for _i in range(n):
            fn()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        host.append((t1 - t0) / n), total.append((t2 - t0) / n)
    print(f"{name:48s} host {1e6 * min(host):6.2f} us   drained {1e6 * min(total):6.2f} us per call")


bench("nerf_train._conv (python wrapper, bf16 twin)", lambda: nerf_train._conv(L.fwd, x, Pp // 8, 1, out=o, out_bf=o2))
bench("lib.dfb_conv_fwd_ex2 (ctypes, args prebuilt)", (lambda a=(L.fwd, C.c_void_p(x.data_ptr()), 1, Pp // 8, 8, 1,
      C.c_void_p(o.data_ptr()), None, None, None, None, C.c_void_p(o2.data_ptr()), st): lib.dfb_conv_fwd_ex2(*a)))
bench("lib.dfb_conv_wgrad_acc (ctypes, args prebuilt)", (lambda a=(C.c_void_p(g.data_ptr()), C.c_void_p(o2.data_ptr()), 1,
      Pp // 8, 8, 128, 128, 128, 1, 1, C.c_void_p(dW.data_ptr()), C.c_void_p(dB.data_ptr()), st): lib.dfb_conv_wgrad_acc(*a)))
bench("torch.empty(1, device)", lambda: torch.empty(1, device=dev))
bench("torch add (tiny)", lambda: o.add_(1))

# ---- fixed device-side cost of one convolution launch: the same 128 -> 128 1x1 layer over growing pixel counts, timed
# back to back on the device (events), for both CTA-group variants
for cg, pdl in (("2", "1"), ("2", "0"), ("1", "1"), ("1", "0")):
    os.environ["DFB_CONV_CTA_GROUP"], os.environ["DFB_PDL"] = cg, pdl
    for rows in (64, 1024, 12288):
        xx = torch.randn(rows * 8, 128, device=dev).half()
        oo = torch.empty(rows * 8, 128, device=dev, dtype=torch.float16)
        a = (L.fwd, C.c_void_p(xx.data_ptr()), 1, rows, 8, 1, C.c_void_p(oo.data_ptr()), None, None, None, None, st)
        for _ in range(20):
            lib.dfb_conv_fwd_ex(*a)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(200):
            lib.dfb_conv_fwd_ex(*a)
        e1.record()
        torch.cuda.synchronize()
        print(f"cta_group {cg} pdl {pdl}: {rows * 8:6d} pixels  {1e3 * e0.elapsed_time(e1) / 200:6.2f} us per launch (device, back to back)")
del os.environ["DFB_CONV_CTA_GROUP"], os.environ["DFB_PDL"]
