"""BASELINE config[2]: DFNet siamese forward + level-0 cosine feature loss on one 640x480 pair.
Prints one JSON line (pairs/s, conv TFLOP/s against 325.3 GFLOP per image, loss HBM GB/s)."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from helpers import synthetic_dfnet  # noqa: E402
from dfnet_b200.dfnet import feature_loss  # noqa: E402
from dfnet_b200._lib import lib  # noqa: E402

dev = torch.device("cuda:0")
net = synthetic_dfnet("DFNet").to(dev)
torch.manual_seed(0)
x = torch.rand(2, 3, 480, 640, device=dev)
steps, warm = int(os.environ.get("STEPS", 10)), 3


@torch.no_grad()   # forward-only configuration: the inference path (ping-pong activation buffers, no tape)
def step():
    feats, _ = net(x, return_feature=True, isSingleStream=False, return_pose=False, upsampleH=480, upsampleW=640)
    return feature_loss(feats[1][0, 0], feats[0][0, 0])


for _ in range(warm):
    step()
torch.cuda.synchronize()
l0 = lib.dfb_launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    loss = step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
# loss alone
f = torch.randn(128, 480 * 640, device=dev)
g = torch.randn(128, 480 * 640, device=dev)
for _ in range(3):
    feature_loss(f, g)
torch.cuda.synchronize()
e0.record()
for _ in range(20):
    feature_loss(f, g)
e1.record()
torch.cuda.synchronize()
lms = e0.elapsed_time(e1) / 20
# encoder-only (no pose, no upsampling of levels 1/2 excluded is not separable here): report whole step
peaks = json.load(open("MEASURED_PEAKS.json")) if os.path.exists("MEASURED_PEAKS.json") else {"bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0}
tf = 2 * 325.3e9 / (ms * 1e-3) / 1e12
print(json.dumps({"metric": "DFNet pairs/sec (640x480, siamese forward + level-0 cosine loss)", "value": 1e3 / ms, "unit": "pairs/s",
                  "ms_per_pair": ms, "gpu_launches_per_pair": (lib.dfb_launch_count() - l0 - 0) / steps,
                  "conv_tflops_whole_step": tf, "frac_of_bf16_sustained": tf / peaks["bf16_tflops_sustained"],
                  "cosine_loss_ms": lms, "cosine_loss_GBps": 2 * 128 * 480 * 640 * 4 / (lms * 1e-3) / 1e9,
                  "cosine_loss_frac_of_hbm": 2 * 128 * 480 * 640 * 4 / (lms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                  "loss": float(loss)}))
