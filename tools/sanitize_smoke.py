"""Small end-to-end run of every kernel family for compute-sanitizer (memcheck)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from dfnet_b200 import nerfw, ops  # noqa: E402
from dfnet_b200.dfnet import feature_loss  # noqa: E402
from dfnet_b200.misc import mse, triplet_loss_hard_negative_mining_plus, upsample_bicubic  # noqa: E402
from helpers import synthetic_dfnet  # noqa: E402

def ray_records(o, d, near, far, hist):
    """[N, 11+hist_bin] records [o3, d3, near, far, viewdir3, hist] (reference rendering.py:366-389)."""
    o, d = o.reshape(-1, 3).float(), d.reshape(-1, 3).float()
    n = o.shape[0]
    vd = d / d.norm(dim=-1, keepdim=True)
    nf = torch.ones(n, 1, device=o.device)
    return torch.cat([o, d, near * nf, far * nf, vd, torch.as_tensor(hist, device=o.device).float().reshape(1, -1).expand(n, -1)], -1).contiguous()


dev = torch.device("cuda:0")
mods = nerfw.make_synthetic_nerf(D=8, W=256)
h = ops.NerfHandle(*[m.to(dev) for m in mods])
c2w = torch.tensor([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 1.0]], device=dev)
hist = torch.tensor([5, 10, 20, 30, 15, 10, 5, 3, 1, 1.0], device=dev)
for mma in ("f16", "bf16", "fp32"):
    o = h.render(64, 128, True, c2w=c2w, H=9, W=13, focal=11.0, near=0.0, far=2.5, hist=hist, mma=mma)
# training forward with saved ReLU masks + tcgen05 backward (saved masks and forward recompute), fp32 backward, 1-CTA kernel
ro, rd = ops.get_rays(9, 13, 11.0, c2w)
rec = ray_records(ro, rd, 0.0, 2.5, hist)
g = torch.randn(9 * 13, 3, device=dev) * 1e-6
for mma in ("f16", "bf16"):
    t = h.render(64, 128, True, rays=rec, mma=mma, want=("z_vals", "raw", "relu_masks"))
    h.render_backward(rec, t["z_vals"], t["raw"], g, mma=mma, relu_masks=t["relu_masks"])
    h.render_backward(rec, t["z_vals"], t["raw"], g, mma=mma)
h.render_backward(rec, t["z_vals"], t["raw"], g, mma="fp32")
os.environ["DFB_TC_CTA_GROUP"] = "1"
h.render(64, 128, True, c2w=c2w, H=9, W=13, focal=11.0, near=0.0, far=2.5, hist=hist, mma="f16")
del os.environ["DFB_TC_CTA_GROUP"]
o = h.render(64, 128, False, c2w=c2w, H=5, W=7, focal=11.0, near=0.0, far=2.5, hist=hist, mma="fp32",
             want=("rgb0", "beta", "z_std", "raw"))
net = synthetic_dfnet("DFNet").to(dev)
x = torch.rand(2, 3, 37, 53, device=dev)
feats, pose = net(x, return_feature=True, isSingleStream=False, return_pose=True, upsampleH=37, upsampleW=53)
fs, _ = net(x, return_feature=True, isSingleStream=True, return_pose=False, upsampleH=20, upsampleW=30)
l = feature_loss(feats[1][0, 0], feats[0][0, 0])
l2 = feature_loss(feats[1][0, 0], feats[0][0, 0], per_channel=True)
t = triplet_loss_hard_negative_mining_plus(torch.randn(3, 2, 16, 5, 9, device=dev), torch.randn(3, 2, 16, 5, 9, device=dev))
m = mse(x, x * 0.5)
u = upsample_bicubic(x, (74, 100))
torch.cuda.synchronize()
print("sanitize smoke ok", float(l), float(l2), float(t), float(m), tuple(u.shape), float(o["rgb"].mean()))
