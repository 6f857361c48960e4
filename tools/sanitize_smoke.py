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
# ---- round 2: native 128-wide program (+ its zero-padded embedding), early ray termination, split-precision coarse
# pass, the 1-CTA variants of the backward and convolution kernels, a train_on_batch step (weight re-packing as one
# launch, conv data / weight / bias gradients, saved-mask backward on cta_group::2), a NeRF-Hist training step
mods128 = nerfw.make_synthetic_nerf(D=8, W=128)
h128 = ops.NerfHandle(*[m.to(dev) for m in mods128])
for env in ({}, {"DFB_TC_NATIVE128": "0"}):
    os.environ.update(env)
    h128.render(64, 64, True, c2w=c2w, H=9, W=13, focal=11.0, near=0.0, far=2.5, hist=hist, mma="f16")
    h128.render(64, 64, True, rays=rec, mma="bf16", want=("raw", "depth"))
    for k in env:
        del os.environ[k]
h.render(64, 128, True, c2w=c2w, H=9, W=13, focal=11.0, near=0.0, far=2.5, hist=hist, mma="f16", ert_eps=1e-2)
h.render(64, 128, True, c2w=c2w, H=9, W=13, focal=11.0, near=0.0, far=2.5, hist=hist, mma="f16s")
os.environ["DFB_TC_CTA_GROUP"] = "1"
t1 = h.render(64, 128, True, rays=rec, mma="f16", want=("z_vals", "raw", "relu_masks"))
h.render_backward(rec, t1["z_vals"], t1["raw"], g, mma="f16", relu_masks=t1["relu_masks"])
h.render_backward(rec, t1["z_vals"], t1["raw"], g, mma="f16")
del os.environ["DFB_TC_CTA_GROUP"]
os.environ["DFB_CONV_CTA_GROUP"] = "1"
net(x, return_feature=True, isSingleStream=False, return_pose=True, upsampleH=37, upsampleW=53)
del os.environ["DFB_CONV_CTA_GROUP"]
import types  # noqa: E402
from dfnet_b200 import direct_feature_matching as dfm, nerf_train  # noqa: E402
from dfnet_b200.losses import loss_dict  # noqa: E402
Fnet, Gnet = synthetic_dfnet("DFNet", seed=0).to(dev), synthetic_dfnet("DFNet", seed=1).to(dev).eval()
for p_ in Gnet.parameters():
    p_.requires_grad_(False)
Fnet.train()
for m_ in Fnet.modules():
    if isinstance(m_, torch.nn.BatchNorm2d):
        m_.eval()
        m_.weight.requires_grad_(False), m_.bias.requires_grad_(False)
nets = [m_.to(dev) for m_ in mods]
for m_ in nets:
    for p_ in m_.parameters():
        p_.requires_grad_(False)
kw = dict(network_query_fn=None, perturb=0.0, N_importance=128, network_fine=nets[1], N_samples=64, network_fn=nets[0],
          use_viewdirs=True, white_bkgd=False, raw_noise_std=0.0, embedding_a=nets[2], embedding_t=nets[3], test_time=True,
          ndc=False, lindisp=False, near=0.0, far=2.5, mma="f16")
targs = types.SimpleNamespace(DFNet=True, preprocess_ImgNet=False, svd_reg=True, combine_loss=True, per_channel=False, chunk=32768,
                              batch_size=1, combine_loss_w=[0.0, 0.0, 1.0], feature_matching_lvl=[0])
opt = torch.optim.Adam([p_ for p_ in Fnet.parameters() if p_.requires_grad], lr=1e-5)
for _ in range(2):   # the second step re-packs the updated weights
    dfm.train_on_batch(targs, torch.rand(1, 3, 64, 96), Fnet, Gnet, torch.tensor([[1., 0, 0, 0.1, 0, 1, 0, -0.05, 0, 0, 1, 2.0]]),
                       hist[None].cpu(), (64, 96, 80.0), opt, True, dev, dict(pose_scale=0.5, pose_scale2=1.0, move_all_cam_vec=[0.0, 0.0, 0.05]), **kw)
tm = [m_.to(dev) for m_ in nerfw.make_synthetic_nerf(D=8, W=128, fine=True)]
tp = [p_ for m_ in tm for p_ in m_.parameters()]
for p_ in tp:
    p_.requires_grad_(True)
nkw = dict(network_query_fn=None, perturb=1.0, N_importance=64, network_fine=tm[1], N_samples=64, network_fn=tm[0], use_viewdirs=True,
           white_bkgd=False, raw_noise_std=1.0, embedding_a=tm[2], embedding_t=tm[3], test_time=False, ndc=False, lindisp=False)
nerf_train.train_on_batch_nerfw(types.SimpleNamespace(chunk=32768, lrate=5e-4, lrate_decay=250), torch.rand(3, 24, 32),
                                torch.tensor([1., 0, 0, 0.1, 0, 1, 0, -0.05, 0, 0, 1, 2.0]), hist[None].cpu(), 24, 32, 30.0, 200,
                                torch.optim.Adam(tp, lr=5e-4), loss_dict["nerfw"](coef=1), 0, nkw, near=0.0, far=2.5)
torch.cuda.synchronize()
print("sanitize smoke ok", float(l), float(l2), float(t), float(m), tuple(u.shape), float(o["rgb"].mean()))
