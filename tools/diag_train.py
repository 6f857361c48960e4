"""Bring-up diagnostic: per-layer agreement of the pose-regressor parameter gradients with a torch reference."""
import os
import sys

import torch
import torch.nn.functional as F

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root)
sys.path.insert(0, os.path.join(root, "tests"))
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
from helpers import synthetic_dfnet  # noqa: E402
from test_train_gpu import torch_dfnet_forward  # noqa: E402
from dfnet_b200 import misc  # noqa: E402

dev = torch.device("cuda:0")
B, H, W = 1, 64, 96
net = synthetic_dfnet("DFNet", seed=7).to(dev)
net.train()
for m in net.modules():
    if isinstance(m, torch.nn.BatchNorm2d):
        m.eval()
        m.weight.requires_grad_(False), m.bias.requires_grad_(False)
torch.manual_seed(8)
x = torch.rand(B, 3, H, W, device=dev)
target = torch.randn(B, 12, device=dev)
_, pose = net(x, return_feature=False, isSingleStream=True, return_pose=True, upsampleH=H, upsampleW=W)
misc.mse(pose, target).backward()
got = {n: p.grad.clone() for n, p in net.named_parameters() if p.grad is not None}
acts = net._handle.tape_activations()
for quant, pin in ((torch.bfloat16, acts), (torch.bfloat16, None), (None, None)):
    net.zero_grad()
    _, pose_t = torch_dfnet_forward(net, x, False, True, True, H, W, quant, pin)
    F.mse_loss(pose_t, target).backward()
    print("quant", quant, "pinned" if pin else "free", "pose err", float((pose - pose_t).abs().max()), float(pose_t.abs().max()))
    for n, p in net.named_parameters():
        if p.grad is None:
            continue
        a, b = got[n].double().flatten(), p.grad.double().flatten()
        print(f"  {n:32s} cos={float(F.cosine_similarity(a, b, dim=0)):.6f} |got|={float(a.norm()):.3e} |want|={float(b.norm()):.3e}")
