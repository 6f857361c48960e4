"""Diagnostic: reproduce the failing sequence of tests/test_render_gpu.py::test_render_narrow_networks_on_tensor_cores."""
import os, sys
import numpy as np, torch
root = os.getcwd()
sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "tests"))
from dfnet_b200 import ops, rendering
from helpers import synthetic_nets
dev = torch.device("cuda:0")
def cos(a, b):
    a, b = a.double().reshape(-1), b.double().reshape(-1)
    return float((a * b).sum() / (a.norm() * b.norm() + 1e-300))
gold = np.load(os.path.join(root, "tests", "golden", "render_golden.npz"))
rays = gold["e2e_c_rays"]
hist_g = torch.tensor(gold["hist"], device=dev)
W, Nf = 128, 64
mods, _ = synthetic_nets(8, W)
dm = [m.to(dev) for m in mods]
if os.environ.get("PART1", "1") == "1":
    h = ops.NerfHandle(*dm)
    kw = dict(c2w=torch.tensor(gold["e2e_b_c2w"], device=dev), H=12, W=16, focal=14.6, near=0.0, far=2.5, hist=hist_g)
    for mma in os.environ.get("P1MMA", "fp32,f16,bf16").split(","):
        h.render(64, Nf, True, mma=mma, **kw)
    torch.cuda.synchronize()
kw2 = dict(network_query_fn=None, perturb=0.0, N_importance=Nf, network_fine=dm[1], N_samples=64, network_fn=dm[0],
           use_viewdirs=True, white_bkgd=False, raw_noise_std=0.0, embedding_a=dm[2], embedding_t=dm[3], test_time=True,
           ndc=False, lindisp=False)
res = {}
for mma in ("f16", "fp32"):
    ro = torch.tensor(rays[0], device=dev).clone().requires_grad_(True)
    rd = torch.tensor(rays[1], device=dev).clone().requires_grad_(True)
    rgb, _, _, _ = rendering.render(4, 6, 5.0, rays=(ro, rd), img_idx=hist_g, near=0.0, far=2.5, mma=mma, **kw2)
    torch.manual_seed(0)
    (rgb * torch.randn_like(rgb)).sum().backward()
    res[mma] = (ro.grad.clone(), rgb.detach().clone())
print("cos autograd f16~fp32:", round(cos(res["f16"][0], res["fp32"][0]), 4), "rgb maxdiff", float((res["f16"][1] - res["fp32"][1]).abs().max()))
# direct: same handle as the autograd path
h2 = ops.handle_for(dm[0], dm[1], dm[2], dm[3])
r_o, r_d = torch.tensor(rays[0], device=dev).reshape(-1, 3), torch.tensor(rays[1], device=dev).reshape(-1, 3)
rec = torch.cat([r_o, r_d, torch.zeros(r_o.shape[0], 1, device=dev), 2.5 * torch.ones(r_o.shape[0], 1, device=dev),
                 torch.nn.functional.normalize(r_d, dim=-1), hist_g.reshape(1, -1).expand(r_o.shape[0], -1)], -1).contiguous()
out = h2.render(64, Nf, True, rays=rec, mma="f16", want=("z_vals", "raw", "relu_masks"))
torch.manual_seed(0)
g = torch.randn(r_o.shape[0], 3, device=dev)
a = h2.render_backward(rec, out["z_vals"], out["raw"], g, mma="f16", relu_masks=out["relu_masks"])
b = h2.render_backward(rec, out["z_vals"], out["raw"], g, mma="f16")
c = h2.render_backward(rec, out["z_vals"], out["raw"], g, mma="fp32")
print("direct: saved~recompute", round(cos(a[0], b[0]), 4), "saved~fp32", round(cos(a[0], c[0]), 4), "recompute~fp32", round(cos(b[0], c[0]), 4))
