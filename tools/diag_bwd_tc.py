"""Diagnostics of the tcgen05 render backward (dfb_render_bwd_mma) against the fp32 kernels: error statistics by
pass region, timing of both kernels."""
import os
import sys

import numpy as np
import torch

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root)
sys.path.insert(0, os.path.join(root, "tests"))
from dfnet_b200 import nerfw, ops  # noqa: E402

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
Hh, Ww = int(os.environ.get("DH", 40)), int(os.environ.get("DW", 50))
c2w = torch.tensor([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 1.0]], device=dev)
hist = np.array([[5, 10, 20, 30, 15, 10, 5, 3, 1, 1]], np.float32)
o, d = ops.get_rays(Hh, Ww, 45.0, c2w)
rec = ray_records(o, d, 0.0, 2.5, hist)
rng = np.random.RandomState(7)
g_rgb = torch.tensor((rng.randn(Hh * Ww, 3) * 1e-7).astype(np.float32), device=dev)
for mma in ("f16", "bf16"):
    out = h.render(64, 128, True, rays=rec, mma=mma, want=("z_vals", "raw"))
    want = h.render_backward(rec, out["z_vals"], out["raw"], g_rgb, mma="fp32")
    got = h.render_backward(rec, out["z_vals"], out["raw"], g_rgb, mma=mma)
    got2 = h.render_backward(rec, out["z_vals"], out["raw"], g_rgb, mma=mma)
    torch.cuda.synchronize()
    n1 = 148 * 256 // 192 - 2
    for nm, a, b, a2 in zip(("g_o", "g_d", "g_vd"), got, want, got2):
        a, b = a.double(), b.double()
        def st(x, y):
            return (float((x - y).norm() / y.norm()), float((x - y).abs().max() / y.abs().max()),
                    float((x * y).sum() / (x.norm() * y.norm())))
        print(f"{mma} {nm}: all relL2 {st(a, b)[0]:.2e} maxrel {st(a, b)[1]:.2e} cos {st(a, b)[2]:.6f} | first pass relL2 "
              f"{st(a[:n1], b[:n1])[0]:.2e} maxrel {st(a[:n1], b[:n1])[1]:.2e} | later relL2 {st(a[n1 + 4:], b[n1 + 4:])[0]:.2e} "
              f"maxrel {st(a[n1 + 4:], b[n1 + 4:])[1]:.2e} | finite {bool(torch.isfinite(a).all())} deterministic {bool(torch.equal(a2.double(), a))}")
    for kind in ("fp32", mma):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for _ in range(3):
            h.render_backward(rec, out["z_vals"], out["raw"], g_rgb, mma=kind)
        ev[1].record()
        torch.cuda.synchronize()
        print(f"  render_backward mma={kind}: {ev[0].elapsed_time(ev[1]) / 3:.3f} ms for {Hh * Ww} rays x 192 samples")

# saved-mask path (no forward recompute) vs recompute, 19 200 rays (cfg4 render size)
o2, d2 = ops.get_rays(120, 160, 146.25, c2w)
rec2 = ray_records(o2, d2, 0.0, 2.5, hist)
g2 = torch.tensor((rng.randn(120 * 160, 3) * 1e-7).astype(np.float32), device=dev)
for want in (("z_vals", "raw"), ("z_vals", "raw", "relu_masks")):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    for rep in range(2):
        ev[0].record()
        out2 = h.render(64, 128, True, rays=rec2, mma="f16", want=want)
        ev[1].record()
        h.render_backward(rec2, out2["z_vals"], out2["raw"], g2, mma="f16", relu_masks=out2.get("relu_masks"))
        ev[2].record()
        torch.cuda.synchronize()
    print(f"19200 rays, extras {want}: forward {ev[0].elapsed_time(ev[1]):.3f} ms, backward {ev[1].elapsed_time(ev[2]):.3f} ms")
