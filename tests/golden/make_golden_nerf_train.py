"""Golden NeRF-Hist training step from the UNMODIFIED reference (CPU, fp32): the body of train_on_epoch_nerfw
(script/run_nerf.py:44-66: render(..., retraw=True, **render_kwargs_train), NerfWLoss, loss.backward()) on seeded networks.

    python tests/golden/make_golden_nerf_train.py   ->  tests/golden/nerf_train_golden.npz

perturb = 0 and raw_noise_std = 0 (the reference draws its jitter from the CPU generator, which a GPU run cannot reproduce);
everything else is the training configuration: test_time=False, coarse net with rgb, fine net with transient head."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from baseline import ref_runner, ref_shims  # noqa: E402

ref_shims.activate()
import torch  # noqa: E402
from helpers import nerf_train_case  # noqa: E402
from models.losses import loss_dict  # noqa: E402
from models.rendering import render  # noqa: E402

G = {}
for case in ("w128", "w256"):
    cfg = nerf_train_case(case)
    mods = cfg["mods"]
    kw = ref_runner.reference_render_kwargs(mods, test_time=False, N_samples=cfg["Nc"], N_importance=cfg["Nf"])
    kw["perturb"] = 0.0
    nets = [kw["network_fn"], kw["network_fine"], kw["embedding_a"], kw["embedding_t"]]
    for m in nets:
        for p in m.parameters():
            p.requires_grad_(True)
    rays = torch.from_numpy(cfg["rays"])
    target = torch.from_numpy(cfg["target"])
    rgb, disp, acc, extras = render(1, 1, 1.0, chunk=32768, rays=(rays[0], rays[1]), retraw=True, near=cfg["near"], far=cfg["far"],
                                    img_idx=torch.from_numpy(cfg["hist"]), **kw)
    results = {"rgb_fine": rgb, "rgb_coarse": extras["rgb0"], "beta": extras["beta"], "transient_sigmas": extras["transient_sigmas"]}
    loss_d = loss_dict["nerfw"](coef=1)(results, target)
    loss = sum(l for l in loss_d.values())
    loss.backward()
    G[f"{case}_loss"] = np.array(float(loss))
    for k, v in loss_d.items():
        G[f"{case}_{k}"] = np.array(float(v))
    G[f"{case}_rgb"], G[f"{case}_rgb0"], G[f"{case}_beta"] = rgb.detach().numpy(), extras["rgb0"].detach().numpy(), extras["beta"].detach().numpy()
    names = []
    for tag, m in zip(("coarse", "fine", "emb_a", "emb_t"), nets):
        for n, p in m.named_parameters():
            key = f"{tag}.{n}"
            names.append(key)
            g = p.grad if p.grad is not None else torch.zeros_like(p)
            flat = g.flatten()
            G[f"{case}_g_{key}_stats"] = np.array([float(flat.norm()), float(flat.abs().max())])
            G[f"{case}_g_{key}_sub"] = flat[:: max(1, flat.numel() // 2048)][:2048].numpy().copy()
    G[f"{case}_names"] = np.frombuffer("\n".join(names).encode(), np.uint8)
    print(case, "loss", float(loss), {k: float(v) for k, v in loss_d.items()})
out = os.path.join(HERE, "nerf_train_golden.npz")
np.savez_compressed(out, **G)
print("wrote", out, os.path.getsize(out) / 1e3, "KB")
