"""Host-side mirror of the reference's `models/nerfw.py` for the render hot path.

`NeRFW` keeps the reference's parameter names, registration order and RNG behaviour
(reference models/nerfw.py:220-295) so that reference checkpoints load unchanged and a
freshly constructed module holds bit-identical weights.  Its arithmetic runs in the
sm_100a kernels behind the C ABI (include/dfnet_b200.h); there is no eager fallback.
"""
import os

import torch
import torch.nn as nn


class NeRFW(nn.Module):
    """NeRF-W / NeRF-Hist MLP container (reference models/nerfw.py:220-295).

    state_dict keys: xyz_encoding_{1..D}.0.{weight,bias}, xyz_encoding_final,
    dir_encoding.0, static_sigma.0, static_rgb.0 and, for the fine network,
    transient_encoding.{0,2,4,6}, transient_{sigma,rgb,beta}.0.
    """

    def __init__(self, typ, D=8, W=256, skips=(4,), in_channels_xyz=63, in_channels_dir=27,
                 encode_appearance=False, in_channels_a=48, encode_transient=False,
                 in_channels_t=16, beta_min=0.1, out_ch_size=3):
        super().__init__()
        # The reference reseeds the global RNG inside __init__ (nerfw.py:245); coarse and
        # fine trunks therefore share their initial weights.  Kept for parity.
        torch.manual_seed(0)
        if out_ch_size != 3:
            raise NotImplementedError("feature-output NeRFW (out_ch_size != 3) is outside the hot path")
        self.typ = typ
        self.D, self.W, self.skips = D, W, list(skips)
        self.in_channels_xyz, self.in_channels_dir = in_channels_xyz, in_channels_dir
        self.encode_appearance = False if typ == "coarse" else encode_appearance
        self.in_channels_a = in_channels_a if encode_appearance else 0
        self.encode_transient = False if typ == "coarse" else encode_transient
        self.in_channels_t = in_channels_t
        self.beta_min = beta_min

        for i in range(D):
            fan_in = in_channels_xyz if i == 0 else (W + in_channels_xyz if i in self.skips else W)
            setattr(self, f"xyz_encoding_{i + 1}", nn.Sequential(nn.Linear(fan_in, W), nn.ReLU(True)))
        self.xyz_encoding_final = nn.Linear(W, W)
        self.dir_encoding = nn.Sequential(
            nn.Linear(W + in_channels_dir + self.in_channels_a, W // 2), nn.ReLU(True))
        self.static_sigma = nn.Sequential(nn.Linear(W, 1), nn.Softplus())
        self.static_rgb = nn.Sequential(nn.Linear(W // 2, 3), nn.Sigmoid())
        if self.encode_transient:
            self.transient_encoding = nn.Sequential(
                nn.Linear(W + in_channels_t, W // 2), nn.ReLU(True),
                nn.Linear(W // 2, W // 2), nn.ReLU(True),
                nn.Linear(W // 2, W // 2), nn.ReLU(True),
                nn.Linear(W // 2, W // 2), nn.ReLU(True))
            self.transient_sigma = nn.Sequential(nn.Linear(W // 2, 1), nn.Softplus())
            self.transient_rgb = nn.Sequential(nn.Linear(W // 2, 3), nn.Sigmoid())
            self.transient_beta = nn.Sequential(nn.Linear(W // 2, 1), nn.Softplus())

    def forward(self, x, sigma_only=False, output_transient=True):
        """Embedded points -> raw outputs (reference nerfw.py:297-354) on the CUDA path."""
        from . import ops
        return ops.nerfw_forward(self, x, sigma_only=sigma_only, output_transient=output_transient)


def synthetic_init_(model, gain=1.6, sigma_bias=-1.0):
    """Benchmark/test initialisation from SURVEY.md §8(d): default init, every Linear
    weight scaled by `gain`, sigma-head bias set to `sigma_bias` (gives the random field
    structure and acc ~ 1)."""
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, nn.Linear):
                m.weight.mul_(gain)
        model.static_sigma[0].bias.fill_(sigma_bias)
    return model


def make_synthetic_nerf(D=8, W=256, in_channels_a=50, in_channels_t=20, n_vocab=1000, seed=0,
                        gain=1.6, sigma_bias=-1.0, fine=True):
    """Coarse + fine NeRFW and the two histogram embeddings, seeded like
    run_nerf.py:24-27 then create_nerf (reference nerfw.py:386-419)."""
    torch.manual_seed(seed)
    emb_a = nn.Embedding(n_vocab, 5)
    emb_t = nn.Embedding(n_vocab, 2)
    coarse = synthetic_init_(NeRFW("coarse", D=D, W=W), gain, sigma_bias)
    net_fine = None
    if fine:
        net_fine = synthetic_init_(
            NeRFW("fine", D=D, W=W, encode_appearance=True, encode_transient=True,
                  in_channels_a=in_channels_a, in_channels_t=in_channels_t), gain, sigma_bias)
    return coarse, net_fine, emb_a, emb_t
