"""Generate golden vectors by running the UNMODIFIED reference (/root/reference).

Run in the build container only (the reference does not travel to the GPU box):

    python tests/golden/make_golden.py

Writes tests/golden/render_golden.npz.  Network weights are not stored: the script
asserts that dfnet_b200.nerfw.NeRFW reproduces the reference module's state_dict bit
for bit under the same seed and records a checksum that the tests re-check.
"""
import hashlib
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

for _m in ["imageio"]:
    sys.modules.setdefault(_m, types.ModuleType(_m))
sys.path[:0] = ["/root/reference/script", "/root/reference"]

import torch  # noqa: E402

from models import nerfw as ref_nerfw  # noqa: E402
from models import ray_utils as ref_rays  # noqa: E402
from models import rendering as ref_rend  # noqa: E402

from dfnet_b200 import nerfw as my_nerfw  # noqa: E402

torch.set_num_threads(8)
G = {}


def put(name, t):
    if isinstance(t, torch.Tensor):
        t = t.detach().cpu().numpy()
    G[name] = np.ascontiguousarray(t)


def sd_checksum(sd):
    h = hashlib.sha256()
    for k in sd:
        h.update(k.encode())
        h.update(sd[k].detach().cpu().numpy().tobytes())
    return h.hexdigest()


def build_nets(D, W, fine=True):
    """Reference modules initialised like dfnet_b200.nerfw.make_synthetic_nerf."""
    torch.manual_seed(0)
    emb_a = torch.nn.Embedding(1000, 5)
    emb_t = torch.nn.Embedding(1000, 2)
    coarse = my_nerfw.synthetic_init_(ref_nerfw.NeRFW("coarse", D=D, W=W, skips=[4]))
    net_fine = None
    if fine:
        net_fine = my_nerfw.synthetic_init_(ref_nerfw.NeRFW(
            "fine", D=D, W=W, skips=[4], encode_appearance=True, encode_transient=True,
            in_channels_a=50, in_channels_t=20))
    mc, mf, ma, mt = my_nerfw.make_synthetic_nerf(D=D, W=W, fine=fine)
    for a, b in [(coarse, mc), (net_fine, mf), (emb_a, ma), (emb_t, mt)]:
        if a is None:
            continue
        sa, sb = a.state_dict(), b.state_dict()
        assert list(sa) == list(sb), (list(sa), list(sb))
        for k in sa:
            assert torch.equal(sa[k], sb[k]), k
    return coarse, net_fine, emb_a, emb_t


def render_kwargs(coarse, fine, emb_a, emb_t, Nc, Nf, test_time, perturb=0.0):
    """The dict create_nerf builds (reference nerfw.py:425-434,476-500), by hand because
    the factory hard-codes torch.device('cuda')."""
    embed_fn, _, _ = ref_nerfw.get_embedder(10, 0, -1)
    embeddirs_fn, _, _ = ref_nerfw.get_embedder(4, 0, -1)

    def q(inputs, viewdirs, ts, network_fn, typ, embedding_a, embedding_t, output_transient, test_time):
        return ref_nerfw.run_network_NeRFW(inputs, viewdirs, ts, network_fn, embed_fn=embed_fn,
                                           embeddirs_fn=embeddirs_fn, typ=typ, embedding_a=embedding_a,
                                           embedding_t=embedding_t, output_transient=output_transient,
                                           netchunk=65536, test_time=test_time)
    return dict(network_query_fn=q, perturb=perturb, N_importance=Nf, network_fine=fine, N_samples=Nc,
                network_fn=coarse, use_viewdirs=True, white_bkgd=False, raw_noise_std=0.0,
                embedding_a=emb_a, embedding_t=emb_t, test_time=test_time, ndc=False, lindisp=False)


def small_pose(seed, max_deg=10.0):
    rng = np.random.RandomState(seed)
    ax = rng.randn(3)
    ax /= np.linalg.norm(ax)
    ang = np.deg2rad(max_deg) * rng.rand()
    K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
    R = np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K
    c2w = np.concatenate([R, np.array([[0.0], [0.0], [1.0]])], 1).astype(np.float32)
    return c2w


class Capture:
    """Record outputs of torch.searchsorted / torch.rand / sample_pdf while the reference runs."""

    def __init__(self):
        self.inds, self.rand, self.zs, self.w = [], [], [], []

    def __enter__(self):
        self._ss, self._rand, self._sp = torch.searchsorted, torch.rand, ref_rend.sample_pdf

        def ss(*a, **k):
            r = self._ss(*a, **k)
            self.inds.append(r.clone())
            return r

        def rnd(*a, **k):
            r = self._rand(*a, **k)
            self.rand.append(r.clone())
            return r

        def sp(bins, weights, *a, **k):
            self.w.append(weights.detach().clone())
            r = self._sp(bins, weights, *a, **k)
            self.zs.append(r.detach().clone())
            return r
        torch.searchsorted, torch.rand, ref_rend.sample_pdf = ss, rnd, sp
        return self

    def __exit__(self, *e):
        torch.searchsorted, torch.rand, ref_rend.sample_pdf = self._ss, self._rand, self._sp


HIST = np.array([[5, 10, 20, 30, 15, 10, 5, 3, 1, 1]], np.float32)


def main():
    rng = np.random.RandomState(1234)

    # --- ATen arithmetic pins -------------------------------------------------------
    for i, (a, b, n) in enumerate([(0, 1, 64), (0, 1, 128), (0, 1, 192), (0, 2.5, 64), (0, 20, 63), (2, 6, 7), (0, 1, 2)]):
        put(f"linspace_{i}_args", np.array([a, b, n], np.float64))
        put(f"linspace_{i}", torch.linspace(a, b, n))
    for n in (62, 64, 14, 192, 1000):
        x = (rng.rand(48, n).astype(np.float32)) ** 3
        put(f"sum_{n}_x", x)
        put(f"sum_{n}", torch.sum(torch.from_numpy(x), -1))
    x = (rng.rand(16, 62).astype(np.float32)) ** 3
    put("cumsum_x", x)
    put("cumsum", torch.cumsum(torch.from_numpy(x), -1))
    x = 1 - (rng.rand(16, 192).astype(np.float32)) ** 4
    put("cumprod_x", x)
    put("cumprod", torch.cumprod(torch.from_numpy(x), -1))

    # --- a1 get_rays ----------------------------------------------------------------
    c2w = small_pose(3)
    put("rays_c2w", c2w)
    put("rays_hwf", np.array([5, 7, 6.5]))
    o, d = ref_rays.get_rays(5, 7, 6.5, torch.from_numpy(c2w))
    put("rays_o", o.contiguous())
    put("rays_d", d)

    # --- a6 embed -------------------------------------------------------------------
    pts = (rng.randn(32, 3) * 1.5).astype(np.float32)
    e10, _, _ = ref_nerfw.get_embedder(10, 0, -1)
    e4, _, _ = ref_nerfw.get_embedder(4, 0, -1)
    put("embed_x", pts)
    put("embed_L10", e10(torch.from_numpy(pts)))
    put("embed_L4", e4(torch.from_numpy(pts)))

    # --- a7 NeRFW.forward, three modes, two sizes ----------------------------------------
    with torch.no_grad():
        for tag, D, W in (("s", 4, 64), ("b", 8, 256)):
            coarse, fine, emb_a, emb_t = build_nets(D, W)
            put(f"mlp_{tag}_coarse_sha", np.frombuffer(sd_checksum(coarse.state_dict()).encode(), np.uint8))
            put(f"mlp_{tag}_fine_sha", np.frombuffer(sd_checksum(fine.state_dict()).encode(), np.uint8))
            x = torch.from_numpy(np.concatenate([
                e10(torch.from_numpy((rng.randn(40, 3) * 1.2).astype(np.float32))).numpy(),
                e4(torch.from_numpy(rng.randn(40, 3).astype(np.float32))).numpy(),
                rng.randn(40, 70).astype(np.float32)], 1))
            put(f"mlp_{tag}_x", x)
            put(f"mlp_{tag}_sigma_only", coarse(x[:, :63], sigma_only=True))
            put(f"mlp_{tag}_coarse_static", coarse(x[:, :90], output_transient=False))
            put(f"mlp_{tag}_fine_full", fine(x, output_transient=True))

        # --- a8 raw2outputs_NeRFW ----------------------------------------------------
        raw = rng.randn(16, 24, 9).astype(np.float32)
        raw[..., :3] = 1 / (1 + np.exp(-raw[..., :3]))
        raw[..., 4:7] = 1 / (1 + np.exp(-raw[..., 4:7]))
        raw[..., 3] = np.log1p(np.exp(2 * raw[..., 3]))
        raw[..., 7] = np.log1p(np.exp(raw[..., 7] - 1))
        raw[..., 8] = np.log1p(np.exp(raw[..., 8]))
        z = np.sort(rng.rand(16, 24).astype(np.float32) * 2.5, -1)
        put("r2o_raw", raw)
        put("r2o_z", z)
        rt, zt, dd = torch.from_numpy(raw), torch.from_numpy(z), torch.zeros(16, 3)
        names = ["rgb", "disp", "acc", "weights", "depth", "transient_sigmas", "beta"]
        cases = {
            "coarse_test": dict(raw=rt[..., 3:4], output_transient=False, test_time=True, typ="coarse"),
            "coarse_train": dict(raw=rt[..., :4], output_transient=False, test_time=False, typ="coarse"),
            "fine_test": dict(raw=rt, output_transient=True, test_time=True, typ="fine"),
            "fine_train": dict(raw=rt, output_transient=True, test_time=False, typ="fine"),
        }
        for cname, kw in cases.items():
            r = kw.pop("raw")
            outs = ref_rend.raw2outputs_NeRFW(r, zt, dd, 0.0, kw.pop("output_transient"), 0.1, False, **kw)
            for nm, v in zip(names, outs):
                if v is not None:
                    put(f"r2o_{cname}_{nm}", v)

        # --- a9 sample_pdf ------------------------------------------------------------
        zc = np.broadcast_to(torch.linspace(0, 1, 64).numpy() * 2.5, (40, 64)).astype(np.float32)
        bins = 0.5 * (zc[:, 1:] + zc[:, :-1])
        w = (rng.rand(40, 62).astype(np.float32)) ** 6
        w[3] = 0.0
        w[4, :] = 0.0
        w[4, 17] = 0.9
        w[5, 40:] = 0.0
        put("pdf_bins", bins)
        put("pdf_w", w)
        with Capture() as cap:
            s_det = ref_rend.sample_pdf(torch.from_numpy(bins), torch.from_numpy(w), 128, det=True)
            s_rnd = ref_rend.sample_pdf(torch.from_numpy(bins), torch.from_numpy(w), 128, det=False, pytest=True)
        np.random.seed(0)
        put("pdf_u_rand", np.random.rand(40, 128).astype(np.float32))
        put("pdf_det_samples", s_det)
        put("pdf_det_inds", cap.inds[0])
        put("pdf_rand_samples", s_rnd)
        put("pdf_rand_inds", cap.inds[1])

        # --- end-to-end render() -------------------------------------------------------
        hist = torch.from_numpy(HIST)
        put("hist", HIST)

        # (a) cfg1-shaped: coarse only, 4x64, train-mode path
        coarse, _, emb_a, emb_t = build_nets(4, 64, fine=False)
        kw = render_kwargs(coarse, None, emb_a, emb_t, 64, 0, test_time=False)
        c2w = small_pose(11)
        put("e2e_a_c2w", c2w)
        rgb, disp, acc, _ = ref_rend.render(8, 8, 8.0, chunk=32768, c2w=torch.from_numpy(c2w), img_idx=hist,
                                            near=0.0, far=2.5, **kw)
        put("e2e_a_rgb", rgb), put("e2e_a_disp", disp), put("e2e_a_acc", acc)

        # (b) cfg2-shaped: 8x256 coarse+fine, 64+128, test_time, 6x8 rays
        coarse, fine, emb_a, emb_t = build_nets(8, 256)
        kw = render_kwargs(coarse, fine, emb_a, emb_t, 64, 128, test_time=True)
        c2w = small_pose(12)
        put("e2e_b_c2w", c2w)
        with Capture() as cap:
            rgb, disp, acc, _ = ref_rend.render(6, 8, 7.3125, chunk=32768, c2w=torch.from_numpy(c2w),
                                                img_idx=hist, near=0.0, far=2.5, **kw)
        put("e2e_b_rgb", rgb), put("e2e_b_disp", disp), put("e2e_b_acc", acc)
        put("e2e_b_inds", cap.inds[0]), put("e2e_b_z_samples", cap.zs[0]), put("e2e_b_w_coarse", cap.w[0])

        # (c) train-mode coarse+fine with explicit rays, 8x64 net (skip active), 16+24, retraw
        coarse, fine, emb_a, emb_t = build_nets(8, 64)
        kw = render_kwargs(coarse, fine, emb_a, emb_t, 16, 24, test_time=False)
        o, d = ref_rays.get_rays(4, 6, 5.0, torch.from_numpy(small_pose(13)))
        sel = torch.tensor([0, 3, 5, 7, 8, 12, 13, 17, 20, 23])
        rays = torch.stack([o.reshape(-1, 3)[sel], d.reshape(-1, 3)[sel]], 0).contiguous()
        put("e2e_c_rays", rays)
        with Capture() as cap:
            rgb, disp, acc, ex = ref_rend.render(4, 6, 5.0, chunk=32768, rays=rays, img_idx=hist,
                                                 near=0.0, far=2.5, retraw=True, **kw)
        put("e2e_c_rgb", rgb), put("e2e_c_disp", disp), put("e2e_c_acc", acc)
        for k, v in ex.items():
            put(f"e2e_c_{k}", v)
        put("e2e_c_inds", cap.inds[0])

        # (d) train-mode with stratified jitter: torch.rand draws captured (t_rand, u)
        kw = render_kwargs(coarse, fine, emb_a, emb_t, 16, 24, test_time=False, perturb=1.0)
        torch.manual_seed(7)
        with Capture() as cap:
            rgb, disp, acc, ex = ref_rend.render(4, 6, 5.0, chunk=32768, rays=rays, img_idx=hist,
                                                 near=0.0, far=2.5, **kw)
        put("e2e_d_t_rand", cap.rand[0]), put("e2e_d_u", cap.rand[1])
        put("e2e_d_rgb", rgb), put("e2e_d_disp", disp), put("e2e_d_acc", acc)
        put("e2e_d_rgb0", ex["rgb0"]), put("e2e_d_beta", ex["beta"]), put("e2e_d_z_std", ex["z_std"])
        put("e2e_d_inds", cap.inds[0])

    out = os.path.join(HERE, "render_golden.npz")
    np.savez_compressed(out, **G)
    print("wrote", out, os.path.getsize(out) / 1e3, "KB,", len(G), "arrays")


if __name__ == "__main__":
    main()
