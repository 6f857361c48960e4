#!/usr/bin/env python
"""Headline benchmark: NeRF-Hist render throughput (rays/sec) on BASELINE.json config[1]
(7-Scenes-heads-shaped 640x480 image, 64+128 samples, 8x256 NeRF-W networks, test-time
render), one image per step.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--mma f16|bf16|fp32]

Prints ONE JSON line (rank 0).  `value` is device-timed with pose/histogram resident in HBM;
`e2e` goes through dfb_render_image_host (pinned host pose in, pinned host image out).
`--impl reference` times the CPU oracle port of the reference on a bounded ray sample.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W, FOCAL, NEAR, FAR = 480, 640, 585.0, 0.0, 2.5
NC, NF = 64, 128
HIST = np.array([5, 10, 20, 30, 15, 10, 5, 3, 1, 1], np.float32)
# Algorithmic FLOPs: 2*MACs of every Linear on the path, W=256, D=8
F_COARSE, F_FINE = 982528, 1369856          # per sample
# FLOPs the tcgen05 kernel actually issues per fine sample: xyz_encoding_final (2*256*256) is folded into the
# two layers that consume it and the output heads run as two zero-padded N=64 steps: trunk,
# sigma step, dir|transient.0, transient 2/4/6, heads step.  The roofline uses the algorithmic figure above.
F_FINE_EXECUTED = 2 * (64 * 256 + 6 * 256 * 256 + 320 * 256 + 256 * 64 + 256 * 256 + 3 * 128 * 128 + 256 * 64)
FLOP_PER_RAY = NC * F_COARSE + (NC + NF) * F_FINE


WORKLOADS = {
    # name -> (H, W, focal, near, far, Nc, Nf, netwidth, label)
    "cfg2": (480, 640, 585.0, 0.0, 2.5, 64, 128, 256, "BASELINE config[1]: 640x480 7-Scenes-heads-shaped image, 64+128 samples"),
    "cfg5": (1080, 1920, 1674.0, 0.0, 20.0, 64, 192, 256, "BASELINE config[4]: 1920x1080 Cambridge-ShopFacade-shaped image, 64+192 samples"),
    # what the reference's own config files run (models/options.py defaults: netwidth 128, 64+64 samples); the
    # tcgen05 kernels execute it embedded in 8x256 with zero weights, so `fine_executed_tflops` is ~4x `achieved`
    "shipped": (480, 640, 585.0, 0.0, 2.5, 64, 64, 128, "reference defaults (config_nerfh.txt): 640x480, 64+64 samples, netwidth 128"),
}
LABEL = WORKLOADS["cfg2"][8]
NETW = 256


def mlp_flops(w, a=50):
    """Algorithmic FLOPs per sample (SURVEY 8d formulas, D = 8): (coarse sigma-only, fine full)."""
    trunk = 63 * w + 7 * w * w + 63 * w
    fine = trunk + w + w * w + (w + 27 + a) * w // 2 + 3 * w // 2 + (w + 20) * w // 2 + 3 * (w // 2) ** 2 + 5 * w // 2
    return 2 * (trunk + w), 2 * fine


def set_workload(name):
    global H, W, FOCAL, NEAR, FAR, NC, NF, FLOP_PER_RAY, LABEL, NETW, F_COARSE, F_FINE
    H, W, FOCAL, NEAR, FAR, NC, NF, NETW, LABEL = WORKLOADS[name]
    F_COARSE, F_FINE = mlp_flops(NETW)
    FLOP_PER_RAY = NC * F_COARSE + (NC + NF) * F_FINE


def pose(i):
    rng = np.random.RandomState(100 + i)
    ax = rng.randn(3)
    ax /= np.linalg.norm(ax)
    ang = np.deg2rad(10.0) * rng.rand()
    K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
    R = np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K
    return np.concatenate([R, np.array([[0.0], [0.0], [1.0]])], 1).astype(np.float32)


def ncu_traffic_per_launch(avg_rays_per_launch):
    """dram__bytes_read.sum + dram__bytes_write.sum of the fine tcgen05 MLP kernel (compositing fused into its heads
    epilogue) from the committed `ncu --set full` capture (profiles/r02_ncu_mlp_tc_summary.csv: one 65 536-ray chunk of a
    640x480 image), scaled to this run's average launch."""
    p = os.path.join(ROOT, "profiles", "r02_ncu_mlp_tc_summary.csv")
    try:
        rows = {r.split(",")[0]: r.strip().split(",") for r in open(p)}
        mb = sum(float(rows[k][3].strip('"')) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        return mb * 1e6 * avg_rays_per_launch / 65536.0
    except Exception:  # noqa: BLE001
        return None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(gpu)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1])), mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for nm, v in zip(names, r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def _cpu_inputs(n_rays):
    from dfnet_b200 import nerfw
    from oracle import nerf_oracle as O
    mods = nerfw.make_synthetic_nerf(D=8, W=NETW)
    o, d = O.get_rays(H, W, FOCAL, pose(0))
    sel = np.linspace(0, H * W - 1, n_rays).astype(np.int64)
    return mods, o.reshape(-1, 3)[sel], d.reshape(-1, 3)[sel]


class CpuReference:
    """The reference's own CPU implementation of the path on a bounded ray sample of the same image, all host threads.
    kind "reference": the UNMODIFIED reference (models.rendering.render from baseline/_ref, see baseline/make_ref.py);
    kind "port": the numpy oracle port with its Linear layers on torch-CPU addmm, when baseline/_ref is absent."""

    def __init__(self, n_rays):
        import torch
        torch.set_num_threads(os.cpu_count())
        self.n_rays = n_rays
        self.mods, self.o, self.d = _cpu_inputs(n_rays)
        sys.path.insert(0, ROOT)
        from baseline import ref_shims
        self.kind = "reference" if ref_shims.available() else "port"
        if self.kind == "reference":
            from baseline import ref_runner
            self.kw = ref_runner.reference_render_kwargs(self.mods, test_time=True, N_samples=NC, N_importance=NF)
            self.run = lambda n: ref_runner.reference_render_rays(
                self.kw, torch.from_numpy(self.o[:n]), torch.from_numpy(self.d[:n]), NEAR, FAR, torch.from_numpy(HIST[None]))
            self.what = ("unmodified reference models.rendering.render (baseline/_ref), torch-CPU fp32, chunk 32768, "
                         "netchunk 65536")
        else:
            from oracle import nerf_oracle as O
            O.set_linear_backend("torch")  # Linear layers through torch-CPU addmm, like the reference
            m = self.mods
            self.nets = dict(coarse={k: v.numpy() for k, v in m[0].state_dict().items()},
                             fine={k: v.numpy() for k, v in m[1].state_dict().items()},
                             emb_a=m[2].weight.detach().numpy(), emb_t=m[3].weight.detach().numpy(), D=8, skips=(4,))
            self.run = lambda n: O.render_rays(O.make_ray_records(self.o[:n], self.d[:n], NEAR, FAR, HIST[None]), self.nets,
                                               NC, NF, test_time=True)
            self.what = "oracle port of render_rays, Linear layers via torch-CPU addmm"

    def time(self, n=None):
        n = n or self.n_rays
        t0 = time.perf_counter()
        self.run(n)
        dt = time.perf_counter() - t0
        return n / dt, dt

    def sample(self, dt=None):
        return (f"{self.n_rays} rays spread over the same {W}x{H} image" + (f", {dt:.1f} s" if dt else " per step") +
                f" ({self.what}; {NC}+{NF} samples, 8x{NETW} NeRF-W, {os.cpu_count()} threads)")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ref = CpuReference(args.cpu_rays)
    for _ in range(args.warmup):
        ref.time(min(args.cpu_rays, 512))
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ref.time()
    dt = time.perf_counter() - t0
    v = args.cpu_rays * args.steps / dt
    line = {"impl": "reference", "metric": "rays/sec", "value": v, "unit": "rays/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_cfg(args), "gpu_launches": 0,
            "cpu_baseline": {"value": v, "unit": "rays/s", "cores": os.cpu_count(), "kind": ref.kind, "sample": ref.sample()},
            "e2e": {"value": v, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_cfg(args):
    # identical for both arms (the driver compares the dicts): what the workload IS, not how an arm computes it
    return {"workload": f"{LABEL}, 8x{NETW} NeRF-W coarse+fine, test-time render_path step (1 image = {H * W} rays per step)",
            "H": H, "W": W, "N_samples": NC, "N_importance": NF, "netdepth": 8, "netwidth": NETW,
            "parallelism": f"images sharded over {args.gpus} rank(s), no collective",
            "l2": "per-chunk working set (~0.65 GB of intermediates per 65 536 rays) exceeds the 126 MB L2; no extra flush"}


def _event_timed(fn, steps, warmup, stream, barrier, after_warmup=None):
    """(total ms, per-step ms list) of `steps` calls after `warmup`, CUDA events on `stream`."""
    import torch
    for i in range(warmup):
        fn(i)
    barrier()
    if after_warmup is not None:
        after_warmup()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    evs[0].record(stream)
    for i in range(steps):
        fn(warmup + i)
        evs[i + 1].record(stream)
    barrier()
    per = [evs[i].elapsed_time(evs[i + 1]) for i in range(steps)]
    return evs[0].elapsed_time(evs[steps]), per


def bench_extras(args, dev, rank, world, mma):
    """BASELINE configs [2], [3], [4] next to the headline, each through the repo's public Python API:
      dfnet : DFNet siamese forward on a 640x480 target/render pair + level-0 cosine feature loss (1 GPU per rank)
      train : train_on_batch (SURVEY cfg4: 480x640 image, 120x160 render, 8x256 NeRF-W 64+128, DFNet F+G, Adam), one
              image per rank and ONE all-reduce of the pose regressor's gradients when WORLD_SIZE > 1
      cfg5  : one 1920x1080 image, 64+192 samples, per rank.
    Every number is max-over-ranks device time; values are whole-job aggregates."""
    import types
    import torch
    import torch.distributed as dist
    from dfnet_b200 import _lib, nerfw, ops, parallel
    from dfnet_b200 import direct_feature_matching as dfm
    from dfnet_b200.dfnet import DFNet, feature_loss
    lib = _lib.lib
    stream = torch.cuda.current_stream()
    pk, pk_kind = peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    out = {}
    # ---- config[2]: DFNet pair + feature-matching loss ------------------------------------------------------------
    torch.manual_seed(1234)
    net = DFNet().to(dev).eval()
    x_h = torch.rand(2, 3, 480, 640).pin_memory()
    x = x_h.to(dev)

    @torch.no_grad()
    def dfnet_step(i):
        feats, _ = net(x, return_feature=True, isSingleStream=False, return_pose=False, upsampleH=480, upsampleW=640)
        return feature_loss(feats[1][0, 0], feats[0][0, 0])

    @torch.no_grad()
    def dfnet_step_e2e(i):
        xd = x_h.to(dev, non_blocking=True)
        feats, _ = net(xd, return_feature=True, isSingleStream=False, return_pose=False, upsampleH=480, upsampleW=640)
        return float(feature_loss(feats[1][0, 0], feats[0][0, 0]))   # host read of the loss every step

    l0 = lib.dfb_launch_count()
    ms, _ = _event_timed(dfnet_step, 10, 3, stream, barrier)
    launches = (lib.dfb_launch_count() - l0) / 13
    ms = max_ranks(ms) / 10
    ms_e2e, _ = _event_timed(dfnet_step_e2e, 10, 3, stream, barrier)
    ms_e2e = max_ranks(ms_e2e) / 10
    fa = torch.randn(128, 480 * 640, device=dev)
    fb = torch.randn(128, 480 * 640, device=dev)
    lms, _ = _event_timed(lambda i: feature_loss(fa, fb), 20, 3, stream, barrier)
    lms /= 20
    del fa, fb
    conv_flop = 2 * 325.3e9          # 325.3 GFLOP per 480x640 image, two images
    loss_bytes = 2 * 128 * 480 * 640 * 4
    peak_t = pk.get("bf16_tflops_sustained", pk.get("bf16_tflops"))
    out["dfnet"] = {
        "metric": "DFNet 640x480 pairs/sec (siamese forward, 3 levels upsampled + level-0 cosine loss)",
        "value": world * 1e3 / ms, "unit": "pairs/s", "ms_per_pair": ms, "gpu_launches_per_pair": launches,
        "e2e": {"value": world * 1e3 / ms_e2e, "unit": "pairs/s", "h2d_bytes_per_step": 2 * 3 * 480 * 640 * 4, "d2h_bytes_per_step": 4},
        "roofline": {"bound": "tensor", "kernel": "k_conv_tc (whole forward: 13 encoder + 6 head convolutions)",
                     "achieved": conv_flop / (ms * 1e-3) / 1e12, "peak": peak_t, "unit": "TFLOP/s",
                     "frac": conv_flop / (ms * 1e-3) / 1e12 / peak_t, "traffic": None,
                     "note": "algorithmic 325.3 GFLOP per image (SURVEY 8a) x 2 images / time of the whole pair step"},
        "loss_roofline": {"bound": "hbm", "kernel": "k_cosine (dfb_cosine_loss, level 0: 2 x [128, 307200] fp32 read once)",
                          "achieved": loss_bytes / (lms * 1e-3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                          "frac": loss_bytes / (lms * 1e-3) / 1e9 / pk["hbm_gbs"], "ms": lms, "traffic": None},
        "dtype": "f16 operands / fp32 accumulate"}
    del net, x
    # ---- config[3]: train_on_batch, data-parallel over ranks -------------------------------------------------------
    torch.manual_seed(0)
    Fnet, Gnet = DFNet().to(dev), DFNet().to(dev).eval()
    with torch.no_grad():
        Fnet.fc_pose.weight.mul_(1e-2)
        Fnet.fc_pose.bias.copy_(torch.tensor([1., 0, 0, 0.1, 0, 1, 0, -0.05, 0, 0, 1, 2.0]))
    for p in Gnet.parameters():
        p.requires_grad_(False)
    Fnet.train()
    for m in Fnet.modules():   # freeze_bn_layer_train (reference feature/direct_feature_matching.py:52-61, train.py:111-112)
        if isinstance(m, torch.nn.BatchNorm2d):
            m.eval()
            m.weight.requires_grad_(False), m.bias.requires_grad_(False)
    c, f, ea, et = [m.to(dev) for m in nerfw.make_synthetic_nerf(D=8, W=256, fine=True)]
    for m in (c, f, ea, et):
        for p in m.parameters():
            p.requires_grad_(False)
    kw = dict(network_query_fn=None, perturb=0.0, N_importance=128, network_fine=f, N_samples=64, network_fn=c,
              use_viewdirs=True, white_bkgd=False, raw_noise_std=0.0, embedding_a=ea, embedding_t=et, test_time=True,
              ndc=False, lindisp=False, near=0.0, far=2.5, mma=mma)
    targs = types.SimpleNamespace(DFNet=True, preprocess_ImgNet=False, svd_reg=True, combine_loss=True, per_channel=False,
                                  chunk=32768, batch_size=1, combine_loss_w=[0.0, 0.0, 1.0], feature_matching_lvl=[0])
    opt = torch.optim.Adam([p for p in Fnet.parameters() if p.requires_grad], lr=1e-5)
    rng = np.random.RandomState(rank)
    data = torch.from_numpy(rng.rand(1, 3, 480, 640).astype(np.float32)).pin_memory()
    tpose = torch.tensor([[1., 0, 0, 0.1, 0, 1, 0, -0.05, 0, 0, 1, 2.0]])
    thist = torch.tensor(HIST[None])
    world_setup = dict(pose_scale=0.5, pose_scale2=1.0, move_all_cam_vec=[0.0, 0.0, 0.05])

    def train_step(i):
        # host image in (pinned), host loss / PSNR out every step: this IS the end-to-end call (train.py's inner loop)
        return dfm.train_on_batch(targs, data, Fnet, Gnet, tpose, thist, (480, 640, 585.0), opt, True, dev, world_setup, **kw)

    l0 = lib.dfb_launch_count()
    n_tr = 10
    # the all-reduce statistics cover the timed steps only (the first collective of a process sets up NCCL's channels)
    tot, per = _event_timed(train_step, n_tr, 4, stream, barrier, after_warmup=lambda: parallel.allreduce_stats(reset=True))
    launches = (lib.dfb_launch_count() - l0) / (n_tr + 4)
    ar = parallel.allreduce_stats(reset=True)
    ms_mean = max_ranks(tot) / n_tr
    ms_med = max_ranks(float(np.median(per)))
    # the same step with the feature net evaluating ALL three hyper-column levels, as the reference does before it
    # index_selects feature_matching_lvl (the default step skips the levels nothing reads: identical loss and gradients)
    os.environ["DFB_ALL_LEVELS"] = "1"
    tot_all, _ = _event_timed(train_step, 5, 2, stream, barrier)
    del os.environ["DFB_ALL_LEVELS"]
    ms_all = max_ranks(tot_all) / 5
    out["train"] = {
        "metric": "train_on_batch steps/sec (BASELINE config[3] / SURVEY cfg4: 480x640 image, 120x160 render, 8x256 NeRF-W 64+128, "
                  "DFNet F+G, feature_matching_lvl=[0], Adam), one image per rank",
        "value": world * 1e3 / ms_mean, "unit": "steps/s", "ms_per_step": ms_mean, "ms_median": ms_med, "steps": n_tr, "warmup": 4,
        "ms_per_step_all_levels": ms_all,
        "note": "the frozen feature net evaluates only feature_matching_lvl (with [0] its encoder stops after conv1_2); loss, PSNR "
                "and every gradient are identical to evaluating all three levels and selecting afterwards, which is what the "
                "reference does and what ms_per_step_all_levels times",
        "gpu_launches_per_step": launches, "n_gpus": world, "scaling": "weak",
        "allreduce": {"calls_per_step": ar["calls"] / max(n_tr, 1), "bytes_per_call": ar["bytes_per_call"],
                      "ms_per_call": ar["ms"] / max(ar["timed_calls"], 1) if ar["timed_calls"] else 0.0,
                      "exposed_ms_per_step": ar["exposed_ms"] / max(ar["timed_calls"], 1) if ar["timed_calls"] else 0.0,
                      "note": "ONE NCCL all-reduce of the pose regressor's flat fp32 gradient bucket per step, issued on a side "
                              "stream as soon as the bucket is complete; ms_per_call = device time of the collective, "
                              "exposed = time the main stream waited for it"},
        "e2e": {"value": world * 1e3 / ms_mean, "unit": "steps/s", "h2d_bytes_per_step": 3 * 480 * 640 * 4 + 12 * 4 + 10 * 4,
                "d2h_bytes_per_step": 8},
        "dtype": f"render {mma} operands / fp32 accumulate; DFNet fp16 forward, bf16 gradients, fp32 weight gradients"}
    del Fnet, Gnet, opt
    # ---- NeRF-Hist training step: the reference's defaults (N_rand 1536, 64+64 samples, netwidth 128) ------
    from dfnet_b200 import nerf_train
    from dfnet_b200.losses import loss_dict
    tmods = [m.to(dev) for m in nerfw.make_synthetic_nerf(D=8, W=128, fine=True)]
    tparams = [p for m in tmods for p in m.parameters()]
    for p in tparams:
        p.requires_grad_(True)
    nargs = types.SimpleNamespace(chunk=32768, lrate=5e-4, lrate_decay=250)
    nopt = torch.optim.Adam(tparams, lr=nargs.lrate, betas=(0.9, 0.999))
    nkw = dict(network_query_fn=None, perturb=1.0, N_importance=64, network_fine=tmods[1], N_samples=64, network_fn=tmods[0],
               use_viewdirs=True, white_bkgd=False, raw_noise_std=0.0, embedding_a=tmods[2], embedding_t=tmods[3], test_time=False,
               ndc=False, lindisp=False)
    nimg = torch.rand(3, 120, 160).pin_memory()
    nloss = loss_dict["nerfw"](coef=1)
    np.random.seed(rank)

    def nerf_step(i):
        # host image / pose in, host loss out: the loop body of train_on_epoch_nerfw (run_nerf.py:33-77)
        loss, psnr = nerf_train.train_on_batch_nerfw(nargs, nimg, tpose[0], thist, 120, 160, 146.0, 1536, nopt, nloss, i, nkw,
                                                     near=0.0, far=2.5)
        return float(loss)

    l0 = lib.dfb_launch_count()
    totn, pern = _event_timed(nerf_step, 10, 4, stream, barrier)
    out["nerf_train"] = {
        "metric": "NeRF-Hist training steps/sec (reference defaults: N_rand 1536 rays, 64+64 samples, 8x128 NeRF-W coarse+fine, "
                  "NerfWLoss, Adam), one image per rank, no collective (the reference trains on one GPU)",
        "value": world * 1e3 / (max_ranks(totn) / 10), "unit": "steps/s", "ms_per_step": max_ranks(totn) / 10,
        "ms_median": float(np.median(pern)), "gpu_launches_per_step": (lib.dfb_launch_count() - l0) / 14,
        "rays_per_sec": world * 1536 * 1e3 / (max_ranks(totn) / 10),
        "dtype": "fp16 activations, bf16 gradients, fp32 accumulation and weight gradients"}
    del tmods, tparams, nopt
    # ---- the other operand kinds / options on the headline workload ---------------------------------------------------
    hq = ops.NerfHandle(c, f, ea, et)
    c2w0 = torch.tensor(pose(rank * 1000), device=dev)
    hist0 = torch.tensor(HIST, device=dev)
    base_rgb = hq.render(NC, NF, True, c2w=c2w0, H=H, W=W, focal=FOCAL, near=NEAR, far=FAR, hist=hist0, mma=mma)["rgb"].clone()
    for key, kwv, what in (("split_coarse", dict(mma="f16s"), "mma f16s: split-precision (hi+lo fp16, 3 MMA sub-steps) coarse pass + fp16 fine pass"),
                           ("ert", dict(mma=mma, ert_eps=1e-3), "opt-in early ray termination, ert_eps 1e-3")):
        msq, _ = _event_timed(lambda i: hq.render(NC, NF, True, c2w=c2w0, H=H, W=W, focal=FOCAL, near=NEAR, far=FAR, hist=hist0, **kwv),
                              4, 3, stream, barrier)
        oq = hq.render(NC, NF, True, c2w=c2w0, H=H, W=W, focal=FOCAL, near=NEAR, far=FAR, hist=hist0, **kwv)
        out[key] = {"what": what, "value": world * H * W / (max_ranks(msq) / 4 * 1e-3), "unit": "rays/s", "ms_per_image": max_ranks(msq) / 4,
                    "max_abs_rgb_diff_vs_headline": float((oq["rgb"] - base_rgb).abs().max())}
    del hq
    # ---- config[4]: the 1920x1080, 64+192 image -------------------------------------------------------------------
    Hc, Wc, fc, nearc, farc, Ncc, Nfc = WORKLOADS["cfg5"][:7]
    h = ops.NerfHandle(c, f, ea, et)
    c2w = torch.tensor(pose(rank * 1000), device=dev)
    hist_d = torch.tensor(HIST, device=dev)

    def cfg5_step(i):
        return h.render(Ncc, Nfc, True, c2w=c2w, H=Hc, W=Wc, focal=fc, near=nearc, far=farc, hist=hist_d, mma=mma)

    lib.dfb_profile_enable(1)
    ms5, _ = _event_timed(cfg5_step, 3, 3, stream, barrier)
    lib.dfb_profile_enable(0)
    cm, fm, cl, fl = C.c_double(), C.c_double(), C.c_int64(), C.c_int64()
    lib.dfb_profile_read(C.byref(cm), C.byref(fm), C.byref(cl), C.byref(fl))
    ms5 = max_ranks(ms5) / 3
    f_c, f_f = mlp_flops(256)
    fine_tf = (Hc * Wc * 6 * (Ncc + Nfc) * f_f) / (fm.value * 1e-3) / 1e12 if fm.value > 0 else 0.0
    out["cfg5"] = {
        "metric": "rays/sec, BASELINE config[4] shape (1920x1080 image, 64+192 samples, 8x256 NeRF-W), one image per rank",
        "value": world * Hc * Wc / (ms5 * 1e-3), "unit": "rays/s", "ms_per_image": ms5, "n_gpus": world, "scaling": "weak",
        "images_per_sec": world * 1e3 / ms5,
        "roofline": {"bound": "tensor", "kernel": "fine NeRF-W MLP", "achieved": fine_tf, "peak": peak_t, "unit": "TFLOP/s",
                     "frac": fine_tf / peak_t, "traffic": None,
                     "whole_step_tflops": Hc * Wc * (Ncc * f_c + (Ncc + Nfc) * f_f) / (ms5 * 1e-3) / 1e12}}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mma", default=os.environ.get("DFB_MMA", "auto"), choices=["auto", "f16", "f16s", "bf16", "fp32"])
    ap.add_argument("--cpu-rays", type=int, default=8192)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the BASELINE config[2]/[3]/[4] measurements (keys dfnet, train, cfg5 of the JSON line)")
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS),
                    help="cfg2 = BASELINE config[1] (the headline, default); cfg5 = BASELINE config[4] (1920x1080, 64+192); "
                         "shipped = the reference's own defaults (netwidth 128, 64+64)")
    args = ap.parse_args()
    set_workload(args.workload)
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from dfnet_b200 import _lib, nerfw, ops
    lib = _lib.lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # stdout carries exactly ONE JSON line: whatever NCCL prints while the communicator comes up (its version
        # banner) is sent to stderr by pointing file descriptor 1 at stderr for the duration of the initialisation
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            os.dup2(saved_fd, 1)
            os.close(saved_fd)

    mods = nerfw.make_synthetic_nerf(D=8, W=NETW)
    h = ops.NerfHandle(*[m.to(dev) for m in mods])
    mma = args.mma
    if mma == "auto":
        mma = "f16"
    cfg = _lib.RenderCfg(N_samples=NC, N_importance=NF, test_time=1, perturb=0, mma_kind=_lib.MMA_KINDS[mma],
                         lindisp=0, raw_noise_std=0.0, hist_len=len(HIST))
    N = H * W
    hist_d = torch.tensor(HIST, device=dev)
    rgb = torch.empty(N, 3, device=dev)
    disp = torch.empty(N, device=dev)
    acc = torch.empty(N, device=dev)
    stage = 512 + (N * 5 * 4 + 255) // 256 * 256
    ws, ws_bytes = h.workspace(cfg, N, dev, extra_bytes=stage)
    stream = torch.cuda.current_stream()
    sp = C.c_void_p(stream.cuda_stream)

    def step_device(i):
        c2w = poses_d[i % len(poses_d)]
        _lib.check(lib.dfb_render_fwd(h._h, C.byref(cfg), None, C.c_void_p(c2w.data_ptr()), H, W, FOCAL, NEAR, FAR,
                                      C.c_void_p(hist_d.data_ptr()), N, None, None, None, C.c_void_p(rgb.data_ptr()),
                                      C.c_void_p(disp.data_ptr()), C.c_void_p(acc.data_ptr()), None,
                                      C.c_void_p(ws.data_ptr()), ws_bytes, sp))

    n_poses = 8
    poses_h = [torch.tensor(pose(rank * 1000 + i)).pin_memory() for i in range(n_poses)]
    poses_d = [p.to(dev) for p in poses_h]
    hist_h = torch.tensor(HIST).pin_memory()
    rgb_h = torch.empty(N, 3).pin_memory()
    disp_h = torch.empty(N).pin_memory()
    acc_h = torch.empty(N).pin_memory()

    def step_host(i):
        _lib.check(lib.dfb_render_image_host(h._h, C.byref(cfg), C.c_void_p(poses_h[i % n_poses].data_ptr()), H, W,
                                             FOCAL, NEAR, FAR, C.c_void_p(hist_h.data_ptr()),
                                             C.c_void_p(rgb_h.data_ptr()), C.c_void_p(disp_h.data_ptr()),
                                             C.c_void_p(acc_h.data_ptr()), C.c_void_p(ws.data_ptr()), ws.numel(), sp))
        stream.synchronize()  # render_path consumes the image on the host every step (rendering.py:423)

    step_device(0)
    args.mma = mma
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, profile=False):
        for i in range(warmup):
            fn(i)
        barrier()
        if profile:
            lib.dfb_profile_enable(1)
        l0 = lib.dfb_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(steps):
            fn(warmup + i)
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        launches = lib.dfb_launch_count() - l0
        prof = None
        if profile:
            lib.dfb_profile_enable(0)
            cm, fm, cl, fl = C.c_double(), C.c_double(), C.c_int64(), C.c_int64()
            lib.dfb_profile_read(C.byref(cm), C.byref(fm), C.byref(cl), C.byref(fl))
            prof = (cm.value, fm.value, cl.value, fl.value)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches, prof

    sampler = ClockSampler(local) if rank == 0 else None
    ms, launches, prof = timed(step_device, args.steps, args.warmup, profile=True)
    clocks = sampler.stop() if sampler else None
    ms_e2e, _, _ = timed(step_host, args.steps, args.warmup)
    bad = sorted(set(clocks["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}) if clocks else []
    if bad and rank == 0:  # re-measure once
        sampler = ClockSampler(local)
        ms, launches, prof = timed(step_device, args.steps, args.warmup, profile=True)
        clocks = sampler.stop()
        clocks["remeasured_after"] = bad

    if rank == 0:
        pk, pk_kind = peaks()
        total_rays = N * args.steps * world
        value = total_rays / (ms * 1e-3)
        coarse_ms, fine_ms, cl, fl = prof
        # dominant kernel = fine-network MLP; per launch = one internal chunk of rays
        fine_flops = N * args.steps * (NC + NF) * F_FINE
        achieved = fine_flops / (fine_ms * 1e-3) / 1e12 if fine_ms > 0 else 0.0
        peak = pk.get("bf16_tflops_sustained", pk.get("bf16_tflops"))
        roof = {"bound": "tensor", "kernel": "fine NeRF-W MLP (k_mlp_*), %d launches, %.3f ms avg" % (fl, fine_ms / max(fl, 1)),
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "peak_source": f"{pk_kind} bf16_tflops_sustained (kernel timed inside a long step)",
                "traffic": ncu_traffic_per_launch(N * args.steps / max(fl, 1)) if mma != "fp32" and args.workload == "cfg2" else None,
                "traffic_note": "DRAM bytes per launch from profiles/r02_ncu_mlp_tc_summary.csv (ncu --set full, 65 536-ray "
                                "launch) scaled to this run's average launch: depths, per-ray bias and ray records in, 32-byte "
                                "partial records out (compositing is fused; the [P,9] raw tensor is no longer written)",
                "kernel_share_of_step": (fine_ms + coarse_ms) / ms,
                "coarse_mlp_tflops": (N * args.steps * NC * F_COARSE) / (coarse_ms * 1e-3) / 1e12 if coarse_ms > 0 else 0.0,
                "whole_step_tflops": value * FLOP_PER_RAY / 1e12,
                "fine_executed_tflops": (achieved * F_FINE_EXECUTED / F_FINE) if mma != "fp32" and
                os.environ.get("DFB_TC_FOLD_FINAL", "1") != "0" else achieved,
                "note": "achieved = algorithmic FLOPs (SURVEY 8d: 1,369,856 per fine sample) / CUDA-event time of the fine "
                        "MLP launches; fine_executed_tflops counts the layers actually issued"}
        line = {"metric": "rays/sec", "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": {"f16": "f16", "f16s": "f16", "bf16": "bf16", "fp32": "f32"}[mma], "data": "synthetic",
                "config": workload_cfg(args), "images_per_sec": world * args.steps / (ms * 1e-3),
                "e2e": {"value": total_rays / (ms_e2e * 1e-3), "unit": "rays/s", "h2d_bytes_per_step": 12 * 4 + 10 * 4,
                        "d2h_bytes_per_step": N * 5 * 4, "images_per_sec": world * args.steps / (ms_e2e * 1e-3)},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roof}
        line["mma"] = mma
    extras = {}
    if not args.no_extras and args.workload == "cfg2":
        extras = bench_extras(args, dev, rank, world, mma)
    if rank == 0:
        # the CPU baseline runs LAST: torch's intra-op thread pool keeps spinning after the reference's CPU matmuls and
        # slowed every host-bound step measured after it (train_on_batch: 30 ms instead of 20 ms per step)
        if world == 1 and not args.no_cpu_baseline:
            ref = CpuReference(args.cpu_rays)
            ref.time(256)
            v, dt = ref.time()
            line["cpu_baseline"] = {"value": v, "unit": "rays/s", "cores": os.cpu_count(), "kind": ref.kind, "sample": ref.sample(dt)}
        line.update(extras)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
