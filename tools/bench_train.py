"""Time one `train_on_batch` step (BASELINE config[3] / SURVEY cfg4 shape) on the GPU.

    python tools/bench_train.py [--H 480 --W 640 --netw 256 --Nc 64 --Nf 128 --levels 0 --steps 10 --warmup 3]   (defaults = SURVEY 8d cfg4)
    torchrun --nproc-per-node 2 tools/bench_train.py ...      (data-parallel: one image per rank, gradient all-reduce)

Prints one JSON line: steps/s over all ranks, ms/step, and a per-phase split measured with CUDA events.
"""
import argparse
import json
import os
import sys
import types

import numpy as np
import torch
import torch.distributed as dist

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root)
sys.path.insert(0, os.path.join(root, "tests"))
from dfnet_b200 import direct_feature_matching as dfm  # noqa: E402
from dfnet_b200 import nerfw, parallel  # noqa: E402
from dfnet_b200 import _lib  # noqa: E402
from dfnet_b200.dfnet import DFNet  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--H", type=int, default=480)
    ap.add_argument("--W", type=int, default=640)
    ap.add_argument("--netw", type=int, default=256)
    ap.add_argument("--Nc", type=int, default=64)
    ap.add_argument("--Nf", type=int, default=128)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--levels", type=int, nargs="+", default=[0])
    ap.add_argument("--no-gc", action="store_true", help="diagnostic: disable Python's cyclic GC during the timed steps")
    a = ap.parse_args()
    rank, world, local = parallel.dist_info()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        sys.stdout.flush()
        saved_fd = os.dup(1)   # NCCL's version banner -> stderr: stdout stays the JSON line
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    torch.manual_seed(0)
    Fnet, Gnet = DFNet().to(dev), DFNet().to(dev).eval()
    with torch.no_grad():
        Fnet.fc_pose.weight.mul_(1e-2)
        Fnet.fc_pose.bias.copy_(torch.tensor([1., 0, 0, 0.1, 0, 1, 0, -0.05, 0, 0, 1, 2.0]))
    for p in Gnet.parameters():
        p.requires_grad_(False)
    Fnet.train()
    for m in Fnet.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.eval()
            m.weight.requires_grad_(False), m.bias.requires_grad_(False)
    c, f, ea, et = [m.to(dev) for m in nerfw.make_synthetic_nerf(D=8, W=a.netw, fine=True)]
    for m in (c, f, ea, et):
        for p in m.parameters():
            p.requires_grad_(False)
    kw = dict(network_query_fn=None, perturb=0.0, N_importance=a.Nf, network_fine=f, N_samples=a.Nc, network_fn=c,
              use_viewdirs=True, white_bkgd=False, raw_noise_std=0.0, embedding_a=ea, embedding_t=et, test_time=True,
              ndc=False, lindisp=False, near=0.0, far=2.5)
    args = types.SimpleNamespace(DFNet=True, preprocess_ImgNet=False, svd_reg=True, combine_loss=True, per_channel=False,
                                 chunk=32768, batch_size=1, combine_loss_w=[0.0, 0.0, 1.0], feature_matching_lvl=a.levels)
    opt = torch.optim.Adam([p for p in Fnet.parameters() if p.requires_grad], lr=1e-5)
    rng = np.random.RandomState(rank)
    data = torch.from_numpy(rng.rand(1, 3, a.H, a.W).astype(np.float32)).pin_memory()
    pose = torch.tensor([[1., 0, 0, 0.1, 0, 1, 0, -0.05, 0, 0, 1, 2.0]])
    hist = torch.tensor([[5., 10, 20, 30, 15, 10, 5, 3, 1, 1]])
    world_setup = dict(pose_scale=0.5, pose_scale2=1.0, move_all_cam_vec=[0.0, 0.0, 0.05])
    hwf = (a.H, a.W, 1.25 * a.W)

    def step():
        return dfm.train_on_batch(args, data, Fnet, Gnet, pose, hist, hwf, opt, True, dev, world_setup, **kw)

    for _ in range(a.warmup):
        loss, psnr = step()
    l0 = C_launches()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    # per-step device times (events around every step; a step ends with the host read of the loss, like the reference):
    # `ms_per_step` is the mean over the timed steps, `ms_median` the median (robust against allocator / clock hiccups)
    if a.no_gc:
        import gc
        gc.collect()
        gc.disable()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.steps):
        evs[i][0].record()
        loss, psnr = step()
        evs[i][1].record()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    per_order = [round(x.elapsed_time(y), 1) for x, y in evs]
    per = sorted(per_order)
    ms_median = per[len(per) // 2]
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"metric": "train_on_batch steps/sec", "value": world * 1e3 / float(t[0]), "unit": "steps/s", "n_gpus": world,
                          "ms_per_step": float(t[0]), "ms_median": ms_median, "ms_min": per[0], "ms_max": per[-1], "ms_steps": per_order, "steps": a.steps, "warmup": a.warmup, "loss": float(loss[0]),
                          "gpu_launches_per_step": (C_launches() - l0) / a.steps,
                          "config": {"workload": f"train_on_batch {a.H}x{a.W}, render {a.H // 4}x{a.W // 4}, NeRF-W 8x{a.netw} "
                                                 f"{a.Nc}+{a.Nf}, DFNet F+G, levels {a.levels}, Adam", "dtype": "f16 fwd / bf16 grad"}}))
    if world > 1:
        dist.destroy_process_group()


def C_launches():
    return int(_lib.lib.dfb_launch_count())


if __name__ == "__main__":
    main()
