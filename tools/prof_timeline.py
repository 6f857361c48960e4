"""GPU timeline of the bench's train_on_batch step (and the DFNet pair) through torch.profiler / CUPTI: per-kernel device
time, busy time and idle gaps per step.  Answers "is the step bound by kernels or by the host issuing them".
Usage: python tools/prof_timeline.py [train|dfnet|nerf] [steps]"""
import collections
import json
import os
import sys
import tempfile
import types

import numpy as np
import torch

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root)
from dfnet_b200 import direct_feature_matching as dfm  # noqa: E402
from dfnet_b200 import nerfw  # noqa: E402
from dfnet_b200.dfnet import DFNet, feature_loss  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "train"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
dev = torch.device("cuda:0")
torch.manual_seed(0)

if what == "train":
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
    c, f, ea, et = [m.to(dev) for m in nerfw.make_synthetic_nerf(D=8, W=256, fine=True)]
    for m in (c, f, ea, et):
        for p in m.parameters():
            p.requires_grad_(False)
    kw = dict(network_query_fn=None, perturb=0.0, N_importance=128, network_fine=f, N_samples=64, network_fn=c, use_viewdirs=True,
              white_bkgd=False, raw_noise_std=0.0, embedding_a=ea, embedding_t=et, test_time=True, ndc=False, lindisp=False,
              near=0.0, far=2.5, mma=os.environ.get("MMA", "f16"))
    args = types.SimpleNamespace(DFNet=True, preprocess_ImgNet=False, svd_reg=True, combine_loss=True, per_channel=False, chunk=32768,
                                 batch_size=1, combine_loss_w=[0.0, 0.0, 1.0], feature_matching_lvl=[0])
    opt = torch.optim.Adam([p for p in Fnet.parameters() if p.requires_grad], lr=1e-5)
    data = torch.from_numpy(np.random.RandomState(0).rand(1, 3, 480, 640).astype(np.float32)).pin_memory()
    pose = torch.tensor([[1., 0, 0, 0.1, 0, 1, 0, -0.05, 0, 0, 1, 2.0]])
    hist = torch.tensor([[5., 10, 20, 30, 15, 10, 5, 3, 1, 1]])
    ws = dict(pose_scale=0.5, pose_scale2=1.0, move_all_cam_vec=[0.0, 0.0, 0.05])

    def step(i):
        return dfm.train_on_batch(args, data, Fnet, Gnet, pose, hist, (480, 640, 585.0), opt, True, dev, ws, **kw)
elif what == "dfnet":
    net = DFNet().to(dev).eval()
    x = torch.rand(2, 3, 480, 640, device=dev)

    @torch.no_grad()
    def step(i):
        feats, _ = net(x, return_feature=True, isSingleStream=False, return_pose=False, upsampleH=480, upsampleW=640)
        return feature_loss(feats[1][0, 0], feats[0][0, 0])
elif what == "nerf":
    from dfnet_b200 import nerf_train
    from dfnet_b200.losses import loss_dict
    tm = [m.to(dev) for m in nerfw.make_synthetic_nerf(D=8, W=128, fine=True)]
    tp = [p for m in tm for p in m.parameters()]
    for p in tp:
        p.requires_grad_(True)
    nargs = types.SimpleNamespace(chunk=32768, lrate=5e-4, lrate_decay=250)
    nopt = torch.optim.Adam(tp, lr=nargs.lrate, betas=(0.9, 0.999))
    nkw = dict(network_query_fn=None, perturb=1.0, N_importance=64, network_fine=tm[1], N_samples=64, network_fn=tm[0],
               use_viewdirs=True, white_bkgd=False, raw_noise_std=0.0, embedding_a=tm[2], embedding_t=tm[3], test_time=False,
               ndc=False, lindisp=False)
    nimg = torch.rand(3, 120, 160).pin_memory()
    nloss = loss_dict["nerfw"](coef=1)
    npose = torch.tensor([1., 0, 0, 0.1, 0, 1, 0, -0.05, 0, 0, 1, 2.0])
    nhist = torch.tensor([[5., 10, 20, 30, 15, 10, 5, 3, 1, 1]])

    def step(i):
        loss, psnr = nerf_train.train_on_batch_nerfw(nargs, nimg, npose, nhist, 120, 160, 146.0, 1536, nopt, nloss, i, nkw,
                                                     near=0.0, far=2.5)
        return float(loss)
else:
    raise SystemExit("unknown workload")

for i in range(6):
    step(i)
torch.cuda.synchronize()
if os.environ.get("NOPROF") == "1":     # plain run of the same steps, e.g. under ncu (which cannot share CUPTI with torch.profiler)
    for i in range(steps):
        step(i)
    torch.cuda.synchronize()
    raise SystemExit(0)
marks = []
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CPU, torch.profiler.ProfilerActivity.CUDA]) as prof:
    for i in range(steps):
        with torch.profiler.record_function(f"STEP{i}"):
            step(i)
            torch.cuda.synchronize()
path = os.path.join(tempfile.gettempdir(), "dfb_trace.json")
prof.export_chrome_trace(path)
tr = json.load(open(path))["traceEvents"]
kern = [e for e in tr if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e]
stepsev = sorted([e for e in tr if e.get("name", "").startswith("STEP") and e.get("cat") in ("user_annotation", "cpu_op")],
                 key=lambda e: e["ts"])
seen = {}
for e in stepsev:
    seen.setdefault(e["name"], e)
for name, se in sorted(seen.items()):
    t0, t1 = se["ts"], se["ts"] + se["dur"]
    ks = sorted([k for k in kern if t0 <= k["ts"] <= t1], key=lambda k: k["ts"])
    if not ks:
        continue
    # union of kernel intervals (several streams)
    busy, cur_s, cur_e = 0.0, None, None
    gaps = []
    for k in ks:
        s, e = k["ts"], k["ts"] + k["dur"]
        if cur_e is None:
            cur_s, cur_e = s, e
        elif s <= cur_e:
            cur_e = max(cur_e, e)
        else:
            busy += cur_e - cur_s
            gaps.append((s - cur_e, k["name"][:60], cur_e, s))
            cur_s, cur_e = s, e
    busy += cur_e - cur_s
    span = ks[-1]["ts"] + ks[-1]["dur"] - ks[0]["ts"]
    print(f"{name}: host span {se['dur'] / 1e3:.2f} ms, first->last kernel {span / 1e3:.2f} ms, GPU busy {busy / 1e3:.2f} ms, "
          f"idle {(span - busy) / 1e3:.2f} ms in {len(gaps)} gaps, {len(ks)} device ops, lead-in {(ks[0]['ts'] - t0) / 1e3:.2f} ms")
    if name == sorted(seen)[-1]:
        cpu = [e for e in tr if e.get("cat") in ("cpu_op", "cuda_runtime", "cuda_driver", "user_annotation") and "dur" in e]
        print("  largest gaps (us, kernel after the gap | host activity overlapping the gap, longest first):")
        for gdur, nm, g0, g1 in sorted(gaps, reverse=True)[:14]:
            ov = sorted([(min(e["ts"] + e["dur"], g1) - max(e["ts"], g0), e["name"][:48]) for e in cpu
                         if e["ts"] < g1 and e["ts"] + e["dur"] > g0 and not e["name"].startswith("STEP")], reverse=True)[:5]
            print(f"    {gdur:7.1f}  {nm:60s} | " + ", ".join(f"{n} {d:.0f}" for d, n in ov))
        agg = collections.defaultdict(lambda: [0, 0.0])
        for k in ks:
            nm = k["name"].split("(")[0][:70] if "dfb::" in k["name"] else k["name"][:160]
            agg[nm][0] += 1
            agg[nm][1] += k["dur"]
        print("  kernels by time:")
        for nm, (n, d) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
            print(f"    {d / 1e3:8.3f} ms  x{n:<4d} {nm}")
        print("  total kernel time %.3f ms" % (sum(k["dur"] for k in ks) / 1e3))
        if os.environ.get("LIST"):
            print("  every device op of the step (start us, duration us, grid, name):")
            for k in ks:
                g = k.get("args", {}).get("grid", "")
                print(f"    {k['ts'] - ks[0]['ts']:9.1f} {k['dur']:8.1f} {str(g):>14s}  {k['name'][:70]}")
