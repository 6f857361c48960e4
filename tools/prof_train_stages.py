"""Host/device time of every stage of the training step (predict_pose, render_prediction, matching_terms, backward,
apply_gradients) with a device synchronisation after each, to find where sporadic long steps come from."""
import os
import sys
import time
import types
import gc

import numpy as np
import torch

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root)
from dfnet_b200 import direct_feature_matching as dfm  # noqa: E402
from dfnet_b200 import nerfw  # noqa: E402
from dfnet_b200.dfnet import DFNet  # noqa: E402
from dfnet_b200.misc import PoseLoss  # noqa: E402

dev = torch.device("cuda:0")
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
c, f, ea, et = [m.to(dev) for m in nerfw.make_synthetic_nerf(D=8, W=256, fine=True)]
for m in (c, f, ea, et):
    for p in m.parameters():
        p.requires_grad_(False)
kw = dict(network_query_fn=None, perturb=0.0, N_importance=128, network_fine=f, N_samples=64, network_fn=c, use_viewdirs=True,
          white_bkgd=False, raw_noise_std=0.0, embedding_a=ea, embedding_t=et, test_time=True, ndc=False, lindisp=False, near=0.0, far=2.5)
args = types.SimpleNamespace(DFNet=True, preprocess_ImgNet=False, svd_reg=True, combine_loss=True, per_channel=False, chunk=32768,
                             batch_size=1, combine_loss_w=[0.0, 0.0, 1.0], feature_matching_lvl=[0])
opt = torch.optim.Adam([p for p in Fnet.parameters() if p.requires_grad], lr=1e-5)
data = torch.from_numpy(np.random.RandomState(0).rand(1, 3, 480, 640).astype(np.float32)).pin_memory()
pose = torch.tensor([[1., 0, 0, 0.1, 0, 1, 0, -0.05, 0, 0, 1, 2.0]]).to(dev)
hist = torch.tensor([[5., 10, 20, 30, 15, 10, 5, 3, 1, 1]]).to(dev)
ws = dict(pose_scale=0.5, pose_scale2=1.0, move_all_cam_vec=[0.0, 0.0, 0.05])
hwf = (480, 640, 585.0)
names = ["h2d", "predict_pose", "render", "matching", "backward", "apply", "readback"]
if os.environ.get("NOGC"):
    gc.disable()
rows = []
for it in range(40):
    t = [time.perf_counter()]

    def mark():
        torch.cuda.synchronize()
        t.append(time.perf_counter())
    d = data.to(dev)
    mark()
    pose_, pose_nerf = dfm.predict_pose(args, d, Fnet, ws, dev)
    mark()
    rgb, _ = dfm.render_prediction(args, pose_nerf[0, :3, :4], hist, hwf, True, kw)
    mark()
    pl, fl = dfm.matching_terms(args, d, rgb, Gnet, dev)
    loss = 0.0 * PoseLoss(args, pose_, pose, dev) + 0.0 * pl + 1.0 * fl
    mark()
    loss.backward()
    mark()
    dfm.apply_gradients(Fnet, opt)
    mark()
    float(loss)
    mark()
    rows.append([1e3 * (b - a) for a, b in zip(t, t[1:])])
    st = torch.cuda.memory_stats()
    rows[-1].append(st["num_alloc_retries"])
    rows[-1].append(st["num_device_alloc"])
    rows[-1].append(st["num_device_free"])
rows = np.array(rows)
print("stage ms (median over steps 5..):", dict(zip(names, np.round(np.median(rows[5:, :7], 0), 2))), "sum", np.round(np.median(rows[5:, :7].sum(1)), 2))
for i, r in enumerate(rows):
    if i < 3 or r[:7].sum() > 1.5 * np.median(rows[5:, :7].sum(1)):
        print("step", i, "total %.1f" % r[:7].sum(), dict(zip(names, np.round(r[:7], 1))), "retries/alloc/free", r[7:])
print("gc counts", gc.get_count(), "gc enabled", gc.isenabled())
