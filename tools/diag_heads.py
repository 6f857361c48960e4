import sys, torch, torch.nn.functional as F
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from helpers import synthetic_dfnet
from test_train_gpu import torch_dfnet_forward, _randomise_bn
from dfnet_b200 import misc
dev = torch.device("cuda:0")
for bn_mode in ("train", "eval"):
    net = synthetic_dfnet("DFNet", seed=5).to(dev)
    with torch.no_grad():
        _randomise_bn(net, 9)
    net.train()
    if bn_mode == "eval":
        for m in net.modules():
            if isinstance(m, torch.nn.BatchNorm2d): m.eval()
    B, H, W = 2, 48, 64
    torch.manual_seed(12)
    x = torch.rand(2 * B, 3, H, W, device=dev)
    wt = torch.randn(3, B, 128, H, W, device=dev) / (3 * B * 128 * H * W) ** 0.5
    wr = torch.randn(3, B, 128, H, W, device=dev) / (3 * B * 128 * H * W) ** 0.5
    target = torch.randn(2 * B, 12, device=dev)
    for mode in ("feat", "pose", "both"):
        def loss_of(feats, pose, mse):
            l = 0
            if mode in ("feat", "both"): l = l + (feats[0] * wt).sum() + (feats[1] * wr).sum()
            if mode in ("pose", "both"): l = l + 0.5 * mse(pose, target)
            return l
        net.zero_grad()
        feats, pose = net(x, return_feature=True, isSingleStream=False, return_pose=True, upsampleH=H, upsampleW=W)
        loss_of(feats, pose, misc.mse).backward()
        got = {n: p.grad.clone() for n, p in net.named_parameters() if p.grad is not None}
        acts = net._handle.tape_activations()
        net.zero_grad()
        f_t, pose_t = torch_dfnet_forward(net, x, True, False, True, H, W, torch.float16, acts)
        print(bn_mode, mode, "fwd feat err", [float((a - b).abs().max() / b.abs().max()) for a, b in zip(feats, f_t)])
        loss_of(f_t, pose_t, F.mse_loss).backward()
        for n, p in net.named_parameters():
            if p.grad is None or n not in got:
                print("   ", n, "missing", p.grad is None, n in got); continue
            g, w = got[n].double().flatten(), p.grad.double().flatten()
            cos = float(F.cosine_similarity(g, w, dim=0)); ratio = float(g.norm() / (w.norm() + 1e-30))
            if cos < 0.999 or abs(ratio - 1) > 0.02:
                print("   ", n, "cos %.4f norm ratio %.4f" % (cos, ratio))
