"""Golden vectors for ONE training step from the UNMODIFIED reference `train_on_batch`
(/root/reference/script/feature/direct_feature_matching.py:322-390), run on the CPU in fp32.

    python tests/golden/make_golden_train.py   ->  tests/golden/train_golden.npz

Networks are seeded synthetic ones (tests/helpers.py builds the identical state_dicts; SHA-256 recorded).  The pose
regressor's fc_pose is re-initialised so that it predicts a camera looking at the synthetic scene (helpers.pose_head_init_).
Two deviations from a verbatim call, both outside the arithmetic: torch.set_default_tensor_type('torch.cuda.FloatTensor')
is stubbed (no GPU in the build container) and the optimizer is a recorder that keeps the gradients `loss.backward()`
produced instead of applying them.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
for _m in ["imageio", "matplotlib", "matplotlib.pyplot", "pytorch3d", "pytorch3d.transforms", "efficientnet_pytorch",
           "torchsummary", "kornia", "transforms3d", "transforms3d.euler", "transforms3d.quaternions", "pykalman",
           "configargparse"]:
    sys.modules.setdefault(_m, types.ModuleType(_m))
sys.modules["efficientnet_pytorch"].EfficientNet = object
sys.modules["torchsummary"].summary = lambda *a, **k: None
sys.path[:0] = ["/root/reference/script", "/root/reference"]

import torch  # noqa: E402
import torchvision  # noqa: E402

_orig_vgg16 = torchvision.models.vgg16
torchvision.models.vgg16 = lambda pretrained=False, **kw: _orig_vgg16(weights=None)
torch.set_default_tensor_type = lambda *a, **k: None

from feature import dfnet as ref_dfnet  # noqa: E402
from feature import direct_feature_matching as ref_dfm  # noqa: E402
from utils.utils import freeze_bn_layer_train  # noqa: E402
from helpers import sd_checksum, synthetic_dfnet, pose_head_init_, train_case  # noqa: E402
import make_golden as mg  # noqa: E402  (reference NeRF-W builders)

torch.set_num_threads(8)
G = {}


def put(k, v):
    G[k] = np.ascontiguousarray(v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else v)


class Recorder:
    """Stands in for the optimizer: keeps the gradients of the step."""

    def __init__(self, model):
        self.model, self.grads = model, None

    def step(self):
        self.grads = {n: p.grad.detach().clone() for n, p in self.model.named_parameters() if p.grad is not None}

    def zero_grad(self):
        self.model.zero_grad()


def main():
    for case in ("lvl0", "lvl012"):
        cfg = train_case(case)
        torch.manual_seed(0)
        F_ref = pose_head_init_(ref_dfnet.DFNet())
        F_mine = pose_head_init_(synthetic_dfnet("DFNet", seed=0))
        torch.manual_seed(1)
        G_ref = ref_dfnet.DFNet().eval()
        G_mine = synthetic_dfnet("DFNet", seed=1)
        for a, b in ((F_ref, F_mine), (G_ref, G_mine)):
            sa, sb = a.state_dict(), b.state_dict()
            assert list(sa) == list(sb)
            for k in sa:
                assert torch.equal(sa[k], sb[k]), k
        put(f"{case}_F_sha", np.frombuffer(sd_checksum(F_ref.state_dict()).encode(), np.uint8))
        put(f"{case}_G_sha", np.frombuffer(sd_checksum(G_ref.state_dict()).encode(), np.uint8))
        coarse, fine, emb_a, emb_t = mg.build_nets(cfg["D"], cfg["W"])
        for m in (coarse, fine, emb_a, emb_t):
            ref_dfm.disable_model_grad(m)
        kw = mg.render_kwargs(coarse, fine, emb_a, emb_t, cfg["Nc"], cfg["Nf"], True)
        kw.update(near=cfg["near"], far=cfg["far"])
        F_ref.train()
        F_ref = freeze_bn_layer_train(F_ref)
        rec = Recorder(F_ref)
        loss, psnr = ref_dfm.train_on_batch(cfg["args"], cfg["data"], F_ref, G_ref, cfg["pose"], cfg["hist"], cfg["hwf"], rec, True,
                                            torch.device("cpu"), cfg["world"], **kw)
        put(f"{case}_loss", loss)
        put(f"{case}_psnr", psnr)
        names = sorted(rec.grads)
        put(f"{case}_grad_names", np.frombuffer("\n".join(names).encode(), np.uint8))
        for n in names:
            g = rec.grads[n].flatten()
            put(f"{case}_g_{n}_stats", torch.stack([g.norm(), g.sum(), g.abs().max()]).double())
            put(f"{case}_g_{n}_sub", g[:: max(1, g.numel() // 4096)][:4096])
        print(case, "loss", loss, "psnr", psnr, "n grads", len(names))
    np.savez_compressed(os.path.join(HERE, "train_golden.npz"), **G)
    print("wrote", len(G), "arrays")


if __name__ == "__main__":
    main()
