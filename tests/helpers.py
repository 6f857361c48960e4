"""Shared test helpers: synthetic networks as numpy parameter dicts for the oracle."""
import hashlib

import numpy as np
import torch

from dfnet_b200 import nerfw as my_nerfw


def sd_checksum(sd):
    h = hashlib.sha256()
    for k in sd:
        h.update(k.encode())
        h.update(sd[k].detach().cpu().numpy().tobytes())
    return h.hexdigest()


def np_params(mod):
    return {k: v.detach().cpu().numpy() for k, v in mod.state_dict().items()}


_CACHE = {}


def synthetic_nets(D, W, fine=True):
    """(torch modules, oracle dict) for the seeded synthetic networks of SURVEY §8(d)."""
    key = (D, W, fine)
    if key not in _CACHE:
        state = torch.get_rng_state()
        c, f, ea, et = my_nerfw.make_synthetic_nerf(D=D, W=W, fine=fine)
        torch.set_rng_state(state)
        nets = dict(coarse=np_params(c), fine=np_params(f) if f is not None else None,
                    emb_a=ea.weight.detach().numpy(), emb_t=et.weight.detach().numpy(),
                    D=D, skips=(4,), beta_min=0.1)
        _CACHE[key] = ((c, f, ea, et), nets)
    return _CACHE[key]


def rel_err(a, b, floor=1e-3):
    """max |a-b| / max(|b|, floor): the relative-error measure used for the 1e-3 bar."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor))) if a.size else 0.0


def synthetic_dfnet(cls_name="DFNet", seed=0):
    """dfnet_b200 DFNet/DFNet_s with the weights the reference constructor produces under `seed`
    when torchvision's vgg16 is built without pretrained weights (test infrastructure: pretrained
    ImageNet weights are not available offline)."""
    import torchvision
    from dfnet_b200 import dfnet as my_dfnet
    state = torch.get_rng_state()
    torch.manual_seed(seed)
    vgg = torchvision.models.vgg16(weights=None)
    net = getattr(my_dfnet, cls_name)()          # builds encoder / heads / fc in the reference's order
    # the reference takes the encoder from torchvision and only then creates heads and fc_pose:
    # rebuild those so that they consume the RNG stream exactly like the reference constructor
    torch.manual_seed(seed)
    vgg = torchvision.models.vgg16(weights=None)
    net.encoder.load_state_dict(vgg.features.state_dict())
    net.adaptation_layers = my_dfnet.AdaptLayers(net.hypercolumn_layers, 128)
    net.fc_pose = torch.nn.Linear(512, 12)
    torch.set_rng_state(state)
    return net.eval()


def pose_head_init_(net):
    """Deterministic fc_pose so that a randomly initialised pose regressor predicts a camera that looks at the
    synthetic scene: small weights, bias = a slightly non-orthogonal [R | t] (svd_reg has something to fix)."""
    g = torch.Generator().manual_seed(123)
    with torch.no_grad():
        net.fc_pose.weight.copy_(torch.randn(12, 512, generator=g) * 2e-3)
        net.fc_pose.bias.copy_(torch.tensor([0.98, 0.02, 0.15, 0.2, -0.03, 1.01, 0.04, -0.1, -0.16, -0.02, 0.97, 1.9]))
    return net


def train_case(case):
    """Inputs of the train_on_batch golden cases (shared by tests/golden/make_golden_train.py and the GPU test)."""
    import types
    rng = np.random.RandomState(21)
    H, W = 64, 96
    args = types.SimpleNamespace(
        DFNet=True, preprocess_ImgNet=False, svd_reg=True, combine_loss=True, per_channel=False, chunk=32768, batch_size=1,
        combine_loss_w=[0.3, 0.2, 1.0] if case == "lvl012" else [0.0, 0.0, 1.0],
        feature_matching_lvl=[0, 1, 2] if case == "lvl012" else [0])
    data = torch.from_numpy(rng.rand(1, 3, H, W).astype(np.float32))
    pose = torch.from_numpy(np.array([[1, 0, 0, 0.1, 0, 1, 0, -0.05, 0, 0, 1, 2.0]], np.float32))
    hist = torch.from_numpy(np.array([[5, 10, 20, 30, 15, 10, 5, 3, 1, 1]], np.float32))
    world = dict(pose_scale=0.5, pose_scale2=1.0, move_all_cam_vec=[0.0, 0.0, 0.05])
    return dict(args=args, data=data, pose=pose, hist=hist, hwf=(H, W, 80.0), world=world, D=8, W=64, Nc=16, Nf=24,
                near=0.0, far=2.5)
