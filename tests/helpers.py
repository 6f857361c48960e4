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
