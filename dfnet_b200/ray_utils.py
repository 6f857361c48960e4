"""Host mirror of the reference's models/ray_utils.py (get_rays only; NDC is forced off by
create_nerf, reference nerfw.py:492-495)."""
from . import ops


def get_rays(H, W, focal, c2w):
    """Pinhole rays for an H x W image (reference ray_utils.py:5-15) -> rays_o, rays_d [H,W,3]."""
    return ops.get_rays(int(H), int(W), float(focal), c2w)
