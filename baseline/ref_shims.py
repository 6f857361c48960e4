"""Import the UNMODIFIED reference from baseline/_ref/ (or /root/reference when it exists).

Non-arithmetic shims only: stub modules for packages the image lacks and the reference imports at
module scope without using them on this path (imageio, matplotlib, pytorch3d, efficientnet_pytorch, ...), torchvision's
vgg16(pretrained=True) -> vgg16(weights=None) (no network), and on a CPU-only process
torch.set_default_tensor_type('torch.cuda.FloatTensor') -> 'torch.FloatTensor'.  Nothing under dfnet_b200/ imports
this file: it serves bench.py --impl reference and the drop-in tests."""
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
_STUBS = ["imageio", "matplotlib", "matplotlib.pyplot", "pytorch3d", "pytorch3d.transforms", "efficientnet_pytorch",
          "torchsummary", "kornia", "transforms3d", "transforms3d.euler", "transforms3d.quaternions", "pykalman",
          "configargparse"]


def reference_root():
    for r in (os.path.join(HERE, "_ref"), os.environ.get("DFB_REFERENCE", "/root/reference")):
        if os.path.isdir(os.path.join(r, "script", "models")):
            return r
    return None


def available():
    return reference_root() is not None


def activate(cpu_default_tensor_type=None):
    """Put the reference on sys.path (idempotent) and install the stubs.  Returns the root used."""
    root = reference_root()
    if root is None:
        raise ImportError("the reference is not available: run `python baseline/make_ref.py` where /root/reference exists")
    for m in _STUBS:
        if m not in sys.modules:
            try:
                __import__(m)
            except Exception:  # noqa: BLE001
                sys.modules[m] = types.ModuleType(m)
    eff = sys.modules["efficientnet_pytorch"]
    if not hasattr(eff, "EfficientNet"):
        eff.EfficientNet = object
    ts = sys.modules["torchsummary"]
    if not hasattr(ts, "summary"):
        ts.summary = lambda *a, **k: None
    for p in (os.path.join(root, "script"), root):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torchvision
    if not getattr(torchvision.models.vgg16, "_dfb_offline", False):
        orig = torchvision.models.vgg16

        def vgg16(pretrained=False, **kw):  # the reference asks for ImageNet weights (feature/dfnet.py:90): no network here
            kw.pop("weights", None)
            return orig(weights=None, **kw)
        vgg16._dfb_offline = True
        torchvision.models.vgg16 = vgg16
    if cpu_default_tensor_type is None:
        cpu_default_tensor_type = not torch.cuda.is_available()
    if cpu_default_tensor_type and not getattr(torch.set_default_tensor_type, "_dfb_cpu", False):
        orig_set = torch.set_default_tensor_type

        def set_default_tensor_type(t):
            return orig_set("torch.FloatTensor" if "cuda" in str(t) else t)
        set_default_tensor_type._dfb_cpu = True
        torch.set_default_tensor_type = set_default_tensor_type
    return root
