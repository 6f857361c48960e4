"""CPU-side checks of the C-ABI library: it loads, exports every declared symbol, the host
helper matches ATen, and compute entry points refuse to run without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from dfnet_b200 import _lib, ops

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported():
    hdr = open(os.path.join(ROOT, "include", "dfnet_b200.h")).read()
    declared = set(re.findall(r"\b(dfb_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SYMBOLS)
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for s in declared:
        assert hasattr(raw, s), s


def test_linspace_host_helper_matches_aten(golden):
    i = 0
    while f"linspace_{i}" in golden:
        a, b, n = golden[f"linspace_{i}_args"]
        got = ops.linspace(float(a), float(b), int(n)).numpy()
        assert np.array_equal(got, golden[f"linspace_{i}"])
        assert np.array_equal(got, torch.linspace(float(a), float(b), int(n)).numpy())
        i += 1


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_no_cpu_fallback():
    from dfnet_b200 import nerfw
    assert _lib.lib.dfb_device_ok() == 0
    c, f, ea, et = nerfw.make_synthetic_nerf(D=4, W=64)
    with pytest.raises(_lib.DfbError):
        ops.NerfHandle(c, f, ea, et)
    with pytest.raises(_lib.DfbError):
        ops.sample_pdf(torch.zeros(2, 5), torch.zeros(2, 4), 8, det=True)
    with pytest.raises(_lib.DfbError):
        c(torch.zeros(4, 63), sigma_only=True)


def test_copy2d_struct_matches_the_header():
    """ctypes mirror of DfbCopy2d: two pointers and four ints, 32 bytes, fields in header order."""
    hdr = open(os.path.join(ROOT, "include", "dfnet_b200.h")).read()
    body = re.search(r"typedef struct DfbCopy2d \{(.*?)\} DfbCopy2d;", hdr, re.S).group(1)
    names = re.findall(r"\*?\s*(\w+)\s*[,;]", body.replace("const float", "").replace("float", "").replace("int", ""))
    assert names == [f[0] for f in _lib.Copy2d._fields_]
    assert ctypes.sizeof(_lib.Copy2d) == 32


def test_direction_embedding_matches_the_band_loop():
    """nerf_train._embed (all bands in one pass) == the per-band concatenation of models/nerfw.py Embedding, element for element."""
    from dfnet_b200 import nerf_train
    x = torch.randn(257, 3) * 3
    for L in (1, 4, 10):
        want = [x]
        for l in range(L):
            want += [torch.sin(x * 2.0 ** l), torch.cos(x * 2.0 ** l)]
        assert torch.equal(nerf_train._embed(x, L), torch.cat(want, -1))


def test_nerfw_loss_host_path_and_fused_gate():
    """NerfWLoss on host tensors evaluates the tensor expressions of models/losses.py:42-57 (the fused kernels take CUDA
    fp32 tensors of the per-ray shapes only); values against the formulas written out."""
    from dfnet_b200 import losses
    torch.manual_seed(0)
    N, S = 9, 5
    inp = dict(rgb_coarse=torch.rand(N, 3), rgb_fine=torch.rand(N, 3), beta=torch.rand(N) + 0.1, transient_sigmas=torch.rand(N, S))
    tg = torch.rand(N, 3)
    assert not losses._fused_ok(inp, tg)                       # host tensors
    lf = losses.NerfWLoss(coef=2.0, lambda_u=0.03)
    out = lf(inp, tg)
    assert lf.last_mse_fine is None
    assert torch.allclose(out["c_l"], 2.0 * 0.5 * ((inp["rgb_coarse"] - tg) ** 2).mean())
    assert torch.allclose(out["f_l"], 2.0 * (((inp["rgb_fine"] - tg) ** 2) / (2 * inp["beta"][:, None] ** 2)).mean())
    assert torch.allclose(out["b_l"], 2.0 * (3 + torch.log(inp["beta"]).mean()))
    assert torch.allclose(out["s_l"], 2.0 * 0.03 * inp["transient_sigmas"].mean())
    # coarse only / no beta: the reference's other branches
    assert set(lf({"rgb_coarse": inp["rgb_coarse"]}, tg)) == {"c_l"}
    assert set(lf({"rgb_coarse": inp["rgb_coarse"], "rgb_fine": inp["rgb_fine"]}, tg)) == {"c_l", "f_l"}
