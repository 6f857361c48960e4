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
