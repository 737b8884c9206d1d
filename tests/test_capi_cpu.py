"""CPU-side checks of the boundary: the shared object loads, exports every symbol include/bwgr_b200.h
declares, and fails loudly (no CPU fallback) when there is no device."""
import os
import re

import pytest

from conftest import ROOT


def _lib():
    from bwgr_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        from bwgr_b200 import build
        build.build()
    return _lib


def test_header_symbols_exported():
    L = _lib()
    hdr = open(os.path.join(ROOT, "include", "bwgr_b200.h")).read()
    declared = sorted(set(re.findall(r"BWGR_API [\w\s\*]*?\b(bwgr_\w+)\(", hdr)))
    assert declared == sorted(L.SYMBOLS)
    cdll = L.load()
    for s in declared:
        assert getattr(cdll, s) is not None
    assert cdll.bwgr_version() >= 100


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import bwgr_b200 as bw
    with pytest.raises(bw.BwgrError) as ei:
        bw.Genotypes()
    assert ei.value.code == -2 and "no CPU path" in str(ei.value)


def test_product_does_not_touch_oracle():
    """The oracle is test infrastructure: nothing under bwgr_b200/ or include/ may reference it."""
    for base in ("bwgr_b200", "include"):
        for dp, _, fs in os.walk(os.path.join(ROOT, base)):
            for f in fs:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                    txt = open(os.path.join(dp, f), errors="ignore").read()
                    assert "oracle" not in txt.lower(), os.path.join(dp, f)
