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


def test_cv_host_logic():
    """emCV / mcmcCV host side (R/cv.R): hold-out sets and the predictive-ability summary, no device involved."""
    import numpy as np
    from bwgr_b200 import api
    hs = api._cv_holdouts(196, k=5, n=3, llo=None, seed=7)
    assert len(hs) == 3 and all(len(w) == 39 and len(set(w.tolist())) == 39 and w.max() < 196 for w in hs)  # round(N/k) rows each
    assert [w.tolist() for w in hs] == [w.tolist() for w in api._cv_holdouts(196, 5, 3, None, 7)]            # seed-reproducible
    lev = np.array(list("aabbbcc"))
    assert [w.tolist() for w in api._cv_holdouts(7, 5, 5, lev, 1)] == [[0, 1], [2, 3, 4], [5, 6]]             # leave-level-out
    rng = np.random.default_rng(0)
    obs = rng.normal(size=50)
    M = np.stack([obs + rng.normal(size=50) * s for s in (2.0, 0.1, 0.7)] + [obs], axis=1)
    pooled = api._cv_summary([M[:25], M[25:]], ["m1", "m2", "m3"], avg=True)
    assert list(pooled) == ["m2", "m3", "m1"]                                                                  # sorted, best first
    assert pooled["m2"] == round(float(np.corrcoef(M[:, 1], obs)[0, 1]), 4)
    per = api._cv_summary([M[:25], M[25:]], ["m1", "m2", "m3"], avg=False)
    assert list(per) == ["CV_1", "CV_2"] and per["CV_2"]["m3"] == round(float(np.corrcoef(M[25:, 2], obs[25:])[0, 1]), 4)
