"""The oracle pinned against the reference's own text.

oracle/_ref/libbwgr_ref.so = /root/reference/src/Rcpp20260726ai.cpp (whole file) and MRR3 / MRR3F of RcppEigen20230423.cpp
(:317-1079), compiled UNMODIFIED against the stand-in RcppEigen / Rcpp headers of oracle/shim/ (oracle/Makefile, target `ref`).
These tests run the reference's functions and the hand-written oracle on the same inputs: agreement is to float reassociation
(the shim's reductions are not Eigen's packets), and exact to 1e-13 for the float64 MRR3.  The Gibbs samplers consume the same
std::mt19937_64 stream draw for draw (the shim's R::rnorm / rchisq / rbinom and the oracle's Rng are the same distributions), so
even the chains agree.  CPU only; skipped when the library is absent and cannot be built (no /root/reference).
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle as O  # noqa: E402
import ref as R  # noqa: E402

pytestmark = pytest.mark.skipif(not R.available(), reason="oracle/_ref not built and /root/reference absent")


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


@pytest.fixture(scope="module")
def tpod():
    d = np.load(os.path.join(ROOT, "tests", "golden", "tpod.npz"))
    return d["y"].astype(np.float64), d["gen"].astype(np.float64)


@pytest.mark.parametrize("model", list(O.EM_MODELS))
def test_em_solvers_match_reference_text(tpod, model):
    y, X = tpod
    a, r = O.em(model, y, X), R.em(model, y, X)
    tol = 2e-3 if model == "emBCpi" else 5e-5   # emBCpi: the inclusion feedback (Pi <- mean d) amplifies float reassociation
    for key in r:
        assert rel(a[key], r[key]) < tol, (model, key, rel(a[key], r[key]))


def test_em_other_hyperparameters(tpod):
    y, X = tpod
    for model, kw in (("emRR", dict(df=4.0, R2=0.3)), ("emBB", dict(df=6.0, R2=0.4, Pi=0.9)), ("emBC", dict(Pi=0.6)),
                      ("emBL", dict(R2=0.3, alpha=0.1)), ("emEN", dict(R2=0.4, alpha=0.5))):
        a, r = O.em(model, y, X, **kw), R.em(model, y, X, **kw)
        for key in r:
            assert rel(a[key], r[key]) < 1e-4, (model, key)


def test_em_synthetic_tall(tpod):
    # n > p, general small-integer genotypes (not only {0,1,2})
    rng = np.random.default_rng(5)
    X = rng.integers(-2, 4, size=(300, 120)).astype(np.float64)
    y = X[:, :10] @ rng.normal(size=10) + rng.normal(size=300)
    for model in ("emRR", "emBA", "emBC", "emDE", "emML"):
        a, r = O.em(model, y, X), R.em(model, y, X)
        assert rel(a["b"], r["b"]) < 1e-4 and abs(a["h2"] - r["h2"]) < 1e-4, model


@pytest.mark.parametrize("model", list(O.GIBBS_MODELS))
def test_gibbs_chains_match_reference_text(tpod, model):
    y, X = tpod
    a, r = O.gibbs(model, y, X, it=80, bi=30, seed=11), R.gibbs(model, y, X, it=80, bi=30, seed=11)
    for key in r:
        assert rel(a[key], r[key]) < 2e-4, (model, key, rel(a[key], r[key]))


@pytest.mark.parametrize("model", list(O.TWO_DESIGN))
def test_two_design_solvers_match_reference_text(tpod, model):
    """BayesA2 / BayesB2 / BayesRR2 / emML2 (Rcpp20260726ai.cpp:990-1305) on tpod split into two designs: the chains draw for draw,
    emML2 to float reassociation over its 350 sweeps (with and without marker weights)."""
    y, X = tpod
    X1, X2 = X[:, :200], X[:, 200:]
    cases = [dict(it=80, bi=30, seed=11)]
    if model == "emML2":
        rng = np.random.default_rng(1)
        cases = [dict(), dict(D1=rng.uniform(0.5, 2, 200), D2=rng.uniform(0.5, 2, 176))]
    for kw in cases:
        a, r = O.two_design(model, y, X1, X2, **kw), R.two_design(model, y, X1, X2, **kw)
        for key in r:
            assert rel(a[key], r[key]) < (1e-3 if model == "emML2" else 2e-4), (model, key, rel(a[key], r[key]))


def test_kmup_sweep_matches_reference_text(tpod):
    y, X = tpod
    n, p = X.shape
    rng = np.random.default_rng(2)
    xx = (X * X).sum(0)
    b0 = rng.normal(size=p) * 0.01
    e0 = y - y.mean() - X @ b0
    L = np.full(p, 50.0)
    for pi in (0.0, 0.3):
        a = O.kmup(X, b0, np.ones(p), xx, e0, L, 0.03, pi, seed=4, ratio_form=False)
        r = R.kmup(X, b0, np.ones(p), xx, e0, L, 0.03, pi, seed=4)
        assert np.array_equal(a["d"], r["d"])
        assert rel(a["b"], r["b"]) < 1e-4 and rel(a["e"], r["e"]) < 1e-4


MRR_CASES = [dict(), dict(maxit=3), dict(HCS=True, maxit=30), dict(XFA=True, NumXFA=2, maxit=30), dict(ACS=True, maxit=30),
             dict(updateMu=True, maxit=30), dict(OneVarB=True, OneVarE=True, maxit=30), dict(InnerGS=True, maxit=30),
             dict(TH=True, maxit=30), dict(NLfactor=0.5, maxit=30), dict(NoInv=True, maxit=30),
             dict(DeflateBy=0.05, PenCor=0.5, MinCor=0.1, maxit=20), dict(weight_prior_h2=0.0, weight_prior_gc=0.0, maxit=20)]


@pytest.mark.parametrize("case", range(len(MRR_CASES)))
def test_mrr3_matches_reference_text(tpod, case):
    _, X = tpod
    Y = np.load(os.path.join(ROOT, "tests", "golden", "tpod_mrr3.npz"))["Y"]
    kw = MRR_CASES[case]
    for f32, tol in ((False, 1e-11), (True, 2e-5)):
        a, r = O.mrr3(Y, X, f32_variant=f32, **kw), R.mrr3(Y, X, f32_variant=f32, **kw)
        assert a["Its"] == r["Its"]
        for key in ("mu", "b", "hat", "h2", "GC", "vb", "ve", "MSx", "b_Weights") + (() if f32 else ("cnvB",)):  # float cnvB = log10 of round-off
            assert rel(a[key], r[key]) < tol, (kw, f32, key, rel(a[key], r[key]))


def test_mrr3_missing_phenotypes_match_reference_text(tpod):
    _, X = tpod
    Y = np.load(os.path.join(ROOT, "tests", "golden", "tpod_mrr3.npz"))["Y"].copy()
    Y[3, 1] = np.nan
    Y[10:40, 2] = np.nan
    Y[100:, 0] = np.nan
    for kw in (dict(maxit=25), dict(maxit=25, InnerGS=True), dict(maxit=25, TH=True)):
        a, r = O.mrr3(Y, X, **kw), R.mrr3(Y, X, **kw)
        for key in ("mu", "b", "hat", "h2", "GC", "vb", "ve"):
            assert rel(a[key], r[key]) < 1e-11, (kw, key)


def test_goldens_are_reference_output(tpod):
    """tests/golden/tpod_em.npz and tpod_mrr3.npz are written by oracle/make_golden.py from the _ref library."""
    y, X = tpod
    g = np.load(os.path.join(ROOT, "tests", "golden", "tpod_em.npz"))
    assert str(g["provenance"]) == "reference-executed"
    for model in O.EM_MODELS:
        r = R.em(model, y, X)
        for key, v in r.items():
            assert np.allclose(g[model + "_ref__" + key], v, rtol=0, atol=0), (model, key)


def _gs_inputs(tpod):
    y, X = tpod
    n, p = X.shape
    xx = (X * X).sum(0)
    cxx = float(X.var(0, ddof=1).sum())
    return y, X, n, p, xx, cxx


@pytest.mark.parametrize("which", ["GSRR", "GSFLM"])
def test_gs_warm_start_solvers_match_reference_text(tpod, which):
    """GSRR / GSFLM (Rcpp20260726ai.cpp:1564-1628), including a second call warm-started from the first one's state."""
    y, X, n, p, xx, cxx = _gs_inputs(tpod)
    e = y - 0.05
    b = np.zeros(p)
    L = np.full(p, cxx)
    a = O.gs(which, y, e, X, b, L, xx, cxx, maxit=7)
    r = R.gs(which, y, e, X, b, L, xx, cxx, maxit=7)
    for key in ("b", "e", "Lmb"):
        assert rel(a[key], r[key]) < 5e-5, (which, key, rel(a[key], r[key]))
    assert abs(a["mu"] - r["mu"]) < 1e-6 and abs(a["h2"] - r["h2"]) < 1e-5
    a2 = O.gs(which, y, a["e"] + a["mu"], X, a["b"], a["Lmb"], xx, cxx, maxit=50)
    r2 = R.gs(which, y, r["e"] + r["mu"], X, r["b"], r["Lmb"], xx, cxx, maxit=50)
    assert rel(a2["b"], r2["b"]) < 2e-4 and abs(a2["h2"] - r2["h2"]) < 1e-4


def test_kmup2_and_preprocessing_match_reference_text(tpod):
    y, X = tpod
    n, p = X.shape
    rng = np.random.default_rng(4)
    use = np.sort(rng.choice(n, size=n // 2, replace=False)).astype(np.float64)
    xx = (X * X).sum(0)
    b0 = rng.normal(size=p) * 0.01
    E = y - y.mean() - X @ b0
    L = np.full(p, 40.0)
    for pi in (0.0, 0.3):
        a = O.kmup2(X, use, b0, np.ones(p), xx, E, L, 0.03, pi, seed=8)
        r = R.kmup2(X, use, b0, np.ones(p), xx, E, L, 0.03, pi, seed=8)
        assert np.array_equal(a["d"], r["d"]) and rel(a["b"], r["b"]) < 1e-4 and rel(a["e"], r["e"]) < 1e-4
    # Use with repeats: what wgr(bag, rp = TRUE) passes (R/wgr.R:68, sort(sample(n, n*bag, TRUE)) - 1)
    use_r = np.sort(rng.integers(0, n, size=n // 2)).astype(np.float64)
    assert np.unique(use_r).size < use_r.size
    for pi in (0.0, 0.3):
        a = O.kmup2(X, use_r, b0, np.ones(p), xx, E, L, 0.03, pi, seed=8)
        r = R.kmup2(X, use_r, b0, np.ones(p), xx, E, L, 0.03, pi, seed=8)
        assert np.array_equal(a["d"], r["d"]) and rel(a["b"], r["b"]) < 1e-4 and rel(a["e"], r["e"]) < 1e-4
    Xn = X.copy()
    Xn[rng.random(X.shape) < 0.03] = np.nan
    assert np.allclose(O.imp(Xn), R.imp(Xn), rtol=0, atol=1e-6) and not np.isnan(O.imp(Xn)).any()
    assert np.allclose(O.cnt(X), R.cnt(X), rtol=0, atol=1e-6)


def test_emml_marker_weights_match_reference_text(tpod):
    y, X = tpod
    D = np.random.default_rng(9).uniform(0.5, 2.0, size=X.shape[1])
    a, r = O.emML_weighted(y, X, D), R.emML_weighted(y, X, D)
    for key in r:
        assert rel(a[key], r[key]) < 5e-5, (key, rel(a[key], r[key]))
    assert rel(a["b"], O.em("emML", y, X)["b"]) > 1e-2  # the weights matter
