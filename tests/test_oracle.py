"""CPU tests of the oracle itself: known answers, golden fixtures, and an independent numpy
restatement (float64) of emRR and default-flag MRR3 so that a typo in the C++ oracle is caught.
Parity is "unpinned" (the reference ships no tests for this path): what can be
pinned is pinned here -- the libstdc++ marker order, the tpod data, oracle self-consistency."""
import json
import os

import numpy as np
import pytest

import oracle as O
from conftest import GOLDEN


def test_perm_known_answers():
    kat = json.load(open(os.path.join(GOLDEN, "perm_kat.json")))
    # SURVEY.md 8a / BASELINE.md 5 (libstdc++ 13, std::shuffle + std::mt19937(iter), cumulative)
    assert kat["p10_iters0_2"] == [[0, 2, 1, 5, 9, 8, 4, 7, 6, 3], [3, 0, 1, 8, 7, 9, 4, 5, 2, 6],
                                   [9, 3, 8, 0, 2, 7, 4, 1, 5, 6]]
    assert kat["p376_iter0_head8"] == [229, 361, 136, 362, 239, 298, 89, 202]
    assert O.perm(10, 3).tolist() == kat["p10_iters0_2"]
    assert O.perm(376, 1)[0, :8].tolist() == kat["p376_iter0_head8"]
    assert O.perm(376, 200)[199, :8].tolist() == kat["p376_iter199_head8"]
    for row in O.perm(50, 5):
        assert sorted(row.tolist()) == list(range(50))


def test_tpod_fixture(tpod):
    y, gen = tpod
    assert y.shape == (196,) and gen.shape == (196, 376)
    assert abs(y.mean() - 0.16011) < 1e-5 and abs(y.var(ddof=1) - 0.037430) < 1e-6
    assert set(np.unique(gen)) == {0, 1, 2}
    cm = gen.mean(0)
    assert 0.836 < cm.min() and cm.max() < 1.154


@pytest.mark.parametrize("model", list(O.EM_MODELS))
def test_em_matches_golden(tpod, model):
    y, gen = tpod
    g = np.load(os.path.join(GOLDEN, "tpod_em.npz"))
    r = O.em(model, y, gen.astype(np.float64))
    for key, v in r.items():
        np.testing.assert_allclose(np.asarray(v), g[f"{model}_f32__{key}"], rtol=1e-6, atol=1e-9)
    # float recipe vs double recipe: the noise floor that the 1e-4 GPU tolerance must sit above
    # (emML stops on sum|db| < 1e-7: the float recipe stops at sweep 200, the double one at 239, 2.3e-4 apart on b)
    b64 = g[f"{model}_f64__b"]
    tol = 5e-4 if model == "emML" else 1e-4
    assert np.abs(r["b"] - b64).max() <= tol * np.abs(b64).max()
    assert abs(r["h2"] - float(g[f"{model}_f64__h2"])) <= tol


def _np_emRR(y, X, df=10.0, R2=0.5, it=200):
    n, p = X.shape
    xx = (X * X).sum(0)
    vx = X.var(0, ddof=1)
    MSx = vx.sum(); Lmb = MSx; Rho = MSx * (1 - R2) / R2
    vy = y.var(ddof=1); Se = (1 - R2) * (df + 2) * vy; Sb = R2 * (df + 2) * vy / MSx
    mu = y.mean(); b = np.zeros(p); e = y - mu
    perms = O.perm(p, it)
    for i in range(it):
        for j in perms[i]:
            b0 = b[j]
            b[j] = (X[:, j] @ e + xx[j] * b0) / (xx[j] + Lmb)
            e -= X[:, j] * (b[j] - b0)
        vb = (b @ b + Sb) / (p + df); ve = (e @ e + Se) / (n + df)
        Lmb = np.sqrt(Rho * ve / vb)
        mu += e.mean(); e -= e.mean()
    return dict(mu=mu, b=b, hat=X @ b + mu, Va=vb, Ve=ve, h2=1 - ve / vy)


def test_emRR_vs_numpy(tpod):
    y, gen = tpod
    X = gen.astype(np.float64)
    ref = _np_emRR(y, X, it=30)
    r = O.em("emRR", y, X, it=30, use_double=True)
    for key in ("mu", "b", "hat", "Va", "Ve", "h2"):
        np.testing.assert_allclose(r[key], ref[key], rtol=1e-5, atol=1e-9)  # y/X enter the oracle as float32


def _np_mrr3(Y, X, maxit=500, tol=10e-9, R2=0.5, gc0=0.5, df0=1.0, wh=0.01, wg=0.01):
    n0, k = Y.shape; p = X.shape[1]
    n = np.full(k, float(n0)); mu = Y.mean(0); y = Y - mu
    X = X - X.mean(0)
    XX = np.outer((X * X).sum(0), np.ones(k))
    MSx = (XX / n - ((X.sum(0)[:, None] / n) ** 2)).sum(0); Tr = n * MSx
    vy = (y * y).sum(0) / (n - 1); ve = vy * (1 - R2); iVe = 1 / ve
    vbInit = vy * R2 / MSx; veInit = ve.copy(); vb = np.diag(vbInit); iG = np.linalg.inv(vb)
    for i in range(k):
        for j in range(i):
            vb[i, j] = vb[j, i] = gc0 * np.sqrt(vb[i, i] * vb[j, j])
    tilde = X.T @ y; Sb = vb * df0; Se = ve * df0; iNp = 1 / (n + df0 - 1)
    b = np.zeros((p, k)); e = y.copy(); h2 = 1 - ve / vy
    perms = O.perm(p, maxit); logtol = np.log10(tol); its = 0
    for numit in range(maxit):
        beta0 = b.copy()
        for J in perms[numit]:
            b0 = b[J].copy()
            LHS = iG + np.diag(XX[J] * iVe)
            RHS = (X[:, J] @ e + XX[J] * b0) * iVe
            b1 = np.linalg.solve(LHS, RHS)
            b[J] = b1
            e -= np.outer(X[:, J], b1 - b0)
        ve = ((e * y).sum(0) + Se) * iNp; h2 = 1 - ve / vy
        ve = ve * (1 - wh) + wh * veInit; iVe = 1 / ve
        TH = b.T @ tilde
        for i in range(k):
            for j in range(k):
                vb[i, j] = (TH[i, i] + Sb[i, i]) / (Tr[i] + df0) if i == j else (TH[i, j] + TH[j, i] + Sb[i, j]) / (Tr[i] + Tr[j] + df0)
        for i in range(k):
            vb[i, i] = vb[i, i] * (1 - wh) + wh * vbInit[i]
        sd = np.sqrt(np.diag(vb)); GC = (1 - wg) * vb / np.outer(sd, sd) + gc0 * wg; np.fill_diagonal(GC, 1.0)
        w = np.linalg.eigvalsh(GC)
        if w.min() < 0:
            infl = abs(w.min() * 1.1); GC = (GC + np.eye(k) * infl) / (1 + infl)
        vb = GC * np.outer(sd, sd); iG = np.linalg.pinv(vb)
        cnv = np.log10(((beta0 - b) ** 2).sum(0).max()); its += 1
        if cnv < logtol:
            break
    return dict(mu=mu, b=b, hat=X @ b + mu, h2=h2, GC=GC, vb=vb, ve=ve, Its=its)


def test_mrr3_vs_numpy_and_golden(tpod):
    _, gen = tpod
    g = np.load(os.path.join(GOLDEN, "tpod_mrr3.npz"))
    Y = g["Y"]
    X = gen.astype(np.float64)
    r = O.mrr3(Y, X)
    for key in ("mu", "b", "hat", "h2", "GC", "vb", "ve", "cnvB"):
        np.testing.assert_allclose(r[key], g[key], rtol=1e-9, atol=1e-12)
    ref = _np_mrr3(Y, X)
    assert r["Its"] == ref["Its"]
    for key in ("mu", "b", "hat", "h2", "GC", "vb", "ve"):
        np.testing.assert_allclose(r[key], ref[key], rtol=1e-6, atol=1e-9)
    rf = O.mrr3(Y, X, f32_variant=True)
    assert np.abs(rf["b"] - r["b"]).max() <= 1e-4 * np.abs(r["b"]).max()


def test_mrr3_missing_and_options(tpod):
    _, gen = tpod
    g = np.load(os.path.join(GOLDEN, "tpod_mrr3.npz"))
    Y = g["Y"].copy()
    rng = np.random.default_rng(3)
    Y[rng.random(Y.shape) < 0.2] = np.nan
    X = gen.astype(np.float64)
    r = O.mrr3(Y, X)
    assert np.isfinite(r["b"]).all() and 0 < r["Its"] <= 500
    assert ((r["h2"] > 0) & (r["h2"] < 1)).all()
    for kw in (dict(InnerGS=True), dict(HCS=True), dict(XFA=True, NumXFA=2), dict(TH=True), dict(NLfactor=0.5, maxit=50)):
        ro = O.mrr3(Y, X, **kw)
        assert np.isfinite(ro["b"]).all(), kw


def test_gibbs_posterior_means_stable(tpod):
    """Monte-Carlo sanity of the Gibbs oracle: two seeds agree within MC error on h2 and GEBVs."""
    y, gen = tpod
    X = gen.astype(np.float64)
    for m in O.GIBBS_MODELS:
        a = O.gibbs(m, y, X, it=1500, bi=500, seed=11)
        b = O.gibbs(m, y, X, it=1500, bi=500, seed=12)
        assert abs(a["h2"] - b["h2"]) < 0.08, m
        assert np.corrcoef(a["hat"], b["hat"])[0, 1] > 0.97, m


def test_kmup_ratio_form_equivalent(tpod):
    """On tpod the literal exp(C||e||^2) form does not underflow, so both forms take the same branches."""
    y, gen = tpod
    X = gen.astype(np.float64)
    p = X.shape[1]
    xx = (X * X).sum(0)
    e = y - y.mean()
    L = np.full(p, 50.0)
    a = O.kmup(X, np.zeros(p), np.ones(p), xx, e, L, 0.03, 0.9, seed=4, ratio_form=False)
    b = O.kmup(X, np.zeros(p), np.ones(p), xx, e, L, 0.03, 0.9, seed=4, ratio_form=True)
    assert (a["d"] == b["d"]).mean() > 0.99
    w = O.wgr(y, X, it=300, bi=100, pi=0.9, iv=True, seed=2)
    assert np.isfinite(w["b"]).all() and 0 < w["d"].mean() < 1


def _np_second_panel(model, y, X, it, df=10.0, R2=0.5, Pi=0.75):
    """Independent numpy (float64) restatement of emDE :250, emML :463, emBCpi :1502 and lasso :1463 for a fixed sweep count."""
    n, p = X.shape
    xx = (X * X).sum(0); vx = X.var(0, ddof=1); vy = y.var(ddof=1)
    mu = y.mean(); b = np.zeros(p); e = y - mu; d = np.zeros(p)
    perms = O.perm(p, it)
    if model == "emDE":
        xx = np.where(xx == 0, 0.1, xx); cxx = vx.sum() * (1 - R2) / R2; L = np.full(p, p + cxx)
    elif model == "emML":
        MSx = vx.sum(); L = MSx
    elif model == "emBCpi":
        Pi = min(Pi, 1 - Pi); prior = Pi; MSx = vx.sum() * Pi * (1 - Pi); Sa = R2 * (df + 2) * vy / MSx
        Se = (1 - R2) * (df + 2) * vy; ve, va = Sa, Se; L = ve / va; Pi0 = (1 - Pi) / Pi
    else:
        L = xx.mean() / p; yx = np.zeros(p)
    for i in range(it):
        order = perms[i] if model in ("emDE", "emML") else range(p)
        C = -0.5 / np.sqrt(ve) if model == "emBCpi" else 0.0
        for j in order:
            x = X[:, j]; b0 = b[j]; g = x @ e + xx[j] * b0
            if model == "emDE":
                b[j] = g / (L[j] + xx[j])
            elif model == "emML":
                b[j] = g / (xx[j] + L)
            elif model == "emBCpi":
                b1 = g / (xx[j] + L)
                n1 = ((e - x * (b1 - b0)) ** 2).sum(); n2 = ((e + x * b0) ** 2).sum()
                d[j] = 1 / (1 + Pi0 * np.exp(C * (n2 - n1))); b[j] = b1 * d[j]
            else:
                yx[j] = g
                b[j] = max((g - L) / xx[j], 0.0) if g > 0 else min((g + L) / xx[j], 0.0)
            e -= x * (b[j] - b0)
        if model == "emBCpi":
            dm = d.mean(); Pi = ((1 - dm) * p + prior * df) / (p + df); Pi0 = (1 - Pi) / Pi
            MSx = vx.sum() * Pi * (1 - Pi); Sa = R2 * (df + 2) * vy / MSx
            ve = (e @ e + Se) / (n + df); va = (b @ b + Sa) / (p + df) / (dm - Pi); L = ve / va
        if model == "lasso":
            L = 2 * np.sqrt(abs(2 * (np.abs(yx) - np.abs(b * xx)).sum() / p))
        mu += e.mean(); e -= e.mean()
        if model == "emDE":
            Ve = e @ y / (n - 1); Vb = b * b + Ve / (xx + L + 0.0001); L = np.sqrt(cxx * Ve / Vb)
        if model == "emML":
            ve = (y - mu) @ e / n; vb = (y - mu) @ ((y - mu) - e) / (n * MSx); L = ve / vb
    out = dict(mu=mu, b=b, hat=X @ b + mu)
    if model == "emDE":
        out.update(Vb=Vb, Ve=Ve, h2=Vb.sum() / (Vb.sum() + Ve))
    elif model == "emML":
        out.update(Vb=vb, Va=vb * MSx, Ve=ve, h2=vb * MSx / (vb * MSx + ve))
    elif model == "emBCpi":
        out.update(d=d, pi=Pi, Vg=va * MSx, Va=va, Ve=ve, h2=1 - ve / vy)
    else:
        out.update(Lmb=L, h2=1 - (e @ y / (n - 1)) / vy)
    return out


@pytest.mark.parametrize("model", ["emDE", "emML", "emBCpi", "lasso"])
def test_second_panel_vs_numpy(tpod, model):
    y, gen = tpod
    X = gen.astype(np.float64)
    ref = _np_second_panel(model, y.astype(np.float32).astype(np.float64), X, it=12)
    r = O.em(model, y, X, it=12, use_double=True)
    assert r["its"] == 12
    for key, v in ref.items():
        np.testing.assert_allclose(r[key], v, rtol=2e-5, atol=1e-9, err_msg=key)


@pytest.mark.parametrize("model", ["BayesL", "BayesCpi", "BayesDpi"])
def test_second_gibbs_panel_sane(tpod, model):
    """BayesL / BayesCpi / BayesDpi of the oracle: finite, reproducible per seed, seed-dependent, and their fitted values agree
    with BayesRR's (same data, same prior scale) far better than chance."""
    y, gen = tpod
    X = gen.astype(np.float64)
    a = O.gibbs(model, y, X, it=400, bi=100, seed=5)
    b = O.gibbs(model, y, X, it=400, bi=100, seed=5)
    c = O.gibbs(model, y, X, it=400, bi=100, seed=6)
    assert np.array_equal(a["b"], b["b"]) and not np.array_equal(a["b"], c["b"])
    assert np.all(np.isfinite(a["b"])) and 0 < a["h2"] < 1 and a["ve"] > 0
    rr = O.gibbs("BayesRR", y, X, it=400, bi=100, seed=7)
    assert np.corrcoef(a["hat"], rr["hat"])[0, 1] > 0.9
    if model != "BayesL":
        assert 0 < a["pi"] < 1 and a["d"].min() >= 0 and a["d"].max() <= 1


def _np_first_panel(model, y, X, it, df=10.0, R2=0.5, Pi=0.75, alpha=0.02):
    """Independent numpy (float64) restatement of emBA :80, emBB :131, emBC :190, emBL :357 and emEN :400 for a fixed sweep
    count; the spike-slab likelihood ratio uses the closed form |e2|^2 - |e1|^2 = b1 (2 g + xx (2 b0 - b1)),
    not the two temporaries the oracle forms."""
    n, p = X.shape
    xx = (X * X).sum(0); vx = X.var(0, ddof=1); vy = y.var(ddof=1)
    mu = y.mean(); b = np.zeros(p); e = y - mu; d = np.zeros(p); vb = np.ones(p)
    perms = O.perm(p, it)
    Se = (1 - R2) * (df + 2) * vy
    if model in ("emBA", "emBB"):
        if model == "emBB":
            Pi = min(Pi, 1 - Pi)
        MSx = vx.sum() * (Pi if model == "emBB" else 1.0); Sb = R2 * (df + 2) * vy / MSx; ve = 1.0; L = ve / vb
    elif model == "emBC":
        Pi = min(Pi, 1 - Pi); MSx = vx.sum() * Pi * (1 - Pi); Sa = R2 * (df + 2) * vy / MSx; ve, va = Sa, Se; L = ve / va
    elif model == "emBL":
        cxx = xx.mean(); L1 = cxx * ((1 - R2) / R2) * alpha * 0.5; L2 = cxx * ((1 - R2) / R2) * (1 - alpha)
    else:
        cxx = vx.sum() * (1 - R2) / R2; Sy = np.sqrt(vy); L = cxx; L1 = 0.5 * L * alpha * Sy; L2 = L * (1 - alpha)
        tr = (1.0 / (xx + L)).sum()
    Pi0 = (1 - Pi) / Pi
    for i in range(it):
        if model in ("emBB", "emBC"):
            C = -0.5 / np.sqrt(ve)
        for j in perms[i]:
            x = X[:, j]; b0 = b[j]; g = x @ e; ols = g + xx[j] * b0
            if model == "emBA":
                b1 = ols / (xx[j] + L[j]); b[j] = b1; vb[j] = (Sb + b1 * b1) / (df + 1)
                e -= 2 * x * (b1 - b0)  # the residual is updated twice (:108, :111)
                continue
            if model in ("emBB", "emBC"):
                b1 = ols / (xx[j] + (L[j] if model == "emBB" else L))
                d[j] = 1 / (1 + Pi0 * np.exp(C * b1 * (2 * g + xx[j] * (2 * b0 - b1)))); b[j] = b1 * d[j]
                if model == "emBB":
                    vb[j] = (Sb + b[j] ** 2) / (df + 1)
            elif model == "emBL":
                half = 0.5 * ols / (xx[j] + cxx); G = 0.5 * (ols - np.sign(ols if ols != 0 else -1) * L1) / (L2 + xx[j])
                b[j] = G + half if (G > 0) == (ols > 0) and G != 0 else half
            else:
                b[j] = max((ols - L1) / (L2 + xx[j]), 0.0) if ols > 0 else min((ols + L1) / (L2 + xx[j]), 0.0)
            e -= x * (b[j] - b0)
        if model in ("emBA", "emBB"):
            ve = (e @ e + Se) / (n + df); L = ve / vb
        elif model == "emBC":
            ve = (e @ e + Se) / (n + df); va = (b @ b + Sa) / (p + df) / (d.mean() - Pi); L = ve / va
        mu += e.mean(); e -= e.mean()
        if model == "emEN":
            Ve = e @ y / (n - 1); Va = (b @ b + tr * Ve) / p; L = Ve / Va; L1 = 0.5 * L * alpha * Sy; L2 = L * (1 - alpha)
    out = dict(mu=mu, b=b, hat=X @ b + mu)
    if model in ("emBA", "emBB"):
        out.update(Vb=vb, Ve=ve, h2=1 - ve / vy)
    if model in ("emBB", "emBC"):
        out["d"] = d
    if model == "emBC":
        out.update(Vg=va * MSx, Va=va, Ve=ve, h2=1 - ve / vy)
    if model == "emBL":
        out["h2"] = 1 - e.var(ddof=1) / vy
    if model == "emEN":
        out.update(Va=Va * cxx, Ve=Ve, h2=Va * cxx / (Va * cxx + Ve))
    return out


@pytest.mark.parametrize("model", ["emBA", "emBB", "emBC", "emBL", "emEN"])
def test_first_panel_vs_numpy(tpod, model):
    y, gen = tpod
    X = gen.astype(np.float64)
    ref = _np_first_panel(model, y.astype(np.float32).astype(np.float64), X, it=10)
    r = O.em(model, y, X, it=10, use_double=True)
    for key, v in ref.items():
        np.testing.assert_allclose(r[key], v, rtol=2e-5, atol=1e-9, err_msg=key)


def test_two_design_oracle_matches_reference_executed_golden(tpod):
    """The oracle's emML2 against tests/golden/tpod_two_design.npz (emML2 run by the reference's own source, oracle/make_golden.py)."""
    y, X = tpod
    g = np.load(os.path.join(GOLDEN, "tpod_two_design.npz"))
    X1, X2 = X[:, :200].astype(np.float64), X[:, 200:].astype(np.float64)
    for tag, kw in (("plain", {}), ("weighted", dict(D1=g["D1"], D2=g["D2"]))):
        r = O.two_design("emML2", y, X1, X2, **kw)
        for key in ("mu", "b1", "b2", "Vb1", "Vb2", "Ve", "u1", "u2", "h2", "hat"):
            ref = g[tag + "__" + key]
            assert np.abs(np.asarray(r[key]) - ref).max() <= 1e-3 * max(np.abs(ref).max(), 1e-30), (tag, key)



def test_kmup2_repeated_rows_equal_replicated_rows(tpod):
    """KMUP2 on a Use with repeats (what wgr(bag, rp = TRUE) passes, R/wgr.R:68) is the sweep over a matrix that physically holds every
    repeated row once per draw (Rcpp20260726ai.cpp:51-60 builds exactly that H / e0): the same arithmetic in the same order, equal up to the
    float rounding of xx(j) * bg (bg differs between the two calls and is folded into xx)."""
    y, X = tpod
    n, p = X.shape
    rng = np.random.default_rng(21)
    use = np.sort(rng.integers(0, n, size=int(0.7 * n)))
    assert np.unique(use).size < use.size
    Xr = np.asfortranarray(X[use])
    b0 = rng.normal(size=p) * 0.01
    E = y - y.mean() - X @ b0
    xx = (X * X).sum(0) * (use.size / n)
    L = np.full(p, 40.0)
    for pi in (0.0, 0.3):
        a = O.kmup2(X, use.astype(np.float64), b0, np.ones(p), xx, E, L, 0.03, pi, seed=5)
        # the replicated matrix has n0 = nuse rows, so bg = 1 there: hand it xx * (n / nuse) to keep the same denominators
        r = O.kmup2(Xr, np.arange(use.size, dtype=np.float64), b0, np.ones(p), xx * np.float32(n / use.size), E[use], L, 0.03, pi, seed=5)
        assert np.array_equal(a["d"], r["d"])
        assert np.abs(a["b"] - r["b"]).max() <= 2e-6 * np.abs(a["b"]).max() and np.abs(a["e"] - r["e"]).max() <= 2e-6 * np.abs(a["e"]).max()
