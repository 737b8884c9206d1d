"""Generate the committed fixtures under tests/golden/ (run once, in the build container).

TEST INFRASTRUCTURE ONLY.  Provenance of every fixture:
  tpod.npz          decoded from /root/reference/data/tpod.RData (the reference's only bundled data,
                    man/tpod.Rd) with oracle/rdata.py -- real reference data, bit-exact.
  perm_kat.json     std::shuffle/std::mt19937 marker orders; the first three (p=10) and the p=376 head
                    are the known answers recorded in SURVEY.md 8a / BASELINE.md 5 (libstdc++ 13).
  tpod_em.npz       REFERENCE-EXECUTED (<model>_ref__*): the ten EM solvers of /root/reference/src/Rcpp20260726ai.cpp, compiled
                    unmodified into oracle/_ref/libbwgr_ref.so against the stand-in RcppEigen headers (oracle/shim/), run on
                    tpod.  Beside them the oracle's float32 / float64 recipes (<model>_f32__*, <model>_f64__*), kept to bound
                    float noise.  tests/test_ref_pin.py checks oracle == reference on these and many more inputs.
  tpod_mrr3.npz     REFERENCE-EXECUTED: MRR3 (float64, RcppEigen20230423.cpp:318-701 compiled the same way) on a fixed
                    synthetic 3-trait Y over the tpod genotypes.
Usage:  python oracle/make_golden.py   (needs /root/reference; tests never do)
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import oracle as O  # noqa: E402
import ref as R  # noqa: E402
from rdata import read_rdata  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden")


def synth_traits(gen, k=3, seed=20261018):
    rng = np.random.default_rng(seed)
    p = gen.shape[1]
    B = rng.normal(size=(p, k)) * (rng.random((p, k)) < 0.1)
    G = gen @ B
    return G / G.std(0) + rng.normal(size=(gen.shape[0], k))


def main():
    os.makedirs(OUT, exist_ok=True)
    d = read_rdata("/root/reference/data/tpod.RData")
    y = np.asarray(d["y"], dtype=np.float64)
    gen = np.asarray(d["gen"])
    assert gen.shape == (196, 376) and set(np.unique(gen)) == {0, 1, 2}
    np.savez_compressed(os.path.join(OUT, "tpod.npz"), y=y, gen=gen.astype(np.int8),
                        fam=np.asarray(d["fam"]).astype(np.int32), chr=np.asarray(d["chr"]).astype(np.int32))
    kat = {"p10_iters0_2": O.perm(10, 3).tolist(), "p376_iter0_head8": O.perm(376, 1)[0, :8].tolist(),
           "p376_iter199_head8": O.perm(376, 200)[199, :8].tolist()}
    json.dump(kat, open(os.path.join(OUT, "perm_kat.json"), "w"), indent=1)
    genf = gen.astype(np.float64)
    em = {"provenance": np.array("reference-executed")}
    for m in O.EM_MODELS:
        for key, v in R.em(m, y, genf).items():
            em[m + "_ref__" + key] = np.asarray(v)
        for dbl in (False, True):
            r = O.em(m, y, genf, use_double=dbl)
            tag = m + ("_f64" if dbl else "_f32")
            for key, v in r.items():
                em[tag + "__" + key] = np.asarray(v)
    np.savez_compressed(os.path.join(OUT, "tpod_em.npz"), **em)
    Y = synth_traits(genf)
    r = R.mrr3(Y, genf)
    np.savez_compressed(os.path.join(OUT, "tpod_mrr3.npz"), Y=Y, **{k: np.asarray(v) for k, v in r.items()})
    two_design_golden(y, genf)
    print("wrote", sorted(os.listdir(OUT)))


def two_design_golden(y, genf):
    """emML2 (Rcpp20260726ai.cpp:1221-1305) executed by the reference's own source on tpod split into two designs (markers 1-200 and
    201-376), without and with marker weights.  `python oracle/make_golden.py two_design` writes this file alone."""
    X1, X2 = genf[:, :200], genf[:, 200:]
    rng = np.random.default_rng(20261018)
    D1, D2 = rng.uniform(0.5, 2.0, 200), rng.uniform(0.5, 2.0, 176)
    out = {"provenance": np.array("reference-executed"), "split": np.array(200), "D1": D1, "D2": D2}
    for tag, kw in (("plain", {}), ("weighted", dict(D1=D1, D2=D2))):
        for key, v in R.two_design("emML2", y, X1, X2, **kw).items():
            out[tag + "__" + key] = np.asarray(v)
    np.savez_compressed(os.path.join(OUT, "tpod_two_design.npz"), **out)


if __name__ == "__main__":
    if sys.argv[1:] == ["two_design"]:
        t = np.load(os.path.join(OUT, "tpod.npz"))
        two_design_golden(t["y"], t["gen"].astype(np.float64))
    else:
        main()
