"""A few sweeps of one solver of every family on tpod through the blocked path: the workload run under compute-sanitizer."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bwgr_b200 as bw
d = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "tpod.npz"))
y, gen = d["y"].astype(np.float64), d["gen"]
with bw.Genotypes(gen, path=bw.PATH_BLOCKED) as g:
    a = bw.emRR(y, g, it=3); print("emRR", a["h2"])
    a = bw.emBC(y, g, it=2); print("emBC", a["h2"])
    a = bw.BayesB(y, g, it=4, bi=1, seed=1); print("BayesB", a["h2"])
    Y = np.stack([y, y[::-1]], 1)
    a = bw.MRR3(Y, g, maxit=2); print("MRR3", a["h2"])
with bw.Genotypes(gen, path=bw.PATH_SMALL_N) as g:
    a = bw.emRR(y, g, it=2); print("emRR small-n", a["h2"])
print("done")
