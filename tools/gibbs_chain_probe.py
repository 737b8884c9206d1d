#!/usr/bin/env python
"""Wall time per Gibbs iteration of the spike-slab samplers on the blocked family (n = 10k x p = 50k, int8, one B200): whole fits of
60 and 20 iterations, the difference / 40 (set-up and the final fitted values cancel).  A/B of the chain variants via BWGR_LIB."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import bwgr_b200 as bw  # noqa: E402

dev = torch.device("cuda", 0)
n, p = 10000, 50000
Xt, y = bench.synth_gpu(n, p, bench.SEED, dev)
g = bw.Genotypes(device=0, path=bw.PATH_BLOCKED)
g.load(Xt)
y = np.asarray(y.cpu() if hasattr(y, "cpu") else y, dtype=np.float64)
for m in ("BayesB", "BayesC", "BayesDpi"):
    fn = getattr(bw, m)
    fn(y, g, it=8, bi=2)
    ts = {}
    for it in (20, 60):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn(y, g, it=it, bi=2)
        torch.cuda.synchronize()
        ts[it] = time.perf_counter() - t0
    print("%s: %.3f ms per iteration (blocked family, n=%d p=%d)" % (m, (ts[60] - ts[20]) / 40 * 1e3, n, p), flush=True)
