"""MRR3 general path (csrc/mrr_gen.cu) at a chosen shape: a few sweeps on synthetic genotypes with missing phenotypes, wall time per
call.  For ncu: `ncu -k regex:mrr_gen_sweep -c 1 --set full --import-source on python tools/mrr_gen_probe.py 50000 50000 20 2`."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bwgr_b200 as bw  # noqa: E402

n, p, k, its = (int(v) for v in (sys.argv[1:5] + ["50000", "50000", "20", "2"][len(sys.argv) - 1:]))
gen = torch.Generator(device="cuda").manual_seed(1)
X = torch.randint(0, 3, (p, n), generator=gen, device="cuda", dtype=torch.int8)
rng = np.random.default_rng(3)
Y = rng.normal(size=(n, k))
Y[rng.random(Y.shape) < 0.3] = np.nan
torch.cuda.synchronize()
with bw.Genotypes(device=0) as g:
    g.load(X)
    for rep in range(2):
        t0 = time.perf_counter()
        out = bw.MRR3(Y, g, maxit=its, tol=0.0)
        print("call %d: %d sweeps in %.3f s; h2 mean %.4f" % (rep, out["Its"], time.perf_counter() - t0, out["h2"].mean()), flush=True)
