#!/usr/bin/env python
"""Per-sweep time of every solver of the emCV / mcmcCV panels at the headline shape (n = 50k x p = 50k, int8, one B200), and one
emCV hold-out at the config-4 shape (n = 10k x p = 50k: ten solvers to their own stopping rules on one packed store).
Not the bench line (bench.py is); output goes to profiles/ as the measurement of SURVEY 8(f) ranks 1-2."""
import json
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
out = {"shape": [50000, 50000], "em_ms_per_sweep": {}, "gibbs_ms_per_sweep": {}}


def timed(fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = fn()
    torch.cuda.synchronize()
    return r, time.perf_counter() - t0


n, p = 50000, 50000
Xt, y = bench.synth_gpu(n, p, bench.SEED, dev)
g = bw.Genotypes(device=0)
g.load(Xt)
for m in bw.api.EMCV_MODELS:  # begin / sweeps / end split: 3 untimed sweeps (kernel load, workspaces), then 10 timed ones
    st = bw.EmStepper(m, y, g)
    st.sweeps(3)
    _, t = timed(lambda: st.sweeps(10))
    st.end()
    out["em_ms_per_sweep"][m] = round(t / 10 * 1e3, 3)
for m in bw.api.MCMCCV_MODELS:  # whole fits (the Gibbs entry has no split): device time of the kernel classes from the handle's CUDA events
    bw.gibbs_fit(m, y, g, it=2, bi=1, seed=1)
    g.profile(True)
    bw.gibbs_fit(m, y, g, it=20, bi=1, seed=1)
    pr = g.profile_read()
    g.profile(False)
    out["gibbs_ms_per_sweep"][m] = {k: round(v["ms"] / 20, 3) for k, v in pr.items()}
    out["gibbs_ms_per_sweep"][m]["total"] = round(sum(v["ms"] for v in pr.values()) / 20, 3)
print(json.dumps(out), flush=True)
g.close()
del Xt
torch.cuda.empty_cache()

if "noemcv" in sys.argv:
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "panel_bench_sweeps.json"), "w"), indent=1)
    sys.exit(0)
# one emCV hold-out, config-4 shape, host matrices in and out (what an R caller hands over)
n, p = 10000, 50000
Xt, y = bench.synth_gpu(n, p, bench.SEED, dev)
X = Xt.t().contiguous().cpu().numpy()  # n x p int8 on the host
del Xt
torch.cuda.empty_cache()
t0 = time.perf_counter()
try:
    cv = bw.emCV(y, X, k=5, n=1, seed=1)
    out["emCV_one_holdout_10k_x_50k"] = {"seconds": round(time.perf_counter() - t0, 2), "predictive_ability": cv}
except bw.BwgrError as err:  # recorded, not hidden: a solver of the panel that leaves the fixed-point range at this shape
    out["emCV_one_holdout_10k_x_50k"] = {"error": str(err), "seconds": round(time.perf_counter() - t0, 2)}
print(json.dumps(out), flush=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "panel_bench.json"), "w"), indent=1)
