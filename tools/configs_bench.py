#!/usr/bin/env python
"""Full-size runs of BASELINE.json configs 2-4 on one B200 (reduced iteration counts, per-sweep figures).
Not the bench line (bench.py is); the results go to profiles/ as evidence that every section-8 row runs at the named shapes."""
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
out = {}
which = sys.argv[1:] or ["cfg2", "cfg3", "cfg4"]


def timed(fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = fn()
    torch.cuda.synchronize()
    return r, time.perf_counter() - t0


if "cfg2" in which or "cfg4" in which:
    n, p = 10000, 50000
    Xt, y = bench.synth_gpu(n, p, bench.SEED, dev)
    g = bw.Genotypes(device=0)
    g.load(Xt)
    if "cfg2" in which:
        # config 2: wgr BayesB (pi > 0, iv = TRUE), n = 10k x p = 50k; 20k iterations in the full run, 60 + 260 here
        _, t1 = timed(lambda: bw.wgr(y, g, it=60, bi=20, pi=0.95, iv=True, seed=1))
        r, t2 = timed(lambda: bw.wgr(y, g, it=260, bi=60, pi=0.95, iv=True, seed=1))
        per = (t2 - t1) / 200
        out["cfg2_wgr_BayesB_10k_x_50k"] = {"ms_per_iteration": per * 1e3, "marker_updates_per_s": p / per,
                                           "projected_s_for_20000_iterations": per * 20000, "mean_d": float(np.mean(r["d"])),
                                           "Ve": r["Ve"], "cor_hat_y": float(np.corrcoef(r["hat"], y)[0, 1])}
        # same chain as standalone BayesB(), 4 chains at once
        _, t1 = timed(lambda: bw.BayesB(y, g, it=40, bi=10, nchains=4, seed=2))
        r, t2 = timed(lambda: bw.BayesB(y, g, it=140, bi=10, nchains=4, seed=2))
        per = (t2 - t1) / 100
        out["cfg2_BayesB_4chains_10k_x_50k"] = {"ms_per_sweep_all_chains": per * 1e3, "marker_updates_per_s": 4 * p / per,
                                               "h2": np.asarray(r["h2"]).tolist()}
        print(json.dumps(out), flush=True)
    if "cfg4" in which:
        # config 4: 5 folds x 20 traits = 100 emBC fits sharing the genotypes (row masks), one GPU's share is all 100 here
        rng = np.random.default_rng(1)
        k = 20
        Y = np.empty((n, k))
        Xh = None
        for t in range(k):
            Y[:, t] = y * np.sqrt(0.5) + rng.normal(size=n) * np.sqrt(0.5)
        perm = rng.permutation(n)
        Yall = np.repeat(Y, 5, axis=1)
        mask = np.ones((n, 5 * k), dtype=bool)
        for t in range(k):
            for f in range(5):
                mask[perm[f * n // 5:(f + 1) * n // 5], 5 * t + f] = False
        _, t1 = timed(lambda: bw.em_fit("emBC", Yall, g, it=2, row_mask=mask))
        r, t2 = timed(lambda: bw.em_fit("emBC", Yall, g, it=6, row_mask=mask))
        per = (t2 - t1) / 4
        out["cfg4_emBC_100fits_10k_x_50k"] = {"ms_per_sweep_all_fits": per * 1e3, "marker_updates_per_s": 100 * p / per,
                                             "projected_s_for_200_sweeps": per * 200, "h2_mean": float(np.mean(r["h2"]))}
        print(json.dumps(out), flush=True)
    g.close()
    del Xt
    torch.cuda.empty_cache()

if "cfg3" in which:
    # config 3: MRR3, n = 50k x p = 50k, k = 20 traits
    n, p, k = 50000, 50000, 20
    Xt, y = bench.synth_gpu(n, p, bench.SEED, dev)
    rng = np.random.default_rng(2)
    Y = np.empty((n, k))
    for t in range(k):
        Y[:, t] = y * np.sqrt(0.5) + rng.normal(size=n) * np.sqrt(0.5)
    g = bw.Genotypes(device=0)
    g.load(Xt)
    _, t1 = timed(lambda: bw.MRR3(Y, g, maxit=2))
    r, t2 = timed(lambda: bw.MRR3(Y, g, maxit=8))
    per = (t2 - t1) / 6
    out["cfg3_MRR3_50k_x_50k_k20"] = {"ms_per_sweep": per * 1e3, "marker_trait_updates_per_s": k * p / per,
                                     "h2": r["h2"].tolist(), "cnvB": r["cnvB"].tolist(), "GC01": float(r["GC"][0, 1])}
    print(json.dumps(out), flush=True)
    g.close()

json.dump(out, open(os.path.join(ROOT, "gpurun_out", "configs_r1.json"), "w"), indent=1)
