"""Per-kernel-class time of the config-4 pattern (100 row-masked emBC fits, 10k x 50k) -- where a sweep of all fits goes."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, bwgr_b200 as bw
dev = torch.device("cuda", 0)
n, p, ktr, folds = 10000, int(os.environ.get('CFG4_P', 50000)), 20, 5
Xt, y = bench.synth_gpu(n, p, bench.SEED, dev)
rng = np.random.default_rng(2)
Y = np.stack([y * np.sqrt(0.5) + rng.normal(size=n) * np.sqrt(0.5) for _ in range(ktr)], axis=1)
perm = rng.permutation(n)
Yall = np.repeat(Y, folds, axis=1)
mask = np.ones((n, folds * ktr), dtype=bool)
for t in range(ktr):
    for f in range(folds):
        mask[perm[f * n // folds:(f + 1) * n // folds], folds * t + f] = False
g = bw.Genotypes(device=0)
g.load(Xt)
bw.em_fit("emBC", Yall, g, it=2, row_mask=mask)
for it in ((3,) if 'CFG4_P' in os.environ else (3, 9)):
    g.profile(True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    bw.em_fit("emBC", Yall, g, it=it, row_mask=mask)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    pr = g.profile_read(); g.profile(False)
    print("it=%d wall %.3f s" % (it, dt), {k: (round(v["ms"], 2), v["launches"]) for k, v in pr.items()}, flush=True)
