#!/usr/bin/env python
"""BASELINE config 4: 5 folds x 20 traits = 100 emBC fits on n=10k x p=50k, systems sharded over the ranks (no data-path
collective; bwgr_b200.dist.fit_sharded).  usage: torchrun --nproc-per-node N tools/cv_sharded.py [sweeps]"""
import functools
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import bwgr_b200 as bw  # noqa: E402
from bwgr_b200 import dist as bd  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
sweeps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
n, p, k = 10000, 50000, 20
Xt, y = bench.synth_gpu(n, p, bench.SEED, dev)  # same genotypes on every rank (replicated store)
rng = np.random.default_rng(1)
Y = np.stack([y * np.sqrt(0.5) + rng.normal(size=n) * np.sqrt(0.5) for _ in range(k)], axis=1)
perm = rng.permutation(n)
Yall = np.repeat(Y, 5, axis=1)
mask = np.ones((n, 5 * k), dtype=bool)
for t in range(k):
    for f in range(5):
        mask[perm[f * n // 5:(f + 1) * n // 5], 5 * t + f] = False
g = bw.Genotypes(device=local)
g.load(Xt)
fit = functools.partial(bw.em_fit, "emBC", gen=g)
bd.fit_sharded(lambda Yc, row_mask=None, **kw: fit(Yc, row_mask=row_mask, it=1), Yall, row_mask=mask)  # warm-up
times = []
for its in (2, 2 + sweeps):
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    out = bd.fit_sharded(lambda Yc, row_mask=None, **kw: fit(Yc, row_mask=row_mask, it=its), Yall, row_mask=mask)
    torch.cuda.synchronize(); dist.barrier()
    times.append(time.perf_counter() - t0)
per = (times[1] - times[0]) / sweeps
if rank == 0:
    print(json.dumps({"world": world, "fits": 100, "fits_per_gpu": [e - s for s, e in bd.partition(100, world)], "n": n, "p": p,
                      "ms_per_sweep_all_fits": per * 1e3, "marker_updates_per_s": 100 * p / per,
                      "projected_s_for_200_sweeps": per * 200, "b_shape": list(out["b"].shape),
                      "h2_mean": float(np.mean(out["h2"]))}), flush=True)
g.close()
dist.destroy_process_group()
