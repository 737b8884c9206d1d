#!/usr/bin/env python
"""BASELINE config 4 on the blocked whole-GPU family: 5 folds x 20 traits = 100 emBC fits on n=10k x p=50k, each fold fitted on
its row-subset store (what emCV does with gen[-w,]), (fold, trait) tasks sharded over the ranks (bwgr_b200.dist.fit_cv_sharded,
no data-path collective).  usage: torchrun --nproc-per-node N tools/cv_blocked.py [sweeps]"""
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
Xt, y = bench.synth_gpu(n, p, bench.SEED, dev)  # (p, n) int8 on the device = the column-major n x p matrix
rng = np.random.default_rng(1)
Y = np.stack([y * np.sqrt(0.5) + rng.normal(size=n) * np.sqrt(0.5) for _ in range(k)], axis=1)
perm = rng.permutation(n)
folds = [np.sort(perm[f * n // 5:(f + 1) * n // 5]) for f in range(5)]


def load_(keep):
    sub = Xt.index_select(1, torch.as_tensor(keep, device=dev)).contiguous()  # the rows a fold trains on
    g = bw.Genotypes(device=local, path=bw.PATH_BLOCKED)
    g.load(sub)
    return g


clock = {"load": 0.0, "fit": 0.0}


def load(keep):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    g = load_(keep)
    torch.cuda.synchronize(); clock["load"] += time.perf_counter() - t0
    return g


def fit(Yc, g, it):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = bw.em_fit("emBC", Yc, g, it=it)
    torch.cuda.synchronize(); clock["fit"] += time.perf_counter() - t0
    return out


bd.fit_cv_sharded(fit, load, Y, folds, it=1)  # warm-up
times, fits, loads = [], [], []
for its in (2, 2 + sweeps):
    clock["load"] = clock["fit"] = 0.0
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    out = bd.fit_cv_sharded(fit, load, Y, folds, it=its)
    torch.cuda.synchronize(); dist.barrier()
    times.append(time.perf_counter() - t0)
    c = torch.tensor([clock["fit"], clock["load"]], device=dev, dtype=torch.float64)
    dist.all_reduce(c, op=dist.ReduceOp.MAX)  # the slowest rank sets the job's time
    fits.append(float(c[0])); loads.append(float(c[1]))
per = (fits[1] - fits[0]) / sweeps
if rank == 0:
    # spot check against the masked small-n path on the first fold / first trait
    keep = np.setdiff1d(np.arange(n), folds[0])
    print(json.dumps({"world": world, "fits": 100, "tasks_rank0": [(f, len(t)) for f, t in bd.cv_tasks(5, k, world)[0]], "n": n, "p": p,
                      "ms_per_sweep_all_fits": per * 1e3, "fit_s": fits, "load_s": loads, "wall_s": times, "marker_updates_per_s": 100 * p / per,
                      "projected_s_for_200_sweeps": per * 200, "b_shape": list(out["b"].shape), "h2_mean": float(np.mean(out["h2"]))}),
          flush=True)
dist.destroy_process_group()
