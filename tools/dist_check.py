#!/usr/bin/env python
"""Row-sharded emRR / emBC / BayesRR over the ranks of one node vs the single-GPU fit of the same data (run under torchrun).
Also times row-sharded sweeps.  usage: torchrun --nproc-per-node N tools/dist_check.py [n p]"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bwgr_b200 as bw  # noqa: E402
from conftest import synth  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n, p = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (4096, 1500)
X, y = synth(n, p, seed=5)
rows = np.array_split(np.arange(n), world)[rank]
ok = True
g = bw.Genotypes(device=local, path=bw.PATH_BLOCKED)
g.enable_row_sharding()
g.load(np.ascontiguousarray(X[rows]))
for model, it in (("emRR", 10), ("emBC", 6), ("emBL", 5), ("emBA", 5), ("emEN", 6)):
    out = bw.em_fit(model, y[rows], g, it=it)
    if rank == 0:
        with bw.Genotypes(X, device=local, path=bw.PATH_BLOCKED) as g1:
            ref = bw.em_fit(model, y, g1, it=it)
        eb = np.abs(out["b"] - ref["b"]).max() / np.abs(ref["b"]).max()
        eh = np.abs(out["hat"] - ref["hat"][rows]).max() / np.abs(ref["hat"]).max()
        eh2 = abs(out["h2"] - ref["h2"])
        print("%s world=%d: max|db|/max|b| %.2e  hat %.2e  h2 %.2e (%.5f vs %.5f)" % (model, world, eb, eh, eh2, out["h2"], ref["h2"]), flush=True)
        ok = ok and eb < 1e-4 and eh < 1e-4 and eh2 < 1e-4
    # every rank holds the same b
    t = torch.tensor(out["b"], device="cuda")
    t0 = t.clone()
    dist.broadcast(t0, src=0)
    same = bool((t == t0).all().item())
    if not same:
        print("rank %d: b differs from rank 0 for %s" % (rank, model), flush=True)
        ok = False
for model in ("BayesRR", "BayesB"):
    out = bw.gibbs_fit(model, y[rows], g, it=30, bi=5, seed=3)
    if rank == 0:
        with bw.Genotypes(X, device=local, path=bw.PATH_BLOCKED) as g1:
            ref = bw.gibbs_fit(model, y, g1, it=30, bi=5, seed=3)
        eb = np.abs(out["b"] - ref["b"]).max() / np.abs(ref["b"]).max()
        print("%s world=%d (same Philox streams): max|db|/max|b| %.2e  h2 %.5f vs %.5f" % (model, world, eb, out["h2"], ref["h2"]), flush=True)
        ok = ok and eb < 1e-3
dist.barrier()
if len(sys.argv) > 3:  # timing: emRR sweeps at the given shape, rows sharded
    st = bw.EmStepper("emRR", y[rows], g)
    st.sweeps(3)
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    st.sweeps(int(sys.argv[3]))
    torch.cuda.synchronize(); dist.barrier()
    dt = time.perf_counter() - t0
    st.end()
    if rank == 0:
        print("row-sharded emRR n=%d (%d per rank) x p=%d: %.3f ms per sweep" % (n, len(rows), p, 1e3 * dt / int(sys.argv[3])), flush=True)
g.close()
flag = torch.tensor([0 if ok else 1], device="cuda")
dist.all_reduce(flag)
dist.destroy_process_group()
if rank == 0:
    print("DIST CHECK", "OK" if flag.item() == 0 else "FAILED", flush=True)
sys.exit(0 if flag.item() == 0 else 1)
