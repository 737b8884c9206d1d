#!/usr/bin/env python
"""Row-sharded emRR sweep timing: n_per_rank x p int8 per GPU, synthetic data generated on each device.
usage: torchrun --nproc-per-node N tools/dist_time.py n_per_rank p sweeps"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import bwgr_b200 as bw  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
nl, p, sweeps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
Xt, y = bench.synth_gpu(nl, p, bench.SEED + rank, dev)
g = bw.Genotypes(device=local, path=bw.PATH_BLOCKED)
if world > 1:
    g.enable_row_sharding()
stream = torch.cuda.Stream(device=dev)
g.set_stream(stream.cuda_stream)
g.load(Xt)
st = bw.EmStepper("emRR", y, g)
st.sweeps(3)
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
st.sweeps(sweeps)
e1.record(stream)
torch.cuda.synchronize(); dist.barrier()
ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
g.profile(True)
st.sweeps(5)
pr = g.profile_read()
out = st.end()
if rank == 0:
    per = ms.item() / sweeps
    print(json.dumps({"world": world, "n_total": nl * world, "n_per_rank": nl, "p": p, "ms_per_sweep": per,
                      "marker_updates_per_s": p / (per * 1e-3), "genotype_GB_per_s_aggregate": nl * world * p / (per * 1e-3) / 1e9,
                      "kernel_ms": {k: v["ms"] / max(1, v["launches"]) for k, v in pr.items()}, "h2": out["h2"]}), flush=True)
g.close()
dist.destroy_process_group()
