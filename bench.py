#!/usr/bin/env python
"""bench.py -- headline benchmark of the marker-effect update loop (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W [--impl reference]

Workload (N=1): one emRR Gauss-Seidel sweep (Rcpp20260726ai.cpp:330-344) over synthetic
n=50,000 x p=50,000 int8 genotypes, k=1 -- the configuration the metric "marker-updates/sec & sweep ms
at n=50k,p=50k; genotype GB/s vs HBM peak" is quoted on.  A step = one full sweep (p marker updates:
Gram blocks + blocked sweep + hyper-parameter epilogue).
  value      marker-updates/s with genotypes, y and all state resident in HBM (CUDA events on the
             library's stream, max over ranks).
  e2e        the same metric through the public call a user makes, bwgr_b200.emRR(y, gen) on R's own input: the n x p
             float64 column-major matrix on the host (what _bWGR_emRR receives).  One warm-up fit, then >= 5 timed fits
             (median; every fit listed): exact narrowing to int8 by host threads into pinned staging overlapped with the
             H2D copy, column statistics, the reference's fixed 200 sweeps, GEBVs, D2H -- all inside the timed region.
             The same call on a host int8 matrix is reported beside it (int8_host_input).
  roofline   algorithmic bytes = n*p*1 B per sweep / mean duration of the dominant kernel
             (measured live with CUDA events around each kernel class), against MEASURED_PEAKS.json.
  cpu_baseline  the oracle (C++ restatement of the reference's single-threaded float32 RcppEigen path, pinned against the
             reference's own sources compiled with stand-in headers: tests/test_ref_pin.py) on a bounded sample of the same
             workload: full n, the first m markers, a few sweeps.
N>1 (torchrun, one rank per GPU): ONE fit whose individuals are sharded by rows over the ranks (n = N x 50,000; BASELINE config 5
pattern): the per-block exchange of the reduced partials runs inside the sweep kernel over NVLink peer memory, the Gram band and
five scalars are all-reduced with NCCL once per sweep.  value = N x p marker updates per sweep / sweep time (one marker update =
one marker visited on one 50,000-row shard; scaling "weak").  `row_sharded_parity` = that sharded fit against the single-GPU fit
of the same global data on rank 0 (max|db| / max|b|, expected 0: every cross-GPU sum is an integer sum).  The communication-free
mode (N independent fits, one per GPU) is kept as the extra key `replicas`.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

SEED = 20261018


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=50000)
    ap.add_argument("--p", type=int, default=50000)
    ap.add_argument("--model", default="emRR")
    ap.add_argument("--e2e-fits", type=int, default=7)
    ap.add_argument("--e2e-sweeps", type=int, default=200)
    ap.add_argument("--cpu-markers", type=int, default=2048)
    ap.add_argument("--cpu-sweeps", type=int, default=10)
    ap.add_argument("--cpu-sweeps-main", type=int, default=150, help="sweeps of the cpu_baseline sample of the main arm (~15 s)")
    ap.add_argument("--traffic", type=float, default=None, help="dram bytes per launch of the dominant kernel from an ncu capture (profiles/)")
    ap.add_argument("--missing", type=float, default=0.0,
                    help="--config 3 only: fraction of the phenotypes set to NaN at random (unbalanced traits: MRR3's general device path)")
    ap.add_argument("--config", type=int, default=1, choices=[1, 2, 3, 4, 5],
                    help="BASELINE.json configs[k-1]... 1 (default) = the metric's configuration (emRR 50k x 50k); 2 wgr BayesB 10k x 50k; "
                         "3 MRR3 50k x 50k x 20 traits; 4 5-fold x 20-trait emBC fits; 5 row shards of 62500 x 100k per GPU (500k x 100k over 8)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    return ap.parse_args()


def peaks():
    try:
        d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def synth_gpu(n, p, seed, dev):
    """SURVEY 8d generator on the device: f_j~U(.05,.5), X_ij~Binom(2,f_j), 1% causal, h2=.5.
    Returns Xt (p x n int8, i.e. the column-major n x p matrix) and y (n float64, host)."""
    import torch
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    f = torch.empty(p, device=dev).uniform_(0.05, 0.5, generator=g)
    Xt = torch.empty((p, n), dtype=torch.int8, device=dev)
    nc = max(1, p // 100)
    causal = torch.randperm(p, generator=g, device=dev)[:nc]
    beta = torch.zeros(p, device=dev)
    beta[causal] = torch.randn(nc, generator=g, device=dev)
    gv = torch.zeros(n, device=dev)
    step = 2048
    for j0 in range(0, p, step):
        j1 = min(p, j0 + step)
        fj = f[j0:j1, None]
        blk = (torch.rand((j1 - j0, n), device=dev, generator=g) < fj).to(torch.int8)
        blk += (torch.rand((j1 - j0, n), device=dev, generator=g) < fj).to(torch.int8)
        Xt[j0:j1] = blk
        gv += beta[j0:j1] @ blk.float()
    gv = (gv - gv.mean()) / (gv.std() + 1e-12) * (0.5 ** 0.5)
    y = gv + torch.randn(n, generator=g, device=dev) * (0.5 ** 0.5)
    return Xt, y.double().cpu().numpy()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            parts = [x.strip() for x in r.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def ncu_traffic(kernel, n, p):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel`, from the committed ncu capture of the same
    workload shape (profiles/r2_traffic.json, written by tools/ncu_traffic.py); None if there is no capture for it."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
        e = d.get(kernel)
        if e and e.get("n") == n:
            return e["bytes_per_launch"] * (p / e["p"])  # captured on a p-slice of the same n; traffic is linear in p
    except Exception:
        pass
    return None


def cpu_sample(args, Xs, y, native=False, sweeps=None):
    """Oracle emRR on full n x the first m markers: sweeps timed as (it sweeps) - (0 sweeps)."""
    import oracle as O
    Xf = np.asfortranarray(Xs, dtype=np.float32)
    t0 = time.perf_counter()
    O.em(args.model, y, Xf, it=0, native=native)
    t1 = time.perf_counter()
    sweeps = sweeps or args.cpu_sweeps
    O.em(args.model, y, Xf, it=sweeps, native=native)
    t2 = time.perf_counter()
    sweeps_s = max((t2 - t1) - (t1 - t0), 1e-9)
    return sweeps * Xf.shape[1] / sweeps_s, sweeps_s


def cpu_model_name():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path (oracle port; 1 thread, as the reference has no
    threading: no src/Makevars, no OpenMP pragma) on a bounded sample of the same workload."""
    if rank != 0:
        return
    rng = np.random.default_rng(SEED)
    m = args.cpu_markers
    f = rng.uniform(0.05, 0.5, size=m)
    Xs = np.empty((args.n, m), dtype=np.int8, order="F")
    for j in range(m):
        Xs[:, j] = rng.binomial(2, f[j], size=args.n)
    beta = np.zeros(m); beta[rng.choice(m, max(1, m // 100), replace=False)] = rng.normal(size=max(1, m // 100))
    gv = Xs @ beta
    y = (gv - gv.mean()) / (gv.std() + 1e-12) * 0.5 ** 0.5 + rng.normal(size=args.n) * 0.5 ** 0.5
    vals = []
    for _ in range(max(1, args.warmup > 0)):
        cpu_sample(args, Xs, y)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        v, _s = cpu_sample(args, Xs, y)
        vals.append(v)
    wall = time.perf_counter() - t0
    val = float(np.median(vals))
    sample = "oracle %s, full n=%d rows x first %d markers, %d sweeps per step (setup subtracted), g++ -O2, %s" % (
        args.model, args.n, m, args.cpu_sweeps, cpu_model_name())
    line = {"impl": "reference", "metric": "marker-updates/sec (emRR Gauss-Seidel sweep, n=50k x p=50k int8)",
            "value": val, "unit": "marker-updates/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * wall / max(1, args.steps), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            # the same workload string and seed as the B200 arm's line (the driver compares the two configs)
            "config": {"workload": "%s Gauss-Seidel sweep, synthetic n=%d x p=%d int8 genotypes, k=1" % (args.model, args.n, args.p),
                       "step": "one full sweep = p marker updates (the reference's per-marker loop, src/Rcpp20260726ai.cpp:308-354)",
                       "parallelism": "1 host thread", "seed": SEED,
                       "note": "reference arm: the CPU loop (1 host thread; the reference's Gauss-Seidel loop is sequential and single-threaded) "
                               "on a bounded marker sample at full n -- per-marker cost is independent of p"},
            "cpu_baseline": {"value": val, "unit": "marker-updates/s", "cores": 1, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "marker-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def mem_available_bytes():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable"):
                return int(line.split()[1]) * 1024
    except Exception:
        pass
    return 0


def host_f64_matrix(Xt_cpu, n, p):
    """R's view of the genotypes: an n x p column-major float64 matrix (what _bWGR_emRR receives, src/RcppExports.cpp:110-121)."""
    Xd = np.empty((n, p), dtype=np.float64, order="F")
    X8 = Xt_cpu.numpy()  # (p, n) int8, row-major = column-major n x p
    step = 2048
    for j0 in range(0, p, step):
        Xd[:, j0:j0 + step] = X8[j0:j0 + step].T
    return Xd


def timed_fits(fn, nfits):
    fn()  # one warm-up fit: allocator, pinned staging, first-touch of the host pages
    out = []
    for _ in range(nfits):
        t = time.perf_counter()
        fn()
        out.append(time.perf_counter() - t)
    return out


def row_sharded_parity(bw, dist, dev, local, rank, world, stream):
    """The sharded fit against the single-GPU fit of the same global data (rank 0 gathers the shards): max|db| / max|b|.
    Small enough for one GPU (world x 8,192 rows x 4,096 markers), the same kernels and exchange as the timed fit."""
    import torch
    n_s, p_s = 8192, 4096
    Xs, ys = synth_gpu(n_s, p_s, SEED + 1000 + rank, dev)
    torch.cuda.synchronize()
    g = bw.Genotypes(device=local, path=bw.PATH_BLOCKED)
    g.enable_row_sharding()
    g.load(Xs)
    fit = bw.emRR(ys, g, it=8)
    g.close()
    Xall = [torch.empty_like(Xs) for _ in range(world)]
    dist.all_gather(Xall, Xs)
    yt = torch.from_numpy(ys).to(dev)
    yall = [torch.empty_like(yt) for _ in range(world)]
    dist.all_gather(yall, yt)
    res = None
    if rank == 0:
        Xg = torch.cat(Xall, dim=1).contiguous()  # (p, world * n_s): the global column-major matrix
        yg = torch.cat(yall).cpu().numpy()
        torch.cuda.synchronize()
        g1 = bw.Genotypes(device=local, path=bw.PATH_BLOCKED)
        g1.load(Xg)
        one = bw.emRR(yg, g1, it=8)
        g1.close()
        res = {"max_abs_db_over_max_abs_b": float(np.abs(fit["b"] - one["b"]).max() / np.abs(one["b"]).max()),
               "h2_sharded": float(fit["h2"]), "h2_single_gpu": float(one["h2"]),
               "shape": "n=%d (=%d x %d) x p=%d, emRR, 8 sweeps" % (n_s * world, world, n_s, p_s)}
    dist.barrier()
    return res


def run_config(args, rank, world, local):
    """One JSON line for BASELINE.json configs 2-4 (`--config K`): the chain / multi-trait / batched-fold callers of the marker loop
    at their named shapes.  A step is one pass over all p markers (one Gibbs iteration, one MRR3 sweep, one sweep of every fit);
    K steps are timed as call(W + K) - call(W) with CUDA events on the library's stream, so set-up and read-back cancel."""
    import torch
    import torch.distributed as dist

    import bwgr_b200 as bw
    import oracle as O
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg, K, W = args.config, args.steps, max(3, args.warmup)
    n, p = (50000, 50000) if cfg == 3 else (10000, 50000)
    Xt, y = synth_gpu(n, p, SEED, dev)
    stream = torch.cuda.Stream(device=dev)
    g = bw.Genotypes(device=local)
    g.set_stream(stream.cuda_stream)
    g.load(Xt)
    hbm_peak, peak_src = peaks()
    rng = np.random.default_rng(2)
    ktr = 20
    Y = np.stack([y * np.sqrt(0.5) + rng.normal(size=n) * np.sqrt(0.5) for _ in range(ktr)], axis=1)
    units = 1  # independent fits sharing one pass over the genotypes
    if cfg == 2:
        name = "wgr BayesB (pi=0.95, iv=TRUE) Gibbs, synthetic n=10000 x p=50000 int8 genotypes; step = one MCMC iteration over all markers"
        call = lambda it: bw.wgr(y, g, it=it, bi=max(1, it // 4), pi=0.95, iv=True, seed=1)  # noqa: E731
        kernel = "sweep_pipe_kernel (Kuo-Mallick rule) + wgr step"
    elif cfg == 3:
        name = "MRR3 multivariate ridge, synthetic n=50000 x p=50000 int8 genotypes, k=20 traits; step = one sweep (p marker updates of 20 traits each)"
        call = lambda it: bw.MRR3(Y, g, maxit=it, tol=0.0)  # noqa: E731
        kernel = "sweep_pipe_kernel (20 rotated systems)"
        if args.missing > 0:
            Y = Y.copy()
            Y[rng.random(Y.shape) < args.missing] = np.nan
            name += "; %.0f %% of the phenotypes missing at random (per-trait masked k x k systems, float64)" % (100 * args.missing)
            kernel = "mrr_gen_sweep_kernel (one k x k system per marker, per-marker grid sum)"
    else:
        folds = 5
        perm = rng.permutation(n)
        mine = [i for i in range(folds * ktr) if i % world == rank]  # fits sharded round-robin, no communication
        Yall = np.repeat(Y, folds, axis=1)[:, mine]
        mask = np.ones((n, folds * ktr), dtype=bool)
        for t in range(ktr):
            for f in range(folds):
                mask[perm[f * n // folds:(f + 1) * n // folds], folds * t + f] = False
        mask = np.ascontiguousarray(mask[:, mine])
        units = folds * ktr
        name = ("emCV pattern: 5 folds x 20 traits = 100 emBC fits on synthetic n=10000 x p=50000 int8 genotypes (row masks), %d fits per GPU; "
                "step = one sweep of every fit" % len(mine))
        call = lambda it: bw.em_fit("emBC", Yall, g, it=it, row_mask=mask)  # noqa: E731
        kernel = "masked batched sweep (sweep_pipe_kernel / small_n_kernel per plan)"

    def timed_call(it):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = g.launch_count()
        e0.record(stream)
        r = call(it)
        e1.record(stream)
        torch.cuda.synchronize()
        return r, e0.elapsed_time(e1), g.launch_count() - l0

    sampler = ClockSampler(local)
    sampler.start()
    call(W)  # warm-up (>= 3 steps)
    _, t_w, l_w = timed_call(W)
    res, t_wk, l_wk = timed_call(W + K)
    ms = t_wk - t_w
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    t_clk = time.perf_counter()
    while time.perf_counter() - t_clk < 1.5 or (len(sampler.rows) < 10 and time.perf_counter() - t_clk < 6.0):
        call(W)
    clocks = sampler.stop()
    value = units * p * K / (ms * 1e-3)
    fits_here = units // world if cfg == 4 else 1
    achieved = n * p / (ms / K * 1e-3) / 1e9  # the genotypes are streamed once per step per GPU whatever the number of fits on it
    roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": None,
                "kernel": kernel, "peak_source": peak_src, "algorithmic_bytes_per_launch": n * p,
                "note": "whole step (all launches of one pass) against one pass over the int8 genotypes; these chains are latency-bound, not HBM-bound"}
    e2e = None
    if not args.no_e2e and rank == 0:
        Xh = Xt.cpu().numpy().T  # n x p int8, host (column-major view)
        it_e = W + K
        t0 = time.perf_counter()
        if cfg == 2:
            bw.wgr(y, Xh, it=it_e, bi=max(1, it_e // 4), pi=0.95, iv=True, seed=1)
        elif cfg == 3:
            bw.MRR3(Y, Xh, maxit=it_e, tol=0.0)
        else:
            bw.em_fit("emBC", Yall, Xh, it=it_e, row_mask=mask)
        dt = time.perf_counter() - t0
        e2e = {"value": units // (world if cfg == 4 else 1) * p * it_e / dt, "unit": "marker-updates/s", "h2d_bytes_per_step": int(n * p), "d2h_bytes_per_step": int(8 * (n + p) * (ktr if cfg != 2 else 1)),
               "step": "one public call on a HOST int8 matrix with %d steps: H2D, column statistics, the steps, read-back; rank 0 only" % it_e, "seconds": dt}
    cpu = None
    if rank == 0 and not args.no_cpu:
        ns_, m = (10000, 2048) if cfg == 3 else (n, 1024)
        Xs = Xt[:m, :ns_].cpu().numpy().T.astype(np.float64)
        t0 = time.perf_counter()
        if cfg == 2:
            its = 400
            O.wgr(y[:ns_], Xs, it=its, bi=10, pi=0.95, iv=True, seed=1)
            upd = its * m
        elif cfg == 3:
            its = 12
            O.mrr3(Y[:ns_], Xs, maxit=its, tol=0.0)
            upd = its * m
        else:
            its = 60
            for f in range(4):
                O.em("emBC", Yall[:, f], Xs.astype(np.float32), it=its)
            upd = 4 * its * m
        secs = time.perf_counter() - t0
        cpu = {"value": upd / secs * (ns_ / n), "unit": "marker-updates/s", "cores": 1, "kind": "port",
               "sample": "oracle on %d rows x first %d markers, %d steps = %.1f s, scaled by rows to n=%d (cost is linear in n); %s" % (
                   ns_, m, its, secs, n, cpu_model_name())}
    if rank == 0:
        line = {"metric": "marker-updates/sec (BASELINE config %d)" % cfg, "value": value, "unit": "marker-updates/s", "n_gpus": world, "steps": K,
                "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong" if cfg == 4 else "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": {"workload": name, "seed": SEED, "fits_on_this_gpu": fits_here,
                                                "l2": "genotypes are %.1f GB per step, larger than the 126 MB L2" % (n * p / 1e9)},
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(l_wk), "clocks": clocks}
        if cfg == 2:
            line["check"] = {"mean_inclusion": float(np.mean(res["d"])), "Ve": res["Ve"], "cor_hat_y": float(np.corrcoef(res["hat"], y)[0, 1])}
        elif cfg == 3:
            line["check"] = {"h2_mean": float(np.mean(res["h2"])), "marker_trait_updates_per_s": value * ktr}
        else:
            line["check"] = {"h2_mean": float(np.mean(res["h2"]))}
        print(json.dumps(line), flush=True)
    g.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if args.config == 5:  # one GPU's share of the 500k x 100k row-sharded fit: the main flow below at that shard shape
        args.n, args.p = 62500, 100000
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.config in (2, 3, 4):
        run_config(args, rank, world, local)
        return
    import torch
    import torch.distributed as dist

    import bwgr_b200 as bw
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n, p = args.n, args.p
    Xt, y = synth_gpu(n, p, SEED + rank, dev)
    torch.cuda.synchronize()
    stream = torch.cuda.Stream(device=dev)  # events and the library share this stream
    hbm_peak, peak_src = peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_sweeps(g, st):
        """W warm-up sweeps, then exactly K sweeps between a barrier + synchronize on both sides; max over ranks."""
        st.sweeps(args.warmup)
        barrier()
        l0 = g.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        st.sweeps(args.steps)
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        g.profile(True)  # per-kernel durations for the roofline (separate pass: event pairs around every kernel)
        st.sweeps(args.steps)
        prof = g.profile_read()
        g.profile(False)
        return ms, g.launch_count() - l0 - 0, prof

    # ---- independent fits: one emRR sweep stream per GPU (N = 1: THE workload; N > 1: the communication-free mode, extra key)
    g = bw.Genotypes(device=local, path=bw.PATH_BLOCKED)
    g.set_stream(stream.cuda_stream)
    g.load(Xt)  # device-resident int8 (p x n row-major == n x p column-major)
    st = bw.EmStepper(args.model, y, g)
    sampler = ClockSampler(local)
    sampler.start()  # started before the warm-up: the timed region can be shorter than nvidia-smi's first sample
    ms_rep, launches_rep, prof_rep = timed_sweeps(g, st)
    launches_rep //= 2  # the per-kernel timing pass repeats the K sweeps
    rep_value = world * p * args.steps / (ms_rep * 1e-3)

    # ---- N > 1: ONE fit, individuals sharded by rows over the ranks (BASELINE config 5 pattern) -- the headline of --gpus N
    ms, launches, prof, rs_parity, rs_fit = ms_rep, launches_rep, prof_rep, None, None
    if world > 1:
        g3 = bw.Genotypes(device=local, path=bw.PATH_BLOCKED)
        g3.enable_row_sharding()
        g3.set_stream(stream.cuda_stream)
        g3.load(Xt)
        st3 = bw.EmStepper(args.model, y, g3)
        ms, launches, prof = timed_sweeps(g3, st3)
        launches //= 2
    # nvidia-smi needs ~1 s to deliver its first sample and the timed region is ~50 ms: keep the SAME sweep loop running until
    # the sampler has seen the GPU under this load for a while, then read the clocks / throttle reasons
    t_clk = time.perf_counter()
    stc = st3 if world > 1 else st
    while time.perf_counter() - t_clk < 1.5 or (len(sampler.rows) < 10 and time.perf_counter() - t_clk < 6.0):
        stc.sweeps(20)
        torch.cuda.synchronize()
    clocks = sampler.stop()  # covers warm-up, the timed region, the per-kernel timing pass and the clock pass (all the same sweeps)
    fit = st.end()
    assert np.isfinite(fit["b"]).all() and np.isfinite(fit["h2"])
    g.close()
    if world > 1:
        rs_fit = st3.end()
        g3.close()
        assert np.isfinite(rs_fit["b"]).all()
        try:
            rs_parity = row_sharded_parity(bw, dist, dev, local, rank, world, stream)
        except Exception as ex:  # noqa: BLE001
            rs_parity = {"error": str(ex)[:200]}

    sweep_ms = prof["sweep"]["ms"] / max(1, prof["sweep"]["launches"])
    gram_ms = prof["gram"]["ms"] / max(1, prof["gram"]["launches"])
    epi_ms = prof["epilogue"]["ms"] / max(1, prof["epilogue"]["launches"])
    inv_ms = prof["block_inverse"]["ms"] / max(1, prof["block_inverse"]["launches"])
    dom = "sweep_pipe_kernel" if sweep_ms >= gram_ms else "gram_fp4_kernel"
    dom_ms = max(sweep_ms, gram_ms)
    achieved = n * p / (dom_ms * 1e-3) / 1e9  # per GPU: every rank streams its own n x p bytes per launch
    roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "traffic": args.traffic if args.traffic is not None else ncu_traffic(dom, n, p), "kernel": dom, "peak_source": peak_src, "algorithmic_bytes_per_launch": n * p,
                "kernel_ms": {"sweep_pipe_kernel": sweep_ms, "gram_fp4_kernel": gram_ms, "block_inverse_kernel": inv_ms, "epilogue_kernel": epi_ms},
                "whole_sweep_frac": (n * p / (ms / args.steps * 1e-3) / 1e9) / hbm_peak}
    # one marker update = one marker visited on one 50k-row shard: N = 1 -> p per sweep; row-sharded -> N * p per sweep of the ONE fit
    value = world * p * args.steps / (ms * 1e-3)

    # ---- end to end through the public API with HOST buffers
    e2e = None
    if not args.no_e2e:
        Xh = Xt.cpu()
        del Xt
        torch.cuda.empty_cache()
        torch.cuda.synchronize()
        if world == 1:
            need = 8 * n * p
            use_f64 = mem_available_bytes() > need + (12 << 30)
            fits8 = None
            if use_f64:
                Xd = host_f64_matrix(Xh, n, p)

                def one_fit():
                    out = bw.emRR(y, Xd, it=args.e2e_sweeps, path=bw.PATH_BLOCKED)  # numpy double matrix in, R-style list out
                    assert np.isfinite(out["b"]).all()
                fit_s = timed_fits(one_fit, max(5, args.e2e_fits))
                del Xd
            Xp = Xh.pin_memory()

            def one_fit8():
                g2 = bw.Genotypes(device=local, path=bw.PATH_BLOCKED)
                g2.load(Xp)
                out = bw.emRR(y, g2, it=args.e2e_sweeps)
                g2.close()
                assert np.isfinite(out["b"]).all()
            fits8 = timed_fits(one_fit8, max(5, args.e2e_fits))
            if not use_f64:
                fit_s = fits8
            dt = float(np.median(fit_s))
            e2e = {"value": args.e2e_sweeps * p / dt, "unit": "marker-updates/s",
                   "h2d_bytes_per_step": int(n * p + 8 * n), "d2h_bytes_per_step": int(4 * (p + n) + 64),
                   "host_input_bytes_per_step": int((8 if use_f64 else 1) * n * p + 8 * n),
                   "step": ("one emRR(y, gen) call on R's double matrix (n x p float64 on the host): exact narrowing to int8 by host threads into pinned "
                            "staging overlapped with the H2D copy, column statistics, %d sweeps, GEBVs, D2H" % args.e2e_sweeps) if use_f64 else
                           ("one emRR(y, gen) call on a host int8 matrix (not enough host memory for the float64 copy): H2D, column statistics, %d sweeps, GEBVs, D2H" % args.e2e_sweeps),
                   "seconds_per_fit": dt, "seconds_each_fit": fit_s, "spread": (max(fit_s) - min(fit_s)) / dt,
                   "iqr_over_median": float(np.subtract(*np.percentile(fit_s, [75, 25])) / dt),
                   "int8_host_input": {"value": args.e2e_sweeps * p / float(np.median(fits8)), "seconds_each_fit": fits8}}
        else:
            Xp = Xh.pin_memory()

            # the store handle (NCCL communicator + the peer-mapped exchange ring: seconds to set up) lives for the session, like the
            # process group itself; each timed call uploads the genotypes again and fits
            g2 = bw.Genotypes(device=local, path=bw.PATH_BLOCKED)
            g2.enable_row_sharding()

            def one_fit_sharded():
                g2.load(Xp)  # this rank's rows, host int8
                out = bw.emRR(y, g2, it=args.e2e_sweeps)
                assert np.isfinite(out["b"]).all()
            fit_s = timed_fits(one_fit_sharded, max(3, args.e2e_fits))
            g2.close()
            dt = float(np.median(fit_s))
            t = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
            e2e = {"value": world * args.e2e_sweeps * p / dt, "unit": "marker-updates/s",
                   "h2d_bytes_per_step": int(world * (n * p + 8 * n)), "d2h_bytes_per_step": int(world * 4 * (p + n) + 64),
                   "step": "one row-sharded emRR(y, gen) call over %d GPUs on a store handle that outlives the call (communicator and exchange ring are set up "
                           "once per session): every rank loads ITS rows from a host int8 matrix (H2D), column statistics (all-reduced), %d sweeps, GEBVs, D2H" % (world, args.e2e_sweeps),
                   "seconds_per_fit": dt, "seconds_each_fit": fit_s}

    cpu = None
    if rank == 0 and not args.no_cpu:
        m = args.cpu_markers
        Xs = (Xt[:m].cpu() if args.no_e2e else Xh[:m]).numpy().T
        v, secs = cpu_sample(args, Xs, y, sweeps=args.cpu_sweeps_main)
        cpu = {"value": v, "unit": "marker-updates/s", "cores": 1, "kind": "port",
               "sample": "oracle %s (g++ -O2, float32, 1 thread like the reference; pinned against the reference's own sources, tests/test_ref_pin.py), "
                         "full n=%d x first %d markers, %d sweeps = %.1f s; %s, %d host cores" % (
                             args.model, n, m, args.cpu_sweeps_main, secs, cpu_model_name(), os.cpu_count())}

    if rank == 0:
        workload = "%s Gauss-Seidel sweep, synthetic n=%d x p=%d int8 genotypes, k=1" % (args.model, n, p)
        if world == 1:
            par = "1 GPU"
        else:
            par = ("ONE fit of n=%d individuals sharded by rows over %d GPUs (%d rows each); per 128-marker block the reduced partials cross NVLink as "
                   "peer stores inside the sweep kernel, per sweep ncclAllReduce of the Gram band and 5 scalars; value counts marker updates per "
                   "%d-row shard (N x p per sweep)" % (n * world, world, n, n))
        metric = "marker-updates/sec (emRR Gauss-Seidel sweep, n=50k x p=50k int8)" if args.config == 1 else \
            "marker-updates/sec (BASELINE config 5: emRR Gauss-Seidel, row shards of %d x %d int8 per GPU)" % (n, p)
        line = {"metric": metric, "value": value,
                "unit": "marker-updates/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload,
                           "step": "one full sweep = p marker updates (Gram band + pipelined blocked sweep + epilogue)",
                           "l2": "genotypes are %.1f GB per sweep per GPU, far larger than the 126 MB L2: no flush needed" % (n * p / 1e9),
                           "parallelism": par, "seed": SEED},
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
                "sweep_ms": ms / args.steps}
        if world > 1:
            line["row_sharded_parity"] = rs_parity
            line["row_sharded"] = {"ms_per_sweep": ms / args.steps, "genotype_GB_per_s_aggregate": n * world * p / (ms / args.steps * 1e-3) / 1e9,
                                   "frac_of_aggregate_hbm_peak": n * world * p / (ms / args.steps * 1e-3) / 1e9 / (hbm_peak * world), "h2": float(rs_fit["h2"])}
            line["replicas"] = {"what": "%d independent fits, one per GPU, no collective" % world, "value": rep_value, "ms_per_sweep": ms_rep / args.steps}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
