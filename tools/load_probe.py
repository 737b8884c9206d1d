"""Where the host side of emRR(y, gen) on a float64 matrix goes: CPU budget of the box, loader throughput against the number of
host threads, and repeated whole calls."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, bwgr_b200 as bw

for f in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu.stat", "/sys/fs/cgroup/memory.max"):
    try:
        print(f, open(f).read().strip().replace("\n", " | "))
    except Exception as e:
        print(f, "n/a", e)
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)), "MemAvailable GB", bench.mem_available_bytes() / 1e9)
os.system("lscpu | grep -E 'Model name|Socket|NUMA|Thread|Core' ; uptime")
dev = torch.device("cuda", 0)
n, p = 50000, 50000
Xt, y = bench.synth_gpu(n, p, bench.SEED, dev)
Xh = Xt.cpu(); del Xt
torch.cuda.empty_cache()
pq = 12500
Xd = bench.host_f64_matrix(Xh[:pq], n, pq)
for thr in (4, 8, 12, 16, 24):
    os.environ["BWGR_LOAD_THREADS"] = str(thr)
    ts = []
    for r in range(5):
        t0 = time.perf_counter()
        g = bw.Genotypes(Xd, device=0)
        ts.append(time.perf_counter() - t0)
        g.close()
    print("threads %2d: load of %d x %d float64 (%.1f GB): %s  -> best %.1f GB/s" % (thr, n, pq, Xd.nbytes / 1e9, " ".join("%.3f" % t for t in ts), Xd.nbytes / min(ts) / 1e9), flush=True)
del os.environ["BWGR_LOAD_THREADS"]
del Xd
Xd = bench.host_f64_matrix(Xh, n, p)
for r in range(6):
    t0 = time.perf_counter()
    g = bw.Genotypes(Xd, device=0, path=bw.PATH_BLOCKED); t1 = time.perf_counter()
    st = bw.EmStepper("emRR", y, g); torch.cuda.synchronize(); t2 = time.perf_counter()
    st.sweeps(200); torch.cuda.synchronize(); t3 = time.perf_counter()
    out = st.end(); t4 = time.perf_counter()
    g.close(); t5 = time.perf_counter()
    print("full call: store+load %.3f begin %.3f sweeps %.3f end %.3f close %.3f total %.3f" % (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t5 - t0), flush=True)
print(open("/sys/fs/cgroup/cpu.stat").read().strip().replace("\n", " | ") if os.path.exists("/sys/fs/cgroup/cpu.stat") else "")
