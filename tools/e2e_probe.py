"""Where the time of one end-to-end emRR(y, gen) call goes at 50k x 50k (handle, H2D + pack of pinned int8 genotypes, 200 sweeps, outputs)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, bwgr_b200 as bw
dev = torch.device("cuda", 0)
Xt, y = bench.synth_gpu(50000, 50000, bench.SEED, dev)
Xh = torch.empty((50000, 50000), dtype=torch.int8, pin_memory=True); Xh.copy_(Xt); del Xt
torch.cuda.synchronize()
for rep in range(2):
    t0 = time.perf_counter()
    g = bw.Genotypes(device=0, path=bw.PATH_BLOCKED); t1 = time.perf_counter()
    g.load(Xh); torch.cuda.synchronize(); t2 = time.perf_counter()
    st = bw.EmStepper("emRR", y, g); torch.cuda.synchronize(); t3 = time.perf_counter()
    st.sweeps(200); torch.cuda.synchronize(); t4 = time.perf_counter()
    out = st.end(); t5 = time.perf_counter()
    g.close(); t6 = time.perf_counter()
    print("create %.3f load %.3f begin %.3f sweeps %.3f end %.3f close %.3f total %.3f" % (t1-t0, t2-t1, t3-t2, t4-t3, t5-t4, t6-t5, t6-t0), flush=True)
