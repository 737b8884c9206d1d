"""BWGR_GAPS=1: where the time of a sweep goes on the main stream (kernel boundaries), emRR 50k x 50k."""
import os, sys
os.environ["BWGR_GAPS"] = "1"
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, bwgr_b200 as bw
dev = torch.device("cuda", 0)
Xt, y = bench.synth_gpu(50000, 50000, bench.SEED, dev)
g = bw.Genotypes(device=0, path=bw.PATH_BLOCKED)
g.load(Xt)
st = bw.EmStepper("emRR", y, g)
st.sweeps(80)
torch.cuda.synchronize()
out = st.end()
print("h2", out["h2"])
