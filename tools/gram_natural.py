"""Gram band of the natural marker order with and without the packed 2-bit source (BWGR_GRAM_PACKED), through a BayesRR fit that computes it once."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, bwgr_b200 as bw
dev = torch.device("cuda", 0)
Xt, y = bench.synth_gpu(50000, 50000, bench.SEED, dev)
for packed in ("1", "0"):
    os.environ["BWGR_GRAM_PACKED"] = packed
    g = bw.Genotypes(device=0, path=bw.PATH_BLOCKED)
    g.load(Xt)
    g.profile(True)
    r = bw.BayesRR(y, g, it=3, bi=1, seed=1)   # natural order: the Gram band is computed once
    pr = g.profile_read()
    print("natural order, packed", packed, {k: (v["ms"], v["launches"]) for k, v in pr.items()}, flush=True)
    st = bw.EmStepper("emRR", y, g)
    g.profile(True)
    st.sweeps(3)
    pr = g.profile_read()
    print("shuffled order, packed", packed, {k: (v["ms"] / max(1, v["launches"]), v["launches"]) for k, v in pr.items()}, flush=True)
    st.end()
    g.close()
