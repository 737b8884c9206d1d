"""Gram kernel with its MMAs switched off (BWGR_GRAM_DBG=1) vs on, per look-ahead depth: isolates the gather rate of the producer warps."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, bwgr_b200 as bw
dev = torch.device("cuda", 0)
Xt, y = bench.synth_gpu(50000, 50000, bench.SEED, dev)
for D in (0, 1):
    os.environ["BWGR_LOOKAHEAD"] = str(D)
    g = bw.Genotypes(device=0, path=bw.PATH_BLOCKED)
    g.load(Xt)
    st = bw.EmStepper("emRR", y, g)
    g.profile(True)
    st.sweeps(4)
    torch.cuda.synchronize()
    pr = g.profile_read()
    print("D", D, "dbg", os.environ.get("BWGR_GRAM_DBG"), {k: v["ms"] / max(1, v["launches"]) for k, v in pr.items()}, flush=True)
    try:
        st.end()
    except Exception as e:
        print("end:", str(e)[:80])
    g.close()
