#!/bin/bash
# One GPU iteration on the clustered topology: blocked-path parity tests, then short benches with traces.
# usage: tools/gpu_cl.sh TAG "ENV1=.. ENV2=.." ...   (one bench per extra argument; default: clustered fast, clustered, flat)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-x}; shift
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "2-em or path2 or 2] or deterministic or bench_geometry or synthetic_mid or blocked or ragged or config2" > gpurun_out/t_${TAG}.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/t_${TAG}.log
if [ $# -eq 0 ]; then set -- "BWGR_CLUSTER=1" "BWGR_FASTW=0" "BWGR_CLUSTER=0"; fi
i=0
for CFG in "$@"; do
i=$((i+1))
env $CFG BWGR_TRACE=gpurun_out/trace_${TAG}_$i.bin timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_${TAG}_$i.log 2> gpurun_out/bench_${TAG}_$i.err
echo "bench [$CFG] rc=$?"; python - <<P
import json
try:
    d=json.loads(open("gpurun_out/bench_${TAG}_$i.log").read().strip().splitlines()[-1])
    print(d["ms_per_step"], d["value"], d["roofline"]["kernel_ms"], d["roofline"]["frac"])
except Exception as e:
    print("no json", e); print(open("gpurun_out/bench_${TAG}_$i.err").read()[-1500:])
P
done
