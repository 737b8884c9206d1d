#!/bin/bash
cd "$(dirname "$0")/.."
TAG=${1:-x}
for D in ${2:-0 1}; do
BWGR_LOOKAHEAD=$D BWGR_TRACE=gpurun_out/trace_${TAG}_d$D.bin timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_${TAG}_d$D.log 2> gpurun_out/bench_${TAG}_d$D.err
echo "bench D=$D rc=$?"; python - <<P
import json
try:
    d=json.loads(open("gpurun_out/bench_${TAG}_d$D.log").read().strip().splitlines()[-1])
    print(d["ms_per_step"], d["value"], d["roofline"]["kernel_ms"], d["roofline"]["frac"])
except Exception as e:
    print("no json", e); print(open("gpurun_out/bench_${TAG}_d$D.err").read()[-1500:])
P
done
