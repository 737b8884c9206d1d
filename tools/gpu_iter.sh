#!/bin/bash
# One GPU iteration: blocked-path parity tests (both look-ahead depths), then a short bench with a clock trace.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-x}
BWGR_LOOKAHEAD=0 timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "2-em or path2 or 2] or deterministic or gram" > gpurun_out/t_${TAG}_d0.log 2>&1
echo "D0 tests rc=$?"; tail -3 gpurun_out/t_${TAG}_d0.log
BWGR_LOOKAHEAD=1 timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "2-em or path2 or 2] or deterministic" > gpurun_out/t_${TAG}_d1.log 2>&1
echo "D1 tests rc=$?"; tail -3 gpurun_out/t_${TAG}_d1.log
for D in 0 1; do
BWGR_LOOKAHEAD=$D BWGR_TRACE=gpurun_out/trace_${TAG}_d$D.txt timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_${TAG}_d$D.log 2> gpurun_out/bench_${TAG}_d$D.err
echo "bench D=$D rc=$?"; python - <<P
import json
try:
    d=json.loads(open("gpurun_out/bench_${TAG}_d$D.log").read().strip().splitlines()[-1])
    print(d["ms_per_step"], d["value"], d["roofline"]["kernel_ms"], d["roofline"]["frac"])
except Exception as e:
    print("no json", e); print(open("gpurun_out/bench_${TAG}_d$D.err").read()[-1500:])
P
done
