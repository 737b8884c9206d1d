#!/bin/bash
# One GPU iteration on the clustered topology: blocked-path parity tests, then short benches (clustered vs flat) with traces.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-x}
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "2-em or path2 or 2] or deterministic or bench_geometry or synthetic_mid or blocked or ragged or config2" > gpurun_out/t_${TAG}.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/t_${TAG}.log
for CLU in 1 0; do
BWGR_CLUSTER=$CLU BWGR_TRACE=gpurun_out/trace_${TAG}_cl$CLU.bin timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_${TAG}_cl$CLU.log 2> gpurun_out/bench_${TAG}_cl$CLU.err
echo "bench CL=$CLU rc=$?"; python - <<P
import json
try:
    d=json.loads(open("gpurun_out/bench_${TAG}_cl$CLU.log").read().strip().splitlines()[-1])
    print(d["ms_per_step"], d["value"], d["roofline"]["kernel_ms"], d["roofline"]["frac"])
except Exception as e:
    print("no json", e); print(open("gpurun_out/bench_${TAG}_cl$CLU.err").read()[-1500:])
P
done
