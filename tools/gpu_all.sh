#!/bin/bash
# One GPU call, several stages, each bounded and logged.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-x}
K="${2:-gram or bench_geometry or deterministic or synthetic_mid}"
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "$K" > gpurun_out/tA_${TAG}.log 2>&1
echo "A (subset) rc=$?"; tail -4 gpurun_out/tA_${TAG}.log
i=0
for CFG in "BWGR_X=1"; do
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
python tools/trace_cl.py gpurun_out/trace_${TAG}_1.bin
timeout 300 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_${TAG}_notrace.log 2>&1; python - <<P
import json
d=json.loads(open("gpurun_out/bench_${TAG}_notrace.log").read().strip().splitlines()[-1])
print("no trace:", d["ms_per_step"], d["roofline"]["kernel_ms"])
P
