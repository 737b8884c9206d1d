#!/bin/bash
# One GPU call, several stages, each bounded and logged (the pod queue is long: batch).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-x}
K="2-em or path2 or 2] or deterministic or bench_geometry or synthetic_mid or blocked or ragged or config2 or gs_warm or emml or centred or mrr3"
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "$K" > gpurun_out/tA_${TAG}.log 2>&1
echo "A (subset) rc=$?"; tail -4 gpurun_out/tA_${TAG}.log
i=0
for CFG in "BWGR_X=1" "BWGR_CLUSTER=0"; do
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
