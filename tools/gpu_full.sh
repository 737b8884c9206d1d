#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-x}
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/tests_${TAG}.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/tests_${TAG}.log
timeout 900 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?"; tail -c 2500 gpurun_out/bench_${TAG}.json
python - <<P
import json
d=json.loads(open("gpurun_out/bench_${TAG}.json").read().strip().splitlines()[-1])
print("ms/sweep", d["ms_per_step"], d["roofline"]["kernel_ms"], "e2e", d["e2e"]["seconds_per_fit"], d["e2e"]["seconds_each_fit"])
P
