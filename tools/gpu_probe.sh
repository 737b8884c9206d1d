#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-x}
timeout 900 python tools/load_probe.py > gpurun_out/load_probe_${TAG}.log 2>&1; echo "probe rc=$?"; cat gpurun_out/load_probe_${TAG}.log
for c in 2 3 4; do
  timeout 900 python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/bench_${TAG}_cfg$c.json 2> gpurun_out/bench_${TAG}_cfg$c.err; echo "cfg $c rc=$?"
  tail -c 1800 gpurun_out/bench_${TAG}_cfg$c.json; tail -3 gpurun_out/bench_${TAG}_cfg$c.err
done
