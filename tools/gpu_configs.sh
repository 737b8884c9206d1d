#!/bin/bash
# default bench line + the per-config lines, one GPU
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-x}
timeout 900 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench_${TAG}.json
for c in 2 3 4 5; do
  timeout 900 python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/bench_${TAG}_cfg$c.json 2> gpurun_out/bench_${TAG}_cfg$c.err; echo "cfg $c rc=$?"
  tail -c 1800 gpurun_out/bench_${TAG}_cfg$c.json; tail -3 gpurun_out/bench_${TAG}_cfg$c.err
done
