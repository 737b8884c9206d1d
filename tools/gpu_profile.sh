#!/bin/bash
# Round-end evidence: full default bench, launch list, one ncu --set full capture of the sweep and Gram kernels.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r1}
timeout 900 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench_${TAG}.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_${TAG}_ref.json 2>> gpurun_out/bench_${TAG}.err; echo "ref rc=$?"
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --p 6400"
timeout 300 $CMD > gpurun_out/plain_${TAG}.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_list_${TAG}.log 2>&1
echo "launch list rc=$?"
timeout 300 $CMD > gpurun_out/plain2_${TAG}.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"sweep_pipe|gram_tc" -s 4 -c 2 -o gpurun_out/prof_${TAG} $CMD > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full_${TAG}.log
