#!/bin/bash
# Evidence run: full default bench, reference arm, launch list, one ncu --set full capture of the sweep and Gram kernels.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r2}
timeout 1200 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/bench_${TAG}.json; tail -5 gpurun_out/bench_${TAG}.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_${TAG}_ref.json 2>> gpurun_out/bench_${TAG}.err; echo "ref rc=$?"
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --p 19200"
timeout 300 $CMD > gpurun_out/plain_${TAG}.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_list_${TAG}.log 2>&1
echo "launch list rc=$?"
timeout 300 $CMD > gpurun_out/plain2_${TAG}.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"sweep_pipe|gram_fp4" -s 4 -c 2 -o gpurun_out/prof_${TAG} $CMD > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full_${TAG}.log
timeout 600 python tools/cfg4_probe.py > gpurun_out/cfg4_probe_${TAG}.log 2>&1; echo "cfg4 probe rc=$?"; tail -3 gpurun_out/cfg4_probe_${TAG}.log
