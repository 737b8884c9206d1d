#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-x}
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/tests_${TAG}.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/tests_${TAG}.log
timeout 600 python tools/load_probe.py > gpurun_out/load_probe_${TAG}.log 2>&1; echo "probe rc=$?"; tail -9 gpurun_out/load_probe_${TAG}.log
timeout 900 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?"; tail -c 2500 gpurun_out/bench_${TAG}.json
