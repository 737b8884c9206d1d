#!/bin/bash
# Round-end evidence in one GPU call: the new tests first (fast), the whole GPU suite, then tools/gpu_profile.sh (bench, reference arm,
# launch list, one ncu --set full capture).  usage: tools/gpu_final.sh TAG
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r2b}
timeout 600 python -m pytest tests -m gpu -q --durations=8 -k "kmup2 or bagged" > gpurun_out/tests_new_${TAG}.log 2>&1; echo "new tests rc=$?"; tail -15 gpurun_out/tests_new_${TAG}.log
timeout 1500 python -m pytest tests -q -m gpu --durations=10 > gpurun_out/tests_${TAG}.log 2>&1; echo "tests rc=$?"; tail -16 gpurun_out/tests_${TAG}.log
tools/gpu_profile.sh ${TAG}
