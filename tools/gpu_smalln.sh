#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-x}
CFG4_P=6400 timeout 300 python tools/cfg4_probe.py > gpurun_out/smalln_plain_${TAG}.log 2>&1; tail -2 gpurun_out/smalln_plain_${TAG}.log
CFG4_P=6400 timeout 900 ncu --set full --clock-control none --import-source on -k regex:small_n -s 2 -c 1 -o gpurun_out/prof_smalln_${TAG} python tools/cfg4_probe.py > gpurun_out/ncu_smalln_${TAG}.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_smalln_${TAG}.log
