#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-x}
BWGR_TRACE=gpurun_out/trace_mrr_${TAG}.bin timeout 600 python bench.py --config 3 --steps 4 --warmup 3 --no-e2e --no-cpu > gpurun_out/mrr_trace_${TAG}.json 2> gpurun_out/mrr_trace_${TAG}.err; echo "mrr trace rc=$?"; tail -2 gpurun_out/mrr_trace_${TAG}.err
python tools/trace_pipe.py gpurun_out/trace_mrr_${TAG}.bin 2>&1 | tail -24
