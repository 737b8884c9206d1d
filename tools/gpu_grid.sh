#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-x}
run() { env "$@" timeout 300 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu 2>gpurun_out/grid_${TAG}.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$*', '->', round(d['ms_per_step'],4), {k: round(v,4) for k,v in d['roofline']['kernel_ms'].items()})"; }
run BWGR_X=1
run BWGR_X=2
for G in 100 104 108; do run BWGR_CLUSTER=0 BWGR_GRID=$G; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"epilogue_kernel|block_inverse" -s 6 -c 2 -o gpurun_out/prof_small_${TAG} python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_small_${TAG}.log 2>&1; echo "ncu rc=$?"
