#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-x}
run() { env "$@" timeout 300 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu 2>gpurun_out/maxcl_${TAG}.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$*', '->', round(d['ms_per_step'],4), {k: round(v,4) for k,v in d['roofline']['kernel_ms'].items()})"; }
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "variant or deterministic or bench_geometry or golden or synthetic_mid or ahead" 2>&1 | tail -2
run BWGR_X=1
run BWGR_MAXCL=15
BWGR_TRACE=gpurun_out/trace_wgr_${TAG}.bin timeout 300 python bench.py --config 2 --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/wgr_trace_${TAG}.json 2> gpurun_out/wgr_trace_${TAG}.err; echo "wgr trace rc=$?"
python tools/trace_cl.py gpurun_out/trace_wgr_${TAG}.bin 2>&1 | tail -16
