#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-x}
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/tests_${TAG}.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/tests_${TAG}.log
for c in 2 4 3; do
  timeout 900 python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/bench_${TAG}_cfg$c.json 2> gpurun_out/bench_${TAG}_cfg$c.err; echo "cfg $c rc=$?"
  python - <<P
import json
d=json.loads(open("gpurun_out/bench_${TAG}_cfg$c.json").read().strip().splitlines()[-1])
print($c, "ms/step", round(d["ms_per_step"],3), "value %.4g"%d["value"], "cpu %.4g"%d["cpu_baseline"]["value"], d["cpu_baseline"]["sample"][:80], "e2e %.4g"%d["e2e"]["value"], d.get("check"))
P
done
BWGR_TRACE=gpurun_out/trace_wgr_${TAG}.bin timeout 300 python bench.py --config 2 --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/wgr_trace_${TAG}.json 2> gpurun_out/wgr_trace_${TAG}.err; echo "wgr trace rc=$?"
python tools/trace_cl.py gpurun_out/trace_wgr_${TAG}.bin 2>&1 | sed -n 2,4p
