#!/bin/bash
# multi-GPU validation: the 2-GPU tests, then bench.py under torchrun exactly as the driver launches it (ours and the reference arm)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}; TAG=${2:-x}
timeout 900 python -m pytest tests -x -q -m gpu -k "two_gpus or sharded" > gpurun_out/tests_n${N}_${TAG}.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/tests_n${N}_${TAG}.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n${N}_${TAG}.json 2> gpurun_out/bench_n${N}_${TAG}.err
echo "bench N=$N rc=$?"; tail -c 3500 gpurun_out/bench_n${N}_${TAG}.json; tail -5 gpurun_out/bench_n${N}_${TAG}.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/bench_ref_n${N}_${TAG}.json 2> gpurun_out/bench_ref_n${N}_${TAG}.err
echo "ref N=$N rc=$?"; tail -c 800 gpurun_out/bench_ref_n${N}_${TAG}.json
