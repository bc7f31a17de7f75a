#!/usr/bin/env bash
# GPU call 31 (2 GPUs): bench.py after the launch-count change - the driver's commands at N = 1 and N = 2
set -u
mkdir -p gpurun_out
O=gpurun_out
echo "== bench N=1 literal"; timeout 600 python3 bench.py --gpus 1 --steps 20 --warmup 5 > $O/c31_bench_n1.json 2> $O/c31_bench_n1.err; echo "rc=$?"; grep -o '"value": [0-9.]*, "unit": "edges/s", "n_gpus": 1\|"gpu_launches": [0-9]*\|"ms_per_step": [0-9.]*, "higher' $O/c31_bench_n1.json; tail -2 $O/c31_bench_n1.err | cut -c1-200
echo "== bench N=2 literal"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 20 --warmup 5 > $O/c31_bench_n2.json 2> $O/c31_bench_n2.err; echo "rc=$?"; grep -o '"value": [0-9.]*, "unit": "edges/s", "n_gpus": 2\|"gpu_launches": [0-9]*\|"ms_per_step": [0-9.]*, "higher' $O/c31_bench_n2.json; tail -2 $O/c31_bench_n2.err | cut -c1-200
