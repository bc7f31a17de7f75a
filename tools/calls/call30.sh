#!/usr/bin/env bash
# GPU call 30 (8 GPUs): the driver's literal scaling command after the step-scheduling changes
set -u
mkdir -p gpurun_out
O=gpurun_out
echo "== bench N=8 driver-literal"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 8 --steps 20 --warmup 5 > $O/c30_bench_n8.json 2> $O/c30_bench_n8.err; echo "rc=$?"; cut -c1-300 $O/c30_bench_n8.json; grep -o '"e2e": {"value": [0-9.]*' $O/c30_bench_n8.json; tail -3 $O/c30_bench_n8.err | cut -c1-200
