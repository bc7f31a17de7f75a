#!/usr/bin/env bash
# GPU call 29 (2 GPUs): multi-GPU re-validation after the step-scheduling changes
set -u
mkdir -p gpurun_out
O=gpurun_out
echo "== multi_gpu_check"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/multi_gpu_check.py > $O/c29_mgc.log 2>&1; echo "rc=$?"; grep -v "^\[W\|^W1\|Warning" $O/c29_mgc.log | tail -25
echo "== bench N=2 driver-literal"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 20 --warmup 5 > $O/c29_bench_n2.json 2> $O/c29_bench_n2.err; echo "rc=$?"; cut -c1-900 $O/c29_bench_n2.json; tail -5 $O/c29_bench_n2.err
echo "== bench N=2 NCCL all-reduce"; INCAGG_FUSED_ALLREDUCE=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29614 bench.py --gpus 2 --steps 40 --warmup 5 --no-cpu-baseline --no-e2e > $O/c29_bench_n2_nccl.json 2> $O/c29_bench_n2_nccl.err; echo "rc=$?"; cut -c1-300 $O/c29_bench_n2_nccl.json; tail -3 $O/c29_bench_n2_nccl.err
