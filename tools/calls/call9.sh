#!/usr/bin/env bash
# GPU call 9 (8 GPUs): scaling evidence - the driver's literal command at N=8, N=4 / N=8 step-only runs,
# fused vs NCCL gradient exchange, C5 (PNA, amazon-products shape) with NCCL halo exchange
set -u
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L > $O/c9_smi.txt 2>&1
run() { # name nproc port args...
  local name=$1 np=$2 port=$3; shift 3
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port $port bench.py --gpus $np "$@" > $O/c9_$name.json 2> $O/c9_$name.err
  echo "$name rc=$?"; cut -c1-330 $O/c9_$name.json; grep -i "error\|Traceback" $O/c9_$name.err | head -3
}
run n8_literal 8 29701 --steps 20 --warmup 5
run n8_step 8 29702 --no-e2e --no-cpu-baseline
INCAGG_FUSED_ALLREDUCE=0 run n8_step_nccl_allreduce 8 29703 --no-e2e --no-cpu-baseline
run n4_step 4 29704 --no-e2e --no-cpu-baseline
run n2_step 2 29705 --no-e2e --no-cpu-baseline
run n8_incagg 8 29706 --mode incagg --no-e2e --no-cpu-baseline
run n8_c5_nccl 8 29707 --config C5 --transport nccl --steps 6 --warmup 3 --no-e2e
run n8_c5_p2p 8 29708 --config C5 --transport p2p --steps 6 --warmup 3 --no-e2e
