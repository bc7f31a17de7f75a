#!/usr/bin/env bash
# GPU call 10 (8 GPUs): fused (vectorised) vs NCCL gradient exchange at N=8, C5 with both transports, 8-rank check
set -u
mkdir -p gpurun_out
O=gpurun_out
run() { local name=$1 np=$2 port=$3; shift 3
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port $port bench.py --gpus $np "$@" > $O/c10_$name.json 2> $O/c10_$name.err
  echo "$name rc=$?"; cut -c1-300 $O/c10_$name.json; grep -i "error\|Traceback" $O/c10_$name.err | head -3
}
run n8_step 8 29801 --no-e2e --no-cpu-baseline
INCAGG_FUSED_ALLREDUCE=0 run n8_step_nccl_allreduce 8 29802 --no-e2e --no-cpu-baseline
run n8_c5_nccl 8 29803 --config C5 --transport nccl --steps 6 --warmup 3 --no-e2e
run n8_c5_p2p 8 29804 --config C5 --transport p2p --steps 6 --warmup 3 --no-e2e
run n8_literal 8 29805 --steps 20 --warmup 5
echo "== multi_gpu_check 4 ranks"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29806 tools/multi_gpu_check.py > $O/c10_mgc4.log 2>&1; echo "rc=$?"; grep "fused\|one-graph\|OK\|False" $O/c10_mgc4.log | head -20
