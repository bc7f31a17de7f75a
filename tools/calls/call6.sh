#!/usr/bin/env bash
# GPU call 6 (2 GPUs): fused peer-memory gradient exchange, one-graph multi-GPU step, driver-literal bench at N=2
set -u
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L > $O/c6_smi.txt 2>&1
echo "== spmm_bench (static rows, lean addressing)"
for v in "" _m5; do
  INCAGG_B200_LIB=$PWD/incagg_gnn_b200/csrc/libincagg_b200$v.so timeout 600 python tools/spmm_bench.py --batches 12 --cases fwd,bwd,delta,full --variants rows,s4x3 > $O/c6_spmm_bench$v.jsonl 2> $O/c6_spmm_bench$v.err; echo "rc=$?"; cut -c1-200 $O/c6_spmm_bench$v.jsonl; tail -2 $O/c6_spmm_bench$v.err
done
echo "== multi_gpu_check"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/multi_gpu_check.py > $O/c6_mgc.log 2>&1; echo "rc=$?"; grep -v "^\[W\|^W1\|Warning" $O/c6_mgc.log | tail -40
echo "== bench N=2 driver-literal"; timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 20 --warmup 5 > $O/c6_bench_n2.json 2> $O/c6_bench_n2.err; echo "rc=$?"; cut -c1-1500 $O/c6_bench_n2.json; tail -5 $O/c6_bench_n2.err
echo "== bench N=2 full epoch"; timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --no-cpu-baseline --no-e2e > $O/c6_bench_n2_full.json 2> $O/c6_bench_n2_full.err; echo "rc=$?"; cut -c1-400 $O/c6_bench_n2_full.json
echo "== bench N=2 NCCL all-reduce"; INCAGG_FUSED_ALLREDUCE=0 timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29614 bench.py --gpus 2 --no-cpu-baseline --no-e2e > $O/c6_bench_n2_nccl.json 2> $O/c6_bench_n2_nccl.err; echo "rc=$?"; cut -c1-400 $O/c6_bench_n2_nccl.json
echo "== bench N=1"; timeout 900 python bench.py --steps 20 --warmup 5 > $O/c6_bench_n1.json 2> $O/c6_bench_n1.err; echo "rc=$?"; cut -c1-400 $O/c6_bench_n1.json
echo "== reference arm N=2"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29615 bench.py --impl reference --gpus 2 --steps 20 --warmup 5 > $O/c6_bench_ref_n2.json 2> $O/c6_bench_ref_n2.err; echo "rc=$?"; cut -c1-300 $O/c6_bench_ref_n2.json
