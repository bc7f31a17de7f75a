#!/usr/bin/env bash
# GPU call 5: rows kernel with dynamic row scheduling + lean addressing, occupancy variants
set -u
mkdir -p gpurun_out
O=gpurun_out
export INCAGG_SPMM_STREAM=-1
echo "== pytest spmm"; timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "spmm and not merge_path" > $O/c5_pytest_spmm.log 2>&1; echo "rc=$?"; tail -3 $O/c5_pytest_spmm.log
for v in "" _m5 _m4; do
  echo "== spmm_bench lib$v"
  INCAGG_B200_LIB=$PWD/incagg_gnn_b200/csrc/libincagg_b200$v.so timeout 600 python tools/spmm_bench.py --batches 12 --cases fwd,bwd,delta,full --variants rows > $O/c5_spmm_bench$v.jsonl 2> $O/c5_spmm_bench$v.err; echo "rc=$?"; cut -c1-230 $O/c5_spmm_bench$v.jsonl; tail -2 $O/c5_spmm_bench$v.err
done
echo "== pytest new parity"; timeout 1500 python -m pytest tests/test_gpu_parity_full.py tests/test_gpu_kernels.py -m gpu -q -k "parity_full or out_of_range or full_size or full_degree or c5_slab or trajectory or push_only or aggregate_combined or fp64" > $O/c5_pytest_parity.log 2>&1; echo "rc=$?"; tail -15 $O/c5_pytest_parity.log
echo "== bench n1 rows"; timeout 600 python bench.py --no-e2e --no-cpu-baseline > $O/c5_bench_n1_rows.json 2> $O/c5_bench_n1_rows.err; echo "rc=$?"; cut -c1-300 $O/c5_bench_n1_rows.json
echo "== ncu"
CMD="python tools/spmm_bench.py --batches 2 --cases fwd --variants rows --reps 1"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmm_kernel -c 4 -o $O/c5_spmm_prof -f $CMD > $O/c5_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 $O/c5_ncu.log
