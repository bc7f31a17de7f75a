#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
O=gpurun_out
echo "== gemm tests"; timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "gemm" > $O/c15_pytest.log 2>&1; echo "rc=$?"; tail -3 $O/c15_pytest.log; grep -E "^E " $O/c15_pytest.log | head
for v in "" _nostage; do
  echo "== gemm_bench lib$v"; INCAGG_B200_LIB=$PWD/incagg_gnn_b200/csrc/libincagg_b200$v.so timeout 300 python tools/gemm_bench.py > $O/c15_gemm_bench$v.jsonl 2> $O/c15_gemm_bench$v.err; echo "rc=$?"; cat $O/c15_gemm_bench$v.jsonl | cut -c1-200
  echo "== bench lib$v"; INCAGG_B200_LIB=$PWD/incagg_gnn_b200/csrc/libincagg_b200$v.so timeout 600 python bench.py --no-e2e --no-cpu-baseline > $O/c15_bench$v.json 2> $O/c15_bench$v.err; echo "rc=$?"; cut -c1-220 $O/c15_bench$v.json
done
echo "== model tests"; timeout 1200 python -m pytest tests/test_gpu_models.py -m gpu -x -q > $O/c15_pytest_models.log 2>&1; echo "rc=$?"; tail -3 $O/c15_pytest_models.log
