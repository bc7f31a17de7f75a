#!/usr/bin/env bash
# GEMM register-prefetch depth: default build (GEMM_PREFETCH=2) vs _pf1 (round-2 baseline) vs _pf3
set -u
mkdir -p gpurun_out
O=gpurun_out
echo "== gemm tests"; timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "gemm" > $O/c16_pytest.log 2>&1; echo "rc=$?"; tail -3 $O/c16_pytest.log; grep -E "^E " $O/c16_pytest.log | head
for v in "" _pf1 _pf3; do
  echo "== gemm_bench lib$v"; INCAGG_B200_LIB=$PWD/incagg_gnn_b200/csrc/libincagg_b200$v.so timeout 300 python tools/gemm_bench.py > $O/c16_gemm_bench$v.jsonl 2> $O/c16_gemm_bench$v.err; echo "rc=$?"; cat $O/c16_gemm_bench$v.jsonl | cut -c1-200
  echo "== bench lib$v"; INCAGG_B200_LIB=$PWD/incagg_gnn_b200/csrc/libincagg_b200$v.so timeout 600 python bench.py --no-e2e --no-cpu-baseline > $O/c16_bench$v.json 2> $O/c16_bench$v.err; echo "rc=$?"; cut -c1-220 $O/c16_bench$v.json
done
echo "== gemm tests pf3"; INCAGG_B200_LIB=$PWD/incagg_gnn_b200/csrc/libincagg_b200_pf3.so timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "gemm" > $O/c16_pytest_pf3.log 2>&1; echo "rc=$?"; tail -2 $O/c16_pytest_pf3.log
