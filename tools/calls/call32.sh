#!/usr/bin/env bash
# register cap of the 128-wide GEMM kernels (184 instead of up to 221)
set -u
mkdir -p gpurun_out
O=gpurun_out
for v in "" _mr184; do
  echo "== bench lib$v"; INCAGG_B200_LIB=$PWD/incagg_gnn_b200/csrc/libincagg_b200$v.so timeout 600 python bench.py --no-e2e --no-cpu-baseline > $O/c32_bench$v.json 2> $O/c32_bench$v.err; echo "rc=$?"; grep -o '"value": [0-9.]*, "unit": "edges/s", "n_gpus": 1\|"ms_per_step": [0-9.]*, "higher' $O/c32_bench$v.json
done
echo "== gemm tests mr184"; INCAGG_B200_LIB=$PWD/incagg_gnn_b200/csrc/libincagg_b200_mr184.so timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "gemm" > $O/c32_pytest.log 2>&1; echo "rc=$?"; tail -2 $O/c32_pytest.log
echo "== gemm_bench mr184"; INCAGG_B200_LIB=$PWD/incagg_gnn_b200/csrc/libincagg_b200_mr184.so timeout 300 python tools/gemm_bench.py > $O/c32_gemm_bench_mr184.jsonl 2>/dev/null; cut -c1-200 $O/c32_gemm_bench_mr184.jsonl
