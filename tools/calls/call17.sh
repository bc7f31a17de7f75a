#!/usr/bin/env bash
# slim GEMM code paths + D2-accumulate epilogue + x_0 gradient sink + parallel colsum finish + CE gradient without autograd
set -u
mkdir -p gpurun_out
O=gpurun_out
echo "== kernel tests"; timeout 1200 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "gemm or colsum or sink" > $O/c17_pytest.log 2>&1; echo "rc=$?"; tail -3 $O/c17_pytest.log; grep -E "^E " $O/c17_pytest.log | head
for v in "" _base; do
  echo "== gemm_bench lib$v"; INCAGG_X0_SINK=0 INCAGG_B200_LIB=$PWD/incagg_gnn_b200/csrc/libincagg_b200$v.so timeout 300 python tools/gemm_bench.py > $O/c17_gemm_bench$v.jsonl 2> $O/c17_gemm_bench$v.err; echo "rc=$?"; cat $O/c17_gemm_bench$v.jsonl | cut -c1-200
  echo "== bench lib$v (no sink)"; INCAGG_X0_SINK=0 INCAGG_B200_LIB=$PWD/incagg_gnn_b200/csrc/libincagg_b200$v.so timeout 600 python bench.py --no-e2e --no-cpu-baseline > $O/c17_bench${v}_nosink.json 2> $O/c17_bench${v}_nosink.err; echo "rc=$?"; cut -c1-220 $O/c17_bench${v}_nosink.json
done
echo "== bench lib (sink)"; timeout 600 python bench.py --no-e2e --no-cpu-baseline > $O/c17_bench.json 2> $O/c17_bench.err; echo "rc=$?"; cut -c1-220 $O/c17_bench.json; tail -3 $O/c17_bench.err
echo "== model tests"; timeout 1500 python -m pytest tests/test_gpu_models.py -m gpu -x -q > $O/c17_pytest_models.log 2>&1; echo "rc=$?"; tail -3 $O/c17_pytest_models.log; grep -E "^E " $O/c17_pytest_models.log | head
