#!/usr/bin/env bash
# original GEMM structure + D2-accumulate epilogue; x_0 gradient sink; parallel colsum finish; CE gradient without autograd
set -u
mkdir -p gpurun_out
O=gpurun_out
echo "== kernel tests"; timeout 1200 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "gemm or colsum or sink or adam or linear" > $O/c23_pytest.log 2>&1; echo "rc=$?"; tail -3 $O/c23_pytest.log; grep -E "^E " $O/c23_pytest.log | head
echo "== gemm_bench"; timeout 300 python tools/gemm_bench.py > $O/c23_gemm_bench.jsonl 2> $O/c23_gemm_bench.err; echo "rc=$?"; cat $O/c23_gemm_bench.jsonl | cut -c1-200
echo "== bench (no sink)"; INCAGG_X0_SINK=0 timeout 600 python bench.py --no-e2e --no-cpu-baseline > $O/c23_bench_nosink.json 2> $O/c23_bench_nosink.err; echo "rc=$?"; cut -c1-220 $O/c23_bench_nosink.json
echo "== bench (sink)"; timeout 600 python bench.py --no-e2e --no-cpu-baseline > $O/c23_bench.json 2> $O/c23_bench.err; echo "rc=$?"; cut -c1-220 $O/c23_bench.json; tail -3 $O/c23_bench.err
echo "== timeline"; timeout 600 python tools/step_timeline.py > $O/c23_timeline.txt 2> $O/c23_timeline.err; echo "rc=$?"; tail -3 $O/c23_timeline.txt; tail -3 $O/c23_timeline.err
echo "== model tests"; timeout 1500 python -m pytest tests/test_gpu_models.py -m gpu -x -q > $O/c23_pytest_models.log 2>&1; echo "rc=$?"; tail -3 $O/c23_pytest_models.log; grep -E "^E " $O/c23_pytest_models.log | head
