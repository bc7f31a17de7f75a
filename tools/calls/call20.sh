#!/usr/bin/env bash
# original GEMM structure + D2-accumulate epilogue; x_0 gradient sink; parallel colsum finish; CE gradient without autograd
set -u
mkdir -p gpurun_out
O=gpurun_out
echo "== kernel tests"; timeout 1200 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "gemm or colsum or sink" > $O/c20_pytest.log 2>&1; echo "rc=$?"; tail -3 $O/c20_pytest.log; grep -E "^E " $O/c20_pytest.log | head
echo "== gemm_bench"; timeout 300 python tools/gemm_bench.py > $O/c20_gemm_bench.jsonl 2> $O/c20_gemm_bench.err; echo "rc=$?"; cat $O/c20_gemm_bench.jsonl | cut -c1-200
echo "== bench (no sink)"; INCAGG_X0_SINK=0 timeout 600 python bench.py --no-e2e --no-cpu-baseline > $O/c20_bench_nosink.json 2> $O/c20_bench_nosink.err; echo "rc=$?"; cut -c1-220 $O/c20_bench_nosink.json
echo "== bench (sink)"; timeout 600 python bench.py --no-e2e --no-cpu-baseline > $O/c20_bench.json 2> $O/c20_bench.err; echo "rc=$?"; cut -c1-220 $O/c20_bench.json; tail -3 $O/c20_bench.err
echo "== timeline"; timeout 600 python tools/step_timeline.py > $O/c20_timeline.txt 2> $O/c20_timeline.err; echo "rc=$?"; tail -3 $O/c20_timeline.txt; tail -3 $O/c20_timeline.err
