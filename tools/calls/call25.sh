#!/usr/bin/env bash
# grouped weight-gradient GEMM of the GCNII layers
set -u
mkdir -p gpurun_out
O=gpurun_out
echo "== kernel tests"; timeout 1200 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "gemm or colsum or sink" > $O/c25_pytest.log 2>&1; echo "rc=$?"; tail -2 $O/c25_pytest.log; grep -E "^E " $O/c25_pytest.log | head
echo "== gemm_bench"; timeout 300 python tools/gemm_bench.py > $O/c25_gemm_bench.jsonl 2> $O/c25_gemm_bench.err; echo "rc=$?"; cut -c1-200 $O/c25_gemm_bench.jsonl
echo "== bench grouped"; timeout 600 python bench.py --no-e2e --no-cpu-baseline > $O/c25_bench.json 2> $O/c25_bench.err; echo "rc=$?"; cut -c1-200 $O/c25_bench.json | grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*' | paste - -; tail -2 $O/c25_bench.err
echo "== bench per-layer"; INCAGG_WGRAD_GROUP=0 timeout 600 python bench.py --no-e2e --no-cpu-baseline > $O/c25_bench_nogroup.json 2> $O/c25_bench_nogroup.err; echo "rc=$?"; cut -c1-200 $O/c25_bench_nogroup.json | grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*' | paste - -
echo "== bench incagg grouped"; timeout 600 python bench.py --mode incagg --no-e2e --no-cpu-baseline > $O/c25_bench_incagg.json 2> $O/c25_bench_incagg.err; echo "rc=$?"; cut -c1-200 $O/c25_bench_incagg.json | grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*' | paste - -
echo "== timeline"; timeout 600 python tools/step_timeline.py > $O/c25_timeline.txt 2> $O/c25_timeline.err; echo "rc=$?"; tail -2 $O/c25_timeline.txt
echo "== model tests"; timeout 1500 python -m pytest tests/test_gpu_models.py -m gpu -x -q > $O/c25_pytest_models.log 2>&1; echo "rc=$?"; tail -3 $O/c25_pytest_models.log; grep -E "^E " $O/c25_pytest_models.log | head
