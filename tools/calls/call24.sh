#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
O=gpurun_out
echo "== kernel tests"; timeout 1200 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "colsum" > $O/c24_pytest.log 2>&1; echo "rc=$?"; tail -2 $O/c24_pytest.log
echo "== bench"; timeout 600 python bench.py --no-e2e --no-cpu-baseline > $O/c24_bench.json 2> $O/c24_bench.err; echo "rc=$?"; cut -c1-200 $O/c24_bench.json | grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*' | paste - -
echo "== bench no wgrad stream"; INCAGG_WGRAD_STREAM=0 timeout 600 python bench.py --no-e2e --no-cpu-baseline > $O/c24_bench_nowgrad.json 2> $O/c24_bench_nowgrad.err; echo "rc=$?"; cut -c1-200 $O/c24_bench_nowgrad.json | grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*' | paste - -
echo "== bench incagg"; timeout 600 python bench.py --mode incagg --no-e2e --no-cpu-baseline > $O/c24_bench_incagg.json 2> $O/c24_bench_incagg.err; echo "rc=$?"; cut -c1-200 $O/c24_bench_incagg.json | grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*' | paste - -; tail -2 $O/c24_bench_incagg.err
echo "== timeline"; timeout 600 python tools/step_timeline.py > $O/c24_timeline.txt 2> $O/c24_timeline.err; echo "rc=$?"; tail -2 $O/c24_timeline.txt
