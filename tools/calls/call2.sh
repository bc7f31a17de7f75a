#!/usr/bin/env bash
# GPU call 2: merge-path kernel v3 - edge-case tests, variants, ncu of the best two.
set -u
mkdir -p gpurun_out
O=gpurun_out
echo "== pytest spmm"; timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "spmm" > $O/c2_pytest_spmm.log 2>&1; echo "rc=$?"; tail -5 $O/c2_pytest_spmm.log
echo "== spmm_bench"; timeout 900 python tools/spmm_bench.py --batches 12 > $O/c2_spmm_bench.jsonl 2> $O/c2_spmm_bench.err; echo "rc=$?"; cat $O/c2_spmm_bench.jsonl; tail -3 $O/c2_spmm_bench.err
echo "== pytest all"; timeout 1200 python -m pytest tests -m gpu -q > $O/c2_pytest.log 2>&1; echo "rc=$?"; tail -8 $O/c2_pytest.log
echo "== ncu spmm"
CMD="python tools/spmm_bench.py --batches 2 --cases fwd --variants stream16x2,stream8x3,stream4x6 --reps 1"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmm_stream -c 9 -o $O/c2_spmm_prof -f $CMD > $O/c2_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 $O/c2_ncu.log
