#!/usr/bin/env bash
# GPU call 3: merge-path kernel v4 (double-buffered batches)
set -u
mkdir -p gpurun_out
O=gpurun_out
echo "== pytest spmm"; timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "merge_path or spmm_gated" > $O/c4_pytest_spmm.log 2>&1; echo "rc=$?"; tail -3 $O/c4_pytest_spmm.log
echo "== spmm_bench"; timeout 900 python tools/spmm_bench.py --batches 12 --cases fwd,bwd,delta,full > $O/c4_spmm_bench.jsonl 2> $O/c4_spmm_bench.err; echo "rc=$?"; cut -c1-260 $O/c4_spmm_bench.jsonl; tail -3 $O/c4_spmm_bench.err
echo "== ncu spmm"
CMD="python tools/spmm_bench.py --batches 2 --cases fwd --variants p4x2,p2x3,p8x1 --reps 1"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmm_pair -c 9 -o $O/c4_spmm_prof -f $CMD > $O/c4_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 $O/c4_ncu.log
