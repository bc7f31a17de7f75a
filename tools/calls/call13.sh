#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
O=gpurun_out
echo "== tests new variants"; timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "merge_path_kernel_edge_cases and (20 or 21 or 22)" > $O/c13_pytest.log 2>&1; echo "rc=$?"; tail -4 $O/c13_pytest.log; grep -E "^E " $O/c13_pytest.log | head
echo "== spmm_bench"; timeout 600 python tools/spmm_bench.py --batches 12 --cases fwd,bwd,full --variants rows,cpa16x3,cpa8x6 > $O/c13_spmm_bench.jsonl 2> $O/c13_spmm_bench.err; echo "rc=$?"; cut -c1-230 $O/c13_spmm_bench.jsonl; tail -3 $O/c13_spmm_bench.err
echo "== ncu"
CMD="python tools/spmm_bench.py --batches 2 --cases fwd --variants cpa16x3 --reps 1"
timeout 600 ncu --set full --clock-control none -k regex:spmm_cpasync -c 3 -o $O/c13_cpa_prof -f $CMD > $O/c13_ncu.log 2>&1; echo "ncu rc=$?"; tail -2 $O/c13_ncu.log | cut -c1-200
