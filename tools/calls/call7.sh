#!/usr/bin/env bash
# GPU call 7 (1 GPU): full GPU suite, default bench, kernel micro-benchmarks, ncu launch list + full captures
set -u
mkdir -p gpurun_out
O=gpurun_out
echo "== pytest gpu (all)"; timeout 1800 python -m pytest tests -m gpu -q > $O/c7_pytest.log 2>&1; echo "rc=$?"; tail -6 $O/c7_pytest.log
echo "== bench n1 default"; timeout 900 python bench.py > $O/c7_bench_n1.json 2> $O/c7_bench_n1.err; echo "rc=$?"; cut -c1-700 $O/c7_bench_n1.json; tail -3 $O/c7_bench_n1.err
echo "== bench n1 incagg"; timeout 900 python bench.py --mode incagg --no-cpu-baseline > $O/c7_bench_n1_incagg.json 2> $O/c7_bench_n1_incagg.err; echo "rc=$?"; cut -c1-300 $O/c7_bench_n1_incagg.json
echo "== spmm_bench products"; timeout 600 python tools/spmm_bench.py --batches 12 --cases fwd,bwd,delta,full --variants rows,s4x3 > $O/c7_spmm_bench.jsonl 2> $O/c7_spmm_bench.err; echo "rc=$?"; cut -c1-200 $O/c7_spmm_bench.jsonl
echo "== spmm_bench reddit F=1024"; timeout 600 python tools/spmm_bench.py --shape reddit --F 1024 --batch-parts 20 --batches 2 --cases fwd --variants rows,s4x3 > $O/c7_spmm_bench_reddit.jsonl 2> $O/c7_spmm_bench_reddit.err; echo "rc=$?"; cut -c1-200 $O/c7_spmm_bench_reddit.jsonl
echo "== spmm_bench amazon F=256"; timeout 600 python tools/spmm_bench.py --shape amazonproducts --F 256 --batch-parts 4 --batches 2 --cases fwd --variants rows,s4x3 > $O/c7_spmm_bench_amazon.jsonl 2> $O/c7_spmm_bench_amazon.err; echo "rc=$?"; cut -c1-200 $O/c7_spmm_bench_amazon.jsonl
echo "== kernel_bench"; timeout 600 python tools/kernel_bench.py --tag r02 > $O/c7_kernel_bench.jsonl 2> $O/c7_kernel_bench.err; echo "rc=$?"; cut -c1-220 $O/c7_kernel_bench.jsonl | head -30
echo "== ncu launch list of the timed region"
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline"
INCAGG_PROFILE=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/c7_launches.csv $CMD > $O/c7_ncu_launches.log 2>&1; echo "ncu rc=$?"; tail -2 $O/c7_ncu_launches.log | cut -c1-300; wc -l $O/c7_launches.csv
echo "== ncu full: spmm rows, gather, gemm in the step"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"spmm_kernel|index_rows_kernel|gemm_tf32x3_kernel|slice_bulk" -c 24 -o $O/c7_step_prof -f env INCAGG_PROFILE=1 $CMD > $O/c7_ncu_full.log 2>&1; echo "ncu rc=$?"; tail -2 $O/c7_ncu_full.log | cut -c1-300
