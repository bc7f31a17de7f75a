#!/usr/bin/env bash
# GPU call 1 of round 2: parity suite on the merged branches, SpMM variants, bench (N=1), reference arm, ncu.
set -u
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/c1_smi.txt 2>&1
echo "== pytest gpu" ; timeout 900 python -m pytest tests -m gpu -x -q > $O/c1_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/c1_pytest.log; tail -5 $O/c1_pytest.log
echo "== spmm_bench"; timeout 600 python tools/spmm_bench.py --batches 12 > $O/c1_spmm_bench.jsonl 2> $O/c1_spmm_bench.err; echo "rc=$?"; cat $O/c1_spmm_bench.jsonl
echo "== bench n1"; timeout 900 python bench.py > $O/c1_bench_n1.json 2> $O/c1_bench_n1.err; echo "rc=$?"; cut -c1-600 $O/c1_bench_n1.json
echo "== bench n1 PDL"; INCAGG_PDL=1 timeout 600 python bench.py --no-e2e --no-cpu-baseline > $O/c1_bench_n1_pdl.json 2> $O/c1_bench_n1_pdl.err; echo "rc=$?"; cut -c1-300 $O/c1_bench_n1_pdl.json
echo "== bench n1 rows kernel"; INCAGG_SPMM_STREAM=-1 timeout 600 python bench.py --no-e2e --no-cpu-baseline > $O/c1_bench_n1_rows.json 2> $O/c1_bench_n1_rows.err; echo "rc=$?"; cut -c1-300 $O/c1_bench_n1_rows.json
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/c1_bench_ref.json 2> $O/c1_bench_ref.err; echo "rc=$?"; cut -c1-400 $O/c1_bench_ref.json
echo "== ncu spmm"
CMD="python tools/spmm_bench.py --batches 2 --cases fwd --variants rows,stream16x2 --reps 1"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmm_ -c 12 -o $O/c1_spmm_prof -f $CMD > $O/c1_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 $O/c1_ncu.log
