#!/usr/bin/env bash
# knobs: CTA budget of the side-stream weight-gradient GEMMs; 64-wide n-tiles for many-tile problems
set -u
mkdir -p gpurun_out
O=gpurun_out
echo "== gemm_bench bn64 for >= 300 m-tiles"; INCAGG_TUNE="gemm_bn64_min_tiles=300" timeout 300 python tools/gemm_bench.py > $O/c22_gemm_bench_bn64.jsonl 2> $O/c22_gemm_bench_bn64.err; echo "rc=$?"; grep lins0 $O/c22_gemm_bench_bn64.jsonl | cut -c1-200
for t in "" "gemm_dual_m_ctas=96" "gemm_dual_m_ctas=64" "gemm_dual_m_ctas=32" "gemm_dual_m_ctas=16" "gemm_bn64_min_tiles=300" ; do
  n=$(echo "$t" | tr '=,' '__'); echo "== bench INCAGG_TUNE=$t"
  INCAGG_TUNE="$t" timeout 600 python bench.py --no-e2e --no-cpu-baseline > $O/c22_bench_$n.json 2> $O/c22_bench_$n.err; echo "rc=$?"; cut -c1-200 $O/c22_bench_$n.json | grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*' | paste - -
done
