#!/usr/bin/env bash
# grouped weight-gradient GEMM of the GCNII layers
set -u
mkdir -p gpurun_out
O=gpurun_out
echo "== bench grouped"; timeout 600 python bench.py --no-e2e --no-cpu-baseline > $O/c27_bench.json 2> $O/c27_bench.err; echo "rc=$?"; cut -c1-200 $O/c27_bench.json | grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*' | paste - -; tail -2 $O/c27_bench.err
echo "== timeline"; timeout 600 python tools/step_timeline.py > $O/c27_timeline.txt 2> $O/c27_timeline.err; echo "rc=$?"; tail -2 $O/c27_timeline.txt
