#!/usr/bin/env bash
# high-priority pull stream (A/B against default priority)
set -u
mkdir -p gpurun_out
O=gpurun_out
echo "== bench prio -1"; timeout 600 python bench.py --no-e2e --no-cpu-baseline > $O/c34_bench.json 2> $O/c34_bench.err; echo "rc=$?"; grep -o '"value": [0-9.]*, "unit": "edges/s", "n_gpus": 1\|"ms_per_step": [0-9.]*, "higher' $O/c34_bench.json; tail -2 $O/c34_bench.err | cut -c1-200
echo "== bench prio 0"; INCAGG_PULL_PRIORITY=0 timeout 600 python bench.py --no-e2e --no-cpu-baseline > $O/c34_bench_p0.json 2> $O/c34_bench_p0.err; echo "rc=$?"; grep -o '"value": [0-9.]*, "unit": "edges/s", "n_gpus": 1\|"ms_per_step": [0-9.]*, "higher' $O/c34_bench_p0.json
echo "== timeline"; timeout 600 python tools/step_timeline.py > $O/c34_timeline.txt 2> $O/c34_timeline.err; echo "rc=$?"; tail -2 $O/c34_timeline.txt
