#!/usr/bin/env bash
# halo pulls issued before the first Linear
set -u
mkdir -p gpurun_out
O=gpurun_out
echo "== bench"; timeout 600 python bench.py --no-e2e --no-cpu-baseline > $O/c33_bench.json 2> $O/c33_bench.err; echo "rc=$?"; grep -o '"value": [0-9.]*, "unit": "edges/s", "n_gpus": 1\|"ms_per_step": [0-9.]*, "higher' $O/c33_bench.json; tail -2 $O/c33_bench.err | cut -c1-200
echo "== timeline"; timeout 600 python tools/step_timeline.py > $O/c33_timeline.txt 2> $O/c33_timeline.err; echo "rc=$?"; tail -2 $O/c33_timeline.txt
echo "== model tests"; timeout 1500 python -m pytest tests/test_gpu_models.py -m gpu -x -q > $O/c33_pytest_models.log 2>&1; echo "rc=$?"; tail -3 $O/c33_pytest_models.log; grep -E "^E " $O/c33_pytest_models.log | head
