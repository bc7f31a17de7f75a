#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
O=gpurun_out
echo "== tests"; timeout 600 python -m pytest tests/test_gpu_models.py -m gpu -q -k "C5fused or pinned_host_tables" > $O/c8c_pytest.log 2>&1; echo "rc=$?"; grep -E "^E |passed|failed|Error" $O/c8c_pytest.log | head -30
echo "== bench scale 16"; CUDA_LAUNCH_BLOCKING=1 timeout 900 python bench.py --scale 16 --steps 20 --warmup 5 --no-cpu-baseline > $O/c8c_bench.json 2> $O/c8c_bench.err; echo "rc=$?"; cut -c1-200 $O/c8c_bench.json; grep -v "Warning\|warn" $O/c8c_bench.err | tail -25
