#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
O=gpurun_out
echo "== C5fused test"; timeout 600 python -m pytest tests/test_gpu_models.py -m gpu -q -x -k "C5fused or pinned_host_tables" > $O/c8b_pytest.log 2>&1; echo "rc=$?"; grep -E "^E |passed|failed|Error" $O/c8b_pytest.log | head -30
echo "== bench default short"; timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/c8b_bench.json 2> $O/c8b_bench.err; echo "rc=$?"; cut -c1-200 $O/c8b_bench.json; grep -v "Warning\|warn" $O/c8b_bench.err | tail -30
