#!/usr/bin/env bash
# final sanity on the committed tree: smoke(), the driver's bench command
set -u
mkdir -p gpurun_out
O=gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/c35_smoke.log 2>&1; echo "rc=$?"; tail -3 $O/c35_smoke.log
echo "== bench literal"; timeout 400 python3 bench.py --gpus 1 --steps 20 --warmup 5 > $O/c35_bench_n1.json 2> $O/c35_bench_n1.err; echo "rc=$?"; grep -o '"value": [0-9.]*, "unit": "edges/s", "n_gpus": 1\|"gpu_launches": [0-9]*\|"ms_per_step": [0-9.]*, "higher' $O/c35_bench_n1.json
