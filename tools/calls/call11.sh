#!/usr/bin/env bash
# GPU call 11: compute-sanitizer memcheck of a pipelined run on the /16 twin (graph replays, e2e legs)
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --log-file $O/c11_memcheck.log python bench.py --scale 16 --steps 12 --warmup 3 --no-cpu-baseline > $O/c11_bench.json 2> $O/c11_bench.err; echo "rc=$?"
tail -15 $O/c11_memcheck.log; cut -c1-200 $O/c11_bench.json; tail -3 $O/c11_bench.err
