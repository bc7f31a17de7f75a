#!/usr/bin/env bash
# GPU call 8 (1 GPU): full suite on the rebuilt library, default bench (with the graphed host-table leg),
# other configs, launch list + full captures for profiles/
set -u
mkdir -p gpurun_out
O=gpurun_out
echo "== pytest gpu (all)"; timeout 1800 python -m pytest tests -m gpu -q > $O/c8_pytest.log 2>&1; echo "rc=$?"; tail -6 $O/c8_pytest.log
echo "== bench n1 default"; timeout 900 python bench.py > $O/c8_bench_n1.json 2> $O/c8_bench_n1.err; echo "rc=$?"; cut -c1-300 $O/c8_bench_n1.json; tail -3 $O/c8_bench_n1.err; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/c8_bench_n1.json').read().strip().splitlines()[-1])
    print('value',d['value'],'ms',d['ms_per_step'],'launches',d['gpu_launches'])
    print('e2e',{k:(v if not isinstance(v,dict) else {kk:vv for kk,vv in v.items() if kk!='layout'}) for k,v in d['e2e'].items() if k!='layout'})
    print('roofline',{k:v for k,v in d['roofline'].items() if k not in ('l2_gather',)})
    print('cpu',d['cpu_baseline']); print('refresh',d['refresh']); print('clocks',d['clocks'])
except Exception as e: print('parse failed',e)
PY
echo "== bench n1 incagg"; timeout 900 python bench.py --mode incagg --no-cpu-baseline --no-e2e > $O/c8_bench_n1_incagg.json 2> $O/c8_bench_n1_incagg.err; echo "rc=$?"; cut -c1-260 $O/c8_bench_n1_incagg.json
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/c8_bench_ref.json 2> $O/c8_bench_ref.err; echo "rc=$?"; cut -c1-200 $O/c8_bench_ref.json
for c in C1 C2 C4 C5; do
  echo "== bench $c"; timeout 900 python bench.py --config $c --steps 6 --warmup 3 --no-e2e > $O/c8_bench_$c.json 2> $O/c8_bench_$c.err; echo "rc=$?"; cut -c1-260 $O/c8_bench_$c.json; tail -2 $O/c8_bench_$c.err
done
echo "== spmm_bench products"; timeout 600 python tools/spmm_bench.py --batches 12 --cases fwd,bwd,delta,full --variants rows,s4x3 > $O/c8_spmm_bench.jsonl 2> $O/c8_spmm_bench.err; echo "rc=$?"; cut -c1-200 $O/c8_spmm_bench.jsonl
echo "== ncu launch list of the timed region"
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline"
INCAGG_PROFILE=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/c8_launches.csv $CMD > $O/c8_ncu_launches.log 2>&1; echo "ncu rc=$?"; wc -l $O/c8_launches.csv
echo "== ncu full: kernels of the step"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"spmm_kernel|index_rows_kernel|gemm_tf32x3_kernel|gemm_nc_kernel|slice_bulk|map_edges" -c 30 -o $O/c8_step_prof -f env INCAGG_PROFILE=1 $CMD > $O/c8_ncu_full.log 2>&1; echo "ncu rc=$?"; tail -2 $O/c8_ncu_full.log | cut -c1-200
