#!/usr/bin/env bash
# GPU call 8d (1 GPU): the driver's literal N=1 command, the default run, launch list + small full captures
set -u
mkdir -p gpurun_out
O=gpurun_out
show() { python - "$1" <<'PY'
import json, sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print('value %.4g  ms/step %.4f  launches %s' % (d['value'], d['ms_per_step'], d.get('gpu_launches')))
    e=d.get('e2e') or {}
    print('e2e', {k:(v if not isinstance(v,dict) else {kk:vv for kk,vv in v.items() if kk!='layout'}) for k,v in e.items() if k!='layout'})
    print('roofline', {k:v for k,v in (d.get('roofline') or {}).items() if k in ('achieved','frac','us_per_launch','bytes_per_launch','l2_to_sm')})
    print('cpu', d.get('cpu_baseline')); print('refresh', d.get('refresh')); print('clocks', d.get('clocks'))
except Exception as e: print('parse failed', e)
PY
}
echo "== driver literal N=1"; timeout 900 python3 bench.py --gpus 1 --steps 20 --warmup 5 > $O/c8d_bench_n1_literal.json 2> $O/c8d_bench_n1_literal.err; echo "rc=$?"; show $O/c8d_bench_n1_literal.json; grep -v "Warn\|warn\|detach\|lv = " $O/c8d_bench_n1_literal.err | tail -5
echo "== default"; timeout 900 python bench.py > $O/c8d_bench_n1.json 2> $O/c8d_bench_n1.err; echo "rc=$?"; show $O/c8d_bench_n1.json; grep -v "Warn\|warn\|detach\|lv = " $O/c8d_bench_n1.err | tail -5
echo "== reference literal"; timeout 600 python3 bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/c8d_bench_ref.json 2> $O/c8d_bench_ref.err; echo "rc=$?"; cut -c1-200 $O/c8d_bench_ref.json
echo "== ncu launch list of the timed region"
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline"
INCAGG_PROFILE=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/c8d_launches.csv $CMD > $O/c8d_ncu_launches.log 2>&1; echo "ncu rc=$?"; wc -l $O/c8d_launches.csv
echo "== ncu full (12 launches, no source)"
timeout 900 ncu --set full --clock-control none --profile-from-start off -k regex:"spmm_kernel|index_rows_kernel<16|gemm_tf32x3_kernel|gemm_nc_kernel|slice_bulk" -c 14 -o $O/c8d_step_prof -f env INCAGG_PROFILE=1 $CMD > $O/c8d_ncu_full.log 2>&1; echo "ncu rc=$?"; ls -la $O/c8d_step_prof.ncu-rep
du -sh $O
