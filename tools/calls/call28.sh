#!/usr/bin/env bash
# GPU call 28: final validation on one GPU - full suite, smoke(), the driver's bench commands
set -u
mkdir -p gpurun_out
O=gpurun_out
echo "== pytest gpu (all)"; timeout 2400 python -m pytest tests -m gpu -q > $O/c28_pytest.log 2>&1; echo "rc=$?"; tail -5 $O/c28_pytest.log
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/c28_smoke.log 2>&1; echo "rc=$?"; tail -4 $O/c28_smoke.log
echo "== reference arm"; timeout 600 python3 bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/c28_bench_ref.json 2> $O/c28_bench_ref.err; echo "rc=$?"; cut -c1-160 $O/c28_bench_ref.json
echo "== bench literal"; timeout 900 python3 bench.py --gpus 1 --steps 20 --warmup 5 > $O/c28_bench_n1.json 2> $O/c28_bench_n1.err; echo "rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/c28_bench_n1.json').read().strip().splitlines()[-1])
print('value %.4g ms %.4f launches %s e2e %.4g host %s roof %s cpu %.3g' % (d['value'], d['ms_per_step'], d['gpu_launches'], d['e2e']['value'], {k:round(v['value']/1e6,1) for k,v in d['e2e'].items() if isinstance(v,dict) and 'value' in v}, d['roofline']['frac'], d['cpu_baseline']['value']))
PY
echo "== bench default"; timeout 900 python bench.py > $O/c28_bench_default.json 2> $O/c28_bench_default.err; echo "rc=$?"; cut -c1-200 $O/c28_bench_default.json | grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*' | paste - -
echo "== ncu launch list of the timed region"
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline"
INCAGG_PROFILE=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/c28_launches.csv $CMD > $O/c28_ncu_launches.log 2>&1; echo "ncu rc=$?"; tail -2 $O/c28_ncu_launches.log | cut -c1-300; wc -l $O/c28_launches.csv
echo "== timeline (final)"; timeout 600 python tools/step_timeline.py > $O/c28_timeline.txt 2> $O/c28_timeline.err; echo "rc=$?"; tail -2 $O/c28_timeline.txt
echo "== timeline incagg"; timeout 600 python tools/step_timeline.py C3 incagg > $O/c28_timeline_incagg.txt 2> $O/c28_timeline_incagg.err; echo "rc=$?"; tail -2 $O/c28_timeline_incagg.txt
