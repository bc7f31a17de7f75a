#!/usr/bin/env bash
# GPU call 14: final validation on one GPU - full suite, smoke(), the driver's bench commands
set -u
mkdir -p gpurun_out
O=gpurun_out
echo "== pytest gpu (all)"; timeout 2400 python -m pytest tests -m gpu -q > $O/c14_pytest.log 2>&1; echo "rc=$?"; tail -5 $O/c14_pytest.log
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/c14_smoke.log 2>&1; echo "rc=$?"; tail -4 $O/c14_smoke.log
echo "== reference arm"; timeout 600 python3 bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/c14_bench_ref.json 2> $O/c14_bench_ref.err; echo "rc=$?"; cut -c1-160 $O/c14_bench_ref.json
echo "== bench literal"; timeout 900 python3 bench.py --gpus 1 --steps 20 --warmup 5 > $O/c14_bench_n1.json 2> $O/c14_bench_n1.err; echo "rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/c14_bench_n1.json').read().strip().splitlines()[-1])
print('value %.4g ms %.4f launches %s e2e %.4g host %s roof %s cpu %.3g' % (d['value'], d['ms_per_step'], d['gpu_launches'], d['e2e']['value'], {k:round(v['value']/1e6,1) for k,v in d['e2e'].items() if isinstance(v,dict) and 'value' in v}, d['roofline']['frac'], d['cpu_baseline']['value']))
PY
