#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r_tests.log 2>&1; echo "tests rc $?"; tail -3 gpurun_out/r_tests.log
timeout 400 python bench.py --steps 20 --warmup 5 --breakdown gpurun_out/q_breakdown.json > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/q_bench.json').read().strip().splitlines()[-1])
print('value %.0f ms %.3f e2e %s train %s'%(d['value'], d['ms_per_step'], d.get('e2e',{}).get('value'), (d.get('train_step') or {}).get('ms_per_step')))
for k in (d.get('kernels') or [])[:12]: print("   %-40s %8.3f ms frac %.3f" % (k['kernel'],k['ms_per_step'],k['frac']))
PY
tail -2 gpurun_out/q_bench.err
