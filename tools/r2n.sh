#!/bin/bash
# round 2, run N: full GPU suite + headline bench with the second-design fused social kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/n_tests.log 2>&1; echo "tests rc $?"; tail -4 gpurun_out/n_tests.log
timeout 400 python bench.py --steps 20 --warmup 5 --breakdown gpurun_out/n_breakdown.json > gpurun_out/n_bench.json 2> gpurun_out/n_bench.err; echo "bench cfg2 rc=$?"
python - <<'PY'
import json
for f in ("n_bench",):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, 'value %.0f ms %.3f e2e %s roofline %s'%(d['value'], d['ms_per_step'], d.get('e2e',{}).get('value'), d.get('roofline')))
        for k in (d.get('kernels') or [])[:10]: print("   %-40s %8.3f ms frac %.3f" % (k['kernel'],k['ms_per_step'],k['frac']))
    except Exception as e: print(f,'ERR',e)
PY
tail -3 gpurun_out/n_bench.err
