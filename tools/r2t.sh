#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/t_tests.log 2>&1; echo "tests rc $?"; tail -3 gpurun_out/t_tests.log
for c in cfg3 cfg5; do
timeout 400 python bench.py --config $c --steps 5 --warmup 3 --no-train --no-cpu-baseline --breakdown gpurun_out/r2_${c}_breakdown.json 2>gpurun_out/r2_${c}_bench.err > gpurun_out/r2_${c}_bench.json
python -c "
import json,sys
d=json.loads(open('gpurun_out/r2_${c}_bench.json').read().strip().splitlines()[-1])
print('$c value %.0f ms %.3f e2e %.0f'%(d['value'], d['ms_per_step'], d['e2e']['value']))
for k in (d.get('kernels') or [])[:5]: print('   %-40s %8.3f ms frac %.3f' % (k['kernel'],k['ms_per_step'],k['frac']))"
done
