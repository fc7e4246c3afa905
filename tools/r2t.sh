#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -s -k "social_pool" 2>&1 | grep "social pool\|passed\|failed\|Error" | tail -14
timeout 900 python -m pytest tests/test_gpu_fullsize.py -x -q -k "cfg3" 2>&1 | tail -2
timeout 400 python bench.py --config cfg3 --steps 5 --warmup 3 --no-train --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('cfg3 value %.0f ms %.3f'%(d['value'], d['ms_per_step']))
for k in (d.get('kernels') or [])[:6]: print('   %-40s %8.3f ms frac %.3f' % (k['kernel'],k['ms_per_step'],k['frac']))"
