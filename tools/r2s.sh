#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_util.py tests/test_gpu_parity.py tests/test_golden.py tests/test_gpu_model_api.py -m gpu -x -q 2>&1 | tail -3
for f in 0 1; do
DESIRE_NO_FUSE4=$f timeout 400 python bench.py --steps 20 --warmup 5 --no-train --no-cpu-baseline --breakdown gpurun_out/q_breakdown.json > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err; echo "bench (no_fuse4=$f) rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/q_bench.json').read().strip().splitlines()[-1])
print('value %.0f ms %.3f e2e %s'%(d['value'], d['ms_per_step'], d.get('e2e',{}).get('value')))
for k in (d.get('kernels') or [])[:8]: print("   %-40s %8.3f ms frac %.3f" % (k['kernel'],k['ms_per_step'],k['frac']))
PY
done
