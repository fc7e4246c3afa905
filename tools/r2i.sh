#!/bin/bash
# round 2, run I: factored feature_pooling projection in the IOC stage
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_gpu_fullsize.py -x -q -m gpu -k "not cfg5 and not cfg3" > gpurun_out/r2i_tests.log 2>&1; echo "tests rc=$?"
tail -4 gpurun_out/r2i_tests.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-train --no-cpu-baseline --breakdown gpurun_out/r2i_breakdown.json > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2i_bench.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'])
for k in d['kernels']: print("   %-40s %8.3f ms" % (k['kernel'],k['ms_per_step']))
PY
