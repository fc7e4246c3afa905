#!/bin/bash
# round 2, run F: GRU v3 with TMA stores of the states
mkdir -p gpurun_out
DESIRE_GRU3_WATCHDOG=1 timeout 200 python -m pytest tests/test_gpu_gru.py tests/test_gpu_parity.py tests/test_golden.py -x -q -k "recurrence or single_bin or golden" > gpurun_out/r2f_tests.log 2>&1; echo "tests rc=$?"
tail -4 gpurun_out/r2f_tests.log
for cfg in "38400 128" "327680 256" "18944 128" "18944 256"; do set -- $cfg
  echo -n "v3 "; timeout 60 python tools/bench_gru.py --rows $1 --hidden $2 --steps 12 2>&1 | tail -1
done | tee gpurun_out/r2f_bench_gru.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-train --no-cpu-baseline --breakdown gpurun_out/r2f_breakdown.json > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2f_bench.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'])
for k in d['kernels']: print("   %-40s %8.3f ms" % (k['kernel'],k['ms_per_step']))
PY
