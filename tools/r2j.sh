#!/bin/bash
# round 2, run J: TMA-store epilogue of the tall tensor-core GEMMs; full GPU suite
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2j_gpu_tests.log 2>&1; echo "gpu tests rc=$?"
tail -5 gpurun_out/r2j_gpu_tests.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --breakdown gpurun_out/r2j_breakdown.json > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2j_bench.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['train_step']['ms_per_step'])
for k in d['kernels']: print("   %-40s %8.3f ms" % (k['kernel'],k['ms_per_step']))
PY
DESIRE_GEMM_NO_TMA_STORE=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2j_bench_nostore.json 2>/dev/null
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2j_bench_nostore.json'))
print("no TMA store:", d['value'], d['ms_per_step'], d['train_step']['ms_per_step'])
PY
