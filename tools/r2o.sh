#!/bin/bash
# round 2, run O: elected-lane MMA issue in gru_tc3 / gemm_tc / deconv_tc — tests, GRU microbench, headline bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/o_tests.log 2>&1; echo "tests rc $?"; tail -4 gpurun_out/o_tests.log
timeout 100 python tools/bench_gru.py --rows 38400 --hidden 128 --steps 12 --iters 20 2>&1 | tail -1
timeout 100 python tools/bench_gru.py --rows 327680 --hidden 256 --steps 12 --iters 5 2>&1 | tail -1
timeout 400 python bench.py --steps 20 --warmup 5 --breakdown gpurun_out/o_breakdown.json > gpurun_out/o_bench.json 2> gpurun_out/o_bench.err; echo "bench cfg2 rc=$?"
python - <<'PY'
import json
for f in ("o_bench",):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, 'value %.0f ms %.3f e2e %s train %s'%(d['value'], d['ms_per_step'], d.get('e2e',{}).get('value'), d.get('train_step')))
        for k in (d.get('kernels') or [])[:12]: print("   %-40s %8.3f ms frac %.3f" % (k['kernel'],k['ms_per_step'],k['frac']))
    except Exception as e: print(f,'ERR',e)
PY
tail -3 gpurun_out/o_bench.err
