#!/bin/bash
# round 2, last evidence pass after the scene-CNN kernel (conv5_tc.cu): GPU tests, bench lines, launch list, scene-CNN
# microbenchmarks (both paths, per launch), MMA cost vs N, one ncu --set full capture of the layer-3 kernel
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r2_final_gputests.txt; cat gpurun_out/r2_final_gputests.txt
timeout -s KILL 400 python bench.py --steps 20 --warmup 5 --breakdown gpurun_out/r2_final_breakdown.json > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err; echo "bench cfg2 rc=$?"
timeout -s KILL 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_final_bench_ref.json 2>/dev/null; echo "bench ref rc=$?"
timeout -s KILL 300 python bench.py --config cfg3 --steps 5 --warmup 3 --no-train --no-cpu-baseline --breakdown gpurun_out/r2_cfg3_breakdown.json > gpurun_out/r2_cfg3_bench.json 2> gpurun_out/r2_cfg3_bench.err; echo "bench cfg3 rc=$?"
timeout -s KILL 300 python bench.py --config cfg5 --steps 5 --warmup 3 --no-train --no-cpu-baseline --breakdown gpurun_out/r2_cfg5_breakdown.json > gpurun_out/r2_cfg5_bench.json 2> gpurun_out/r2_cfg5_bench.err; echo "bench cfg5 rc=$?"
timeout -s KILL 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_launches_fwd.csv python tools/profile_kernels.py --passes 1 --ioc-iters 2 > /dev/null 2>&1; echo "launch list rc=$?"
{
  timeout -s KILL 60 python tools/bench_scene_cnn.py 2>&1 | tail -1
  DESIRE_NO_CONV5=1 timeout -s KILL 60 python tools/bench_scene_cnn.py 2>&1 | tail -1
  DESIRE_CONV5_TRACE=1 timeout -s KILL 60 python tools/bench_scene_cnn.py 2>&1 | grep "conv5 trace"
  for f in 0 1; do
    DESIRE_NO_CONV5=$f timeout -s KILL 100 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/c5_l$f.csv python tools/bench_scene_cnn.py > /dev/null 2>&1
    echo "per launch under ncu (cold, serialised), DESIRE_NO_CONV5=$f:"
    python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/c5_l$f.csv')))
h=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
hdr=rows[h]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
for r in rows[h+1:][-6:]: print("   %-70s %8.1f us" % (r[ki][:70], float(r[vi].replace(',',''))/1e3))
PY
  done
} > gpurun_out/r2_scene_cnn_microbench.txt 2>&1
timeout -s KILL 100 python tools/mma_rate_n.py > gpurun_out/r2_mma_cost_vs_n.txt 2>&1
timeout -s KILL 150 ncu --set full --clock-control none --import-source on -k regex:conv5_tc_kernel --launch-skip 5 -c 1 -f -o gpurun_out/ncu_r2_conv5_l3 python tools/bench_scene_cnn.py > gpurun_out/ncu_r2_conv5_l3.log 2>&1; echo "ncu conv5 rc=$?"
python - <<'PY'
import json
for f in ("r2_final_bench","r2_cfg3_bench","r2_cfg5_bench","r2_final_bench_ref"):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, 'value %.0f ms %.3f e2e %s train %s'%(d['value'], d['ms_per_step'], d.get('e2e',{}).get('value'), (d.get('train_step') or {}).get('ms_per_step')))
        for k in (d.get('kernels') or [])[:9]: print("   %-40s %8.3f ms frac %.3f" % (k['kernel'],k['ms_per_step'],k['frac']))
    except Exception as e: print(f,'ERR',e)
PY
cat gpurun_out/r2_scene_cnn_microbench.txt
