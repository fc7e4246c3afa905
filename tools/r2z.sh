#!/bin/bash
# round 2, final evidence: ncu --set full captures (one launch each), launch list, bench lines (cfg2 + reference arm, cfg3,
# cfg5), MMA issue-rate probe, social kernel timeline
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
cap() {  # tag skip regex [driver args...]
  local tag=$1 skip=$2 rx=$3; shift 3
  timeout 150 $NCU -k regex:$rx --launch-skip $skip -c 1 -f -o gpurun_out/ncu_$tag python tools/profile_kernels.py --passes 1 --serial "$@" > gpurun_out/ncu_$tag.log 2>&1
  echo "ncu $tag rc=$?"
}
cap r2_gru_dec1 2 gru_tc3_kernel
cap r2_gru_dec2 6 gru_tc3_kernel
cap r2_social_fc 3 social_fc_ts_kernel
cap r2_deconv3 1 deconv_tc_kernel
cap r2_gather 0 scene_gather_kernel
cap r2_readout 0 readout_pool_kernel
timeout 150 $NCU -k regex:gru_tc3_kernel -s 2 -c 1 -f -o gpurun_out/ncu_r2_gru_dec1_cfg3 python tools/bench_gru.py --rows 75776 --hidden 256 --steps 12 --iters 2 > gpurun_out/ncu_r2_gru_dec1_cfg3.log 2>&1; echo "ncu cfg3 gru rc=$?"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_launches_fwd.csv python tools/profile_kernels.py --passes 1 --ioc-iters 2 > /dev/null 2>&1; echo "launch list rc=$?"
timeout 100 python tools/mma_rate.py > gpurun_out/r2_mma_issue_rate.txt 2>&1
DESIRE_SOCIAL_TRACE=1 timeout 100 python tools/bench_social.py 8 60 20 128 > gpurun_out/r2_social_ts_trace.txt 2>&1
(timeout 100 python tools/bench_social.py; DESIRE_SOCIAL_V1=1 timeout 100 python tools/bench_social.py) 2>&1 | grep launches > gpurun_out/r2_social_microbench.txt
(for cfg in "38400 128" "327680 256"; do set -- $cfg; timeout 60 python tools/bench_gru.py --rows $1 --hidden $2 --steps 12 2>&1 | tail -1; done) > gpurun_out/r2_gru_microbench.txt
timeout 400 python bench.py --steps 20 --warmup 5 --breakdown gpurun_out/r2_final_breakdown.json > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err; echo "bench cfg2 rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_final_bench_ref.json 2>/dev/null; echo "bench ref rc=$?"
timeout 400 python bench.py --config cfg3 --steps 5 --warmup 3 --no-train --no-cpu-baseline --breakdown gpurun_out/r2_cfg3_breakdown.json > gpurun_out/r2_cfg3_bench.json 2> gpurun_out/r2_cfg3_bench.err; echo "bench cfg3 rc=$?"
timeout 400 python bench.py --config cfg5 --steps 5 --warmup 3 --no-train --no-cpu-baseline --breakdown gpurun_out/r2_cfg5_breakdown.json > gpurun_out/r2_cfg5_bench.json 2> gpurun_out/r2_cfg5_bench.err; echo "bench cfg5 rc=$?"
python - <<'PY'
import json
for f in ("r2_final_bench","r2_cfg3_bench","r2_cfg5_bench","r2_final_bench_ref"):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, 'value %.0f ms %.3f e2e %s'%(d['value'], d['ms_per_step'], d.get('e2e',{}).get('value')))
        for k in (d.get('kernels') or [])[:8]: print("   %-40s %8.3f ms frac %.3f" % (k['kernel'],k['ms_per_step'],k['frac']))
    except Exception as e: print(f,'ERR',e)
PY
cat gpurun_out/r2_social_microbench.txt gpurun_out/r2_gru_microbench.txt
