#!/bin/bash
# round 2, run B: GRU v3 after the xp-ring parity fix; timing experiments; ncu captures
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gru.py -x -q --timeout 60 > gpurun_out/r2b_gru_tests.log 2>&1; echo "gru tests rc=$?"
tail -8 gpurun_out/r2b_gru_tests.log
for cfg in "38400 128" "327680 256"; do set -- $cfg
  for dbg in 0 1 2 3; do
    echo -n "dbg=$dbg "; DESIRE_GRU3_DBG=$dbg timeout 120 python tools/bench_gru.py --rows $1 --hidden $2 --steps 12 2>&1 | tail -1
  done
  echo -n "v2 "; DESIRE_GRU_V2=1 timeout 120 python tools/bench_gru.py --rows $1 --hidden $2 --steps 12 2>&1 | tail -1
done | tee gpurun_out/r2b_bench_gru.log
# one-wave shapes (148 tiles): what a tile costs without the wave tail
for cfg in "18944 128" "18944 256"; do set -- $cfg
  echo -n "1wave "; timeout 120 python tools/bench_gru.py --rows $1 --hidden $2 --steps 12 2>&1 | tail -1
done | tee -a gpurun_out/r2b_bench_gru.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gru_tc3 -s 2 -c 1 -o gpurun_out/ncu_r2b_gru3_h128 python tools/bench_gru.py --rows 38400 --hidden 128 --steps 12 --iters 2 > gpurun_out/ncu_r2b_h128.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gru_tc3 -s 2 -c 1 -o gpurun_out/ncu_r2b_gru3_h256 python tools/bench_gru.py --rows 75776 --hidden 256 --steps 12 --iters 2 > gpurun_out/ncu_r2b_h256.log 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/r2b_gpu_tests.log 2>&1; echo "gpu tests rc=$?"
tail -5 gpurun_out/r2b_gpu_tests.log
