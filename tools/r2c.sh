#!/bin/bash
# round 2, run C: GRU v3 (per-group box barriers, Decoder-2 form) — short, every command under a tight timeout
mkdir -p gpurun_out
export DESIRE_GRU3_WATCHDOG=1
timeout 200 python -m pytest tests/test_gpu_gru.py -x -q > gpurun_out/r2c_gru_tests.log 2>&1; echo "gru tests rc=$?"
tail -12 gpurun_out/r2c_gru_tests.log
unset DESIRE_GRU3_WATCHDOG
for cfg in "38400 128" "327680 256" "18944 128" "18944 256"; do set -- $cfg
  echo -n "v3 "; timeout 60 python tools/bench_gru.py --rows $1 --hidden $2 --steps 12 2>&1 | tail -1
done | tee gpurun_out/r2c_bench_gru.log
for dbg in 1 2 3; do
  echo -n "dbg=$dbg "; DESIRE_GRU3_DBG=$dbg timeout 60 python tools/bench_gru.py --rows 38400 --hidden 128 --steps 12 2>&1 | tail -1
  echo -n "dbg=$dbg "; DESIRE_GRU3_DBG=$dbg timeout 60 python tools/bench_gru.py --rows 327680 --hidden 256 --steps 12 2>&1 | tail -1
done | tee -a gpurun_out/r2c_bench_gru.log
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_existence.py tests/test_randn.py -x -q -m gpu > gpurun_out/r2c_parity.log 2>&1; echo "parity rc=$?"
tail -15 gpurun_out/r2c_parity.log
