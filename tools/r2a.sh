#!/bin/bash
# round 2, run A: GRU v3 correctness + micro-benchmarks (v3 vs v2)
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_gru.py -x -q > gpurun_out/r2a_gru_tests.log 2>&1; echo "gru tests rc=$?"
tail -15 gpurun_out/r2a_gru_tests.log
for cfg in "38400 128" "327680 256"; do set -- $cfg
  timeout 120 python tools/bench_gru.py --rows $1 --hidden $2 --steps 12 2>&1 | tail -1
  DESIRE_GRU_V2=1 timeout 120 python tools/bench_gru.py --rows $1 --hidden $2 --steps 12 2>&1 | tail -1
done | tee gpurun_out/r2a_bench_gru.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_gpu_tests.log 2>&1; echo "gpu tests rc=$?"
tail -5 gpurun_out/r2a_gpu_tests.log
