#!/bin/bash
# round 2, run D: ncu captures of the v3 recurrence, full GPU test suite, smoke, bench cfg2
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 120 $NCU -k regex:gru_tc3 -s 2 -c 1 -o gpurun_out/ncu_r2d_gru3_h128 python tools/bench_gru.py --rows 38400 --hidden 128 --steps 12 --iters 2 > gpurun_out/ncu_r2d_h128.log 2>&1; echo "ncu h128 rc=$?"
timeout 150 $NCU -k regex:gru_tc3 -s 2 -c 1 -o gpurun_out/ncu_r2d_gru3_h256 python tools/bench_gru.py --rows 75776 --hidden 256 --steps 12 --iters 2 > gpurun_out/ncu_r2d_h256.log 2>&1; echo "ncu h256 rc=$?"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_gpu_tests.log 2>&1; echo "gpu tests rc=$?"
tail -6 gpurun_out/r2d_gpu_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2d_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/r2d_smoke.log
timeout 400 python bench.py --steps 10 --warmup 3 --breakdown gpurun_out/r2d_breakdown.json > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r2d_bench.json; tail -5 gpurun_out/r2d_bench.err
