#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:social_fc_ts -s 3 -c 1 -f -o gpurun_out/p_social_ts python tools/bench_social.py > gpurun_out/p_ncu.log 2>&1
tail -2 gpurun_out/p_ncu.log
