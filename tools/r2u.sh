python -m pytest tests/test_gpu_social_fc.py -x -q 2>&1 | tail -2
for d in 0 512 32 2; do echo "DBG=$d"; DESIRE_SOCIAL_DBG=$d timeout 120 python tools/bench_social.py 2>&1 | tail -1; done > gpurun_out/elim6.txt 2>&1
DESIRE_SOCIAL_TRACE=1 timeout 120 python tools/bench_social.py > gpurun_out/ts_trace5.txt 2>&1
cat gpurun_out/elim6.txt
