DESIRE_SOCIAL_DBG=512 DESIRE_SOCIAL_TRACE=1 timeout 60 python tools/bench_social.py > gpurun_out/ts_trace7.txt 2>&1
