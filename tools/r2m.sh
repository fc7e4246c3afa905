#!/bin/bash
# fused social kernel (pooling on the tensor core, A operand in tensor memory): parity, timing experiments, timeline of block 0
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_social_fc.py tests/test_gpu_selftest.py -x -q > gpurun_out/m_social.log 2>&1; echo "social rc $?"; tail -12 gpurun_out/m_social.log
for d in 0 2 4; do
  echo -n "dbg=$d  "; DESIRE_SOCIAL_DBG=$d timeout 120 python tools/bench_social.py 2>&1 | tail -1
done
echo -n "B8 "; timeout 120 python tools/bench_social.py 8 60 20 128 2>&1 | tail -1
echo -n "v1 "; DESIRE_SOCIAL_V1=1 timeout 120 python tools/bench_social.py 2>&1 | tail -1
DESIRE_SOCIAL_TRACE=1 timeout 120 python tools/bench_social.py 8 60 20 128 > gpurun_out/m_trace.log 2>&1
head -45 gpurun_out/m_trace.log
