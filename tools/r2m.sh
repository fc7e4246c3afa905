#!/bin/bash
# fused social kernel (pooling on the tensor core, A operand in tensor memory): parity (repeated: races), timing, timeline
mkdir -p gpurun_out
for i in 1 2 3; do timeout 300 python -m pytest tests/test_gpu_social_fc.py tests/test_gpu_selftest.py -x -q 2>&1 | tail -1; done
for d in 0 4; do
  echo -n "dbg=$d  "; DESIRE_SOCIAL_DBG=$d timeout 120 python tools/bench_social.py 2>&1 | tail -1
done
echo -n "B8 "; timeout 120 python tools/bench_social.py 8 60 20 128 2>&1 | tail -1
DESIRE_SOCIAL_TRACE=1 timeout 120 python tools/bench_social.py 8 60 20 128 > gpurun_out/m_trace.log 2>&1
head -5 gpurun_out/m_trace.log; tail -4 gpurun_out/m_trace.log
