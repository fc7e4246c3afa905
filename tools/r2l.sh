#!/bin/bash
# round 2: fused social kernel, second design (A operand in tensor memory): self-test, parity, A/B timing
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_selftest.py -x -q -s > gpurun_out/l_selftest.log 2>&1; echo "selftest rc $?"
tail -8 gpurun_out/l_selftest.log
timeout 300 python -m pytest tests/test_gpu_social_fc.py -x -q -s > gpurun_out/l_social.log 2>&1; echo "social rc $?"
tail -15 gpurun_out/l_social.log
timeout 120 python tools/bench_social.py 2>&1 | tail -2
DESIRE_SOCIAL_V1=1 timeout 120 python tools/bench_social.py 2>&1 | tail -2
timeout 120 python tools/bench_social.py 8 60 20 128 2>&1 | tail -1
DESIRE_SOCIAL_V1=1 timeout 120 python tools/bench_social.py 8 60 20 128 2>&1 | tail -1
