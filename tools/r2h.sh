#!/bin/bash
# round 2, run H (2 GPUs): NCCL gradient-equality test, strong + weak scaling lines at N=2
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_dist_nccl.py tests/test_gpu_parity.py -x -q -m gpu -k "allreduced or long_horizon or fallback" -s > gpurun_out/r2h_tests.log 2>&1; echo "tests rc=$?"
grep -E "rel-L2|passed|failed|skipped|Error" gpurun_out/r2h_tests.log | tail -12
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --scaling strong --no-cpu-baseline > gpurun_out/r2h_bench_strong2.json 2> gpurun_out/r2h_bench_strong2.err; echo "strong2 rc=$?"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2h_bench_weak2.json 2> gpurun_out/r2h_bench_weak2.err; echo "weak2 rc=$?"
python - <<'PY'
import json
for f in ("r2h_bench_strong2","r2h_bench_weak2"):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, d['scaling'], d['n_gpus'], 'value %.0f e2e %.0f ms %.3f train %s'%(d['value'], d['e2e']['value'], d['ms_per_step'], d['train_step']))
    except Exception as e: print(f, 'ERR', e)
PY
tail -3 gpurun_out/r2h_bench_strong2.err
