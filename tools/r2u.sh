timeout -s KILL 300 python -m pytest tests/test_gpu_gemm_tc.py -x -q 2>&1 | tail -3
for f in 0 1; do echo -n "NO_PERSIST=$f "; DESIRE_GEMM_NO_PERSIST=$f timeout -s KILL 200 python bench.py --config cfg3 --steps 4 --warmup 3 --no-train --no-cpu-baseline 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['e2e'].get('result_checksum'), [(k['kernel'],round(k['ms_per_step'],2)) for k in d['kernels'] if 'proj' in k['kernel']])"; done
