timeout -s KILL 400 python -m pytest tests/test_gpu_train.py tests/test_gpu_scene_cnn.py tests/test_gpu_parity.py -x -q 2>&1 | tail -3
timeout -s KILL 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['train_step']['ms_per_step'])"
