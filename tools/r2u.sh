timeout -s KILL 300 python -m pytest tests/test_gpu_scene_cnn.py -x -q 2>&1 | tail -6
DESIRE_CONV5_TRACE=1 timeout -s KILL 60 python tools/bench_scene_cnn.py 2>&1 | grep "space-to-depth" | head -1
timeout -s KILL 60 python tools/bench_scene_cnn.py 2>&1 | tail -1
