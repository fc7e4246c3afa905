timeout -s KILL 300 python -m pytest tests/test_gpu_scene_cnn.py -x -q 2>&1 | tail -5
timeout -s KILL 60 python tools/bench_scene_cnn.py
DESIRE_NO_CONV5=1 timeout -s KILL 60 python tools/bench_scene_cnn.py
timeout -s KILL 100 ncu --metrics gpu__time_duration.sum --clock-control none -c 30 --csv --log-file gpurun_out/c5_launches.csv python tools/bench_scene_cnn.py > /dev/null 2>&1
tail -8 gpurun_out/c5_launches.csv | cut -d, -f5,15
