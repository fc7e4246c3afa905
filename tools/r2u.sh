timeout -s KILL 200 compute-sanitizer --tool racecheck --print-limit 3 python tools/bench_scene_cnn.py 1 32 2>&1 | tail -5
