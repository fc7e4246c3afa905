timeout -s KILL 300 python -m pytest tests/test_gpu_scene_cnn.py -x -q 2>&1 | tail -3
timeout -s KILL 60 python tools/bench_scene_cnn.py 2>&1 | tail -1
timeout -s KILL 100 ncu --metrics gpu__time_duration.sum --clock-control none -c 30 --csv --log-file gpurun_out/c5_launches.csv python tools/bench_scene_cnn.py > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/c5_launches.csv')))
h=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
hdr=rows[h]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
for r in rows[h+1:][-6:]: print(r[ki][:60], r[vi])
PY
