timeout -s KILL 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r2_final_gputests.txt; cat gpurun_out/r2_final_gputests.txt
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout -s KILL 300 python bench.py --steps 20 --warmup 5 --breakdown gpurun_out/r2_final_breakdown.json > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/r2_final_bench.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['train_step']['ms_per_step'], d['gpu_launches'], d['roofline']['frac'], d['clocks'])"
