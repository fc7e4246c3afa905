timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r2_final_gputests.txt; cat gpurun_out/r2_final_gputests.txt
bash tools/r2z.sh
