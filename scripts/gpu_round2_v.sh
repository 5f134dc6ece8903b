mkdir -p gpurun_out/r2v
python scripts/stage_probe.py 1200 1 2>&1 | head -2 | tee gpurun_out/r2v/stage_1200.txt
python scripts/task_profile.py 1200 1 2>&1 | sed -n 1,5p | tee gpurun_out/r2v/task_1200.txt
python scripts/task_profile.py 5000 1 2>&1 | sed -n 1,5p | tee gpurun_out/r2v/task_5000.txt
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r2v/gputests.txt
python scripts/profile_fit.py 1200 2>&1 | head -4 | tee gpurun_out/r2v/profile_fit.txt
