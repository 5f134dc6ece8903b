mkdir -p gpurun_out/r2t
python scripts/stage_probe.py 1200 1 > gpurun_out/r2t/stage_1200.txt 2>&1
python scripts/task_profile.py 1200 1 > gpurun_out/r2t/task_1200.txt 2>&1
python scripts/task_profile.py 5000 1 > gpurun_out/r2t/task_5000.txt 2>&1
cat gpurun_out/r2t/stage_1200.txt gpurun_out/r2t/task_1200.txt gpurun_out/r2t/task_5000.txt
