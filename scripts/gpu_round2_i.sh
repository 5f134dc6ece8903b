mkdir -p gpurun_out/r2i
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2i/gputests.txt 2>&1; tail -4 gpurun_out/r2i/gputests.txt
(IAK3D_ENGINE=1 timeout 200 python scripts/task_profile.py 5000 9; IAK3D_ENGINE=2 timeout 200 python scripts/task_profile.py 5000 9) > gpurun_out/r2i/taskprof.txt 2>&1
cat gpurun_out/r2i/taskprof.txt
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/r2i/bench.json 2> gpurun_out/r2i/bench.err; tail -c 600 gpurun_out/r2i/bench.err
IAK3D_ENGINE=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:tile_pair -s 3 -c 1 -o gpurun_out/r2i/pair python scripts/engine_compare.py 5000 > gpurun_out/r2i/ncu.log 2>&1; tail -3 gpurun_out/r2i/ncu.log
