mkdir -p gpurun_out/r2j
(IAK3D_ENGINE=2 timeout 200 python scripts/task_profile.py 5000 9) > gpurun_out/r2j/taskprof_pair.txt 2>&1
cat gpurun_out/r2j/taskprof_pair.txt
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/r2j/bench.json 2> gpurun_out/r2j/bench.err; tail -c 300 gpurun_out/r2j/bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tile_pair -s 3 -c 1 -o gpurun_out/r2j/pair python scripts/engine_compare.py 5000 > gpurun_out/r2j/ncu.log 2>&1; tail -3 gpurun_out/r2j/ncu.log
timeout 300 python -m pytest tests/test_gpu_nr.py -x -q 2>&1 | tail -3
