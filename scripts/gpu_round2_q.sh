mkdir -p gpurun_out/r2q
IAK3D_PAIR_WARPS=8 timeout 200 python scripts/task_profile.py 5000 9 > gpurun_out/r2q/taskprof_w8.txt 2>&1; cat gpurun_out/r2q/taskprof_w8.txt | head -12
IAK3D_PAIR_WARPS=8 timeout 300 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,launch__block_size --clock-control none -k regex:tile_ -s 3 -c 2 python scripts/engine_compare.py 5000 2>&1 | grep -E "tile_|duration|tensor|registers|block_size" | head -12
