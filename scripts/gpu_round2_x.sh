# ncu capture of a single evaluation's factorisation kernel (tile_task_kernel<-1,1>: warp-specialised diagonal blocks) at n = 5000
mkdir -p gpurun_out/r2x
python scripts/stage_probe.py 5000 1 2>&1 | head -1
ncu --set full --clock-control none --import-source on -k regex:tile_task_kernel -s 4 -c 1 -o gpurun_out/r2x/single python scripts/stage_probe.py 5000 1 > gpurun_out/r2x/ncu.log 2>&1; tail -3 gpurun_out/r2x/ncu.log
python scripts/ncu_summary.py gpurun_out/r2x/single.ncu-rep gpurun_out/r2x/ncu_tile_task_single_r02.txt "tile_task_kernel<-1,1>: one factorisation at n = 5000 (single-problem launch: 79 chained diagonal blocks, potf2_inv_chain)"
ncu -i gpurun_out/r2x/single.ncu-rep --page source --csv > gpurun_out/r2x/single_source.csv 2>/dev/null; wc -l gpurun_out/r2x/single_source.csv
python scripts/task_profile.py 1200 1 > gpurun_out/r2x/task_1200.txt 2>&1; python scripts/task_profile.py 5000 1 > gpurun_out/r2x/task_5000.txt 2>&1; head -6 gpurun_out/r2x/task_5000.txt
