# profiling pass of the final round-2 tree (one GPU): launch list, full captures of the pair engine (factorisation and kriging
# panel), in-kernel task timeline, S40 single-GPU run
mkdir -p gpurun_out/r2m
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload s5 > gpurun_out/r2m/bench_s5_plain.json 2> gpurun_out/r2m/bench_s5_plain.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2m/launches_s5.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload s5 > gpurun_out/r2m/ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tile_pair -s 3 -c 1 -o gpurun_out/r2m/pair_mode0 python scripts/engine_compare.py 5000 > gpurun_out/r2m/ncu_pair0.log 2>&1; tail -2 gpurun_out/r2m/ncu_pair0.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cov_factor_kernel -s 3 -c 1 -o gpurun_out/r2m/covfactor python scripts/engine_compare.py 5000 > gpurun_out/r2m/ncu_cov.log 2>&1; tail -2 gpurun_out/r2m/ncu_cov.log
timeout 600 ncu --set full --clock-control none -k regex:tile_pair -s 2 -c 1 -o gpurun_out/r2m/pair_mode1 python scripts/krige_profile.py 5000 37888 > gpurun_out/r2m/ncu_pair1.log 2>&1; tail -2 gpurun_out/r2m/ncu_pair1.log
(timeout 200 python scripts/task_profile.py 5000 9; timeout 200 python scripts/task_profile.py 5000 1; timeout 100 python scripts/task_profile.py 1200 1) > gpurun_out/r2m/taskprof.txt 2>&1
timeout 200 python scripts/krige_profile.py 5000 37888 > gpurun_out/r2m/krige_taskprof.txt 2>&1
timeout 900 python scripts/big_runs.py 40000 nr > gpurun_out/r2m/s40.log 2>&1; tail -4 gpurun_out/r2m/s40.log
