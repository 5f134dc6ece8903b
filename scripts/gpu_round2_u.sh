mkdir -p gpurun_out/r2u
timeout 20 scripts/micro/potf2_bench_np 2>&1 | cut -c1-200 | tee gpurun_out/r2u/potf2_bench.txt
