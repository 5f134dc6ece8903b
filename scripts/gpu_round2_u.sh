mkdir -p gpurun_out/r2u
timeout 120 scripts/micro/potf2_bench 2>&1 | tee gpurun_out/r2u/potf2_bench.txt
