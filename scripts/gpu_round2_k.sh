mkdir -p gpurun_out/r2k
(IAK3D_ENGINE=1 timeout 120 python scripts/engine_compare.py 2049 5000; timeout 120 python scripts/engine_compare.py 2049 5000 5100) > gpurun_out/r2k/cmp.txt 2>&1; cat gpurun_out/r2k/cmp.txt
IAK3D_ENGINE=3 timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2k/gputests_e3.txt 2>&1; tail -3 gpurun_out/r2k/gputests_e3.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2k/gputests.txt 2>&1; tail -3 gpurun_out/r2k/gputests.txt
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/r2k/bench.json 2> gpurun_out/r2k/bench.err; tail -c 300 gpurun_out/r2k/bench.err
