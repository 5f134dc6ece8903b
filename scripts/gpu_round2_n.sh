mkdir -p gpurun_out/r2n
timeout 600 python -m pytest tests/test_gpu_ssre.py -x -q > gpurun_out/r2n/ssre.txt 2>&1; tail -15 gpurun_out/r2n/ssre.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2n/gputests.txt 2>&1; tail -3 gpurun_out/r2n/gputests.txt
timeout 600 python scripts/cl_overhead_probe.py 200000 8 > gpurun_out/r2n/cl_probe.txt 2>&1; head -60 gpurun_out/r2n/cl_probe.txt
