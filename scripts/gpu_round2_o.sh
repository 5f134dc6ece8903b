mkdir -p gpurun_out/r2o
timeout 600 python -m pytest tests/test_gpu_design.py -x -q > gpurun_out/r2o/design.txt 2>&1; tail -15 gpurun_out/r2o/design.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2o/smoke.txt 2>&1; tail -3 gpurun_out/r2o/smoke.txt
