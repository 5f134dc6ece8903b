mkdir -p gpurun_out/r2t
python scripts/profile_fit.py 1200 > gpurun_out/r2t/profile_fit.txt 2>&1
head -c 7000 gpurun_out/r2t/profile_fit.txt
