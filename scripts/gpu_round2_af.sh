mkdir -p gpurun_out/r2af
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 --workload s5 --no-cpu-baseline > gpurun_out/r2af/bench_N2_s5.json 2> gpurun_out/r2af/bench_N2_s5.err
tail -c 300 gpurun_out/r2af/bench_N2_s5.err; grep "^{" gpurun_out/r2af/bench_N2_s5.json | tail -1 | cut -c1-400
