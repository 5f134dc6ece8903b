# multi-GPU runs of the round: bash scripts/gpu_scale.sh N   (under gpurun --gpus N)
N=$1
mkdir -p gpurun_out/r2s
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2s/bench_N$N.json 2> gpurun_out/r2s/bench_N$N.err
tail -c 400 gpurun_out/r2s/bench_N$N.err
NR_N=${2:-20000}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 scripts/nr_shard.py $NR_N > gpurun_out/r2s/nr_shard_N${N}_n$NR_N.json 2> gpurun_out/r2s/nr_shard_N${N}.err
tail -c 300 gpurun_out/r2s/nr_shard_N${N}.err; cat gpurun_out/r2s/nr_shard_N${N}_n$NR_N.json | cut -c1-400
timeout 600 python -m pytest tests/test_gpu_cl_sharded.py -x -q 2>&1 | tail -2
if [ "$3" = "base" ]; then
  python scripts/nr_shard.py $NR_N > gpurun_out/r2s/nr_shard_N1_n$NR_N.json 2>> gpurun_out/r2s/nr_shard_N${N}.err; cat gpurun_out/r2s/nr_shard_N1_n$NR_N.json | cut -c1-300
fi
