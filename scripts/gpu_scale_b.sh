# the driver's scaling command on N GPUs with the final build: bash scripts/gpu_scale_b.sh N   (under gpurun --gpus N)
N=$1
mkdir -p gpurun_out/r2s
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2s/bench_N$N.json 2> gpurun_out/r2s/bench_N$N.err
tail -c 300 gpurun_out/r2s/bench_N$N.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2s/bench_N$N.json").read().splitlines() if l.startswith("{")][-1])
print("S5", d["value"], "cl200", d["cl200"]["value"], d["cl200"]["allreduce"], "p1m", d["p1m"]["value"], d["p1m"]["e2e"]["value"])
PY
timeout 600 python -m pytest tests/test_gpu_cl_sharded.py tests/test_gpu_peer.py -x -q 2>&1 | tail -2
