# composite-likelihood scaling with both exchange kinds: bash scripts/gpu_scale_cl.sh N   (under gpurun --gpus N)
N=$1
mkdir -p gpurun_out/r2s
for kind in peer nccl; do
  IAK3D_ALLREDUCE=$kind python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 --workload cl > gpurun_out/r2s/cl_${kind}_N$N.json 2> gpurun_out/r2s/cl_${kind}_N$N.err
  tail -c 300 gpurun_out/r2s/cl_${kind}_N$N.err
  python - <<PY
import json
for l in open("gpurun_out/r2s/cl_${kind}_N$N.json"):
    if l.startswith("{"):
        d = json.loads(l); print("$kind", d["value"], d["allreduce"], d["gradient"]["ms"])
PY
done
