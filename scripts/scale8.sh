# N = 8 (or $1) GPUs: the three workloads back to back, one JSON line each into gpurun_out/scale_N*.json
N=${1:-8}
for wl in "s5:--steps 30 --warmup 4 --no-cpu-baseline" "cl:--workload cl --steps 3 --warmup 3" "krige:--workload krige --steps 2 --warmup 2"; do
  name=${wl%%:*}; args=${wl#*:}
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N $args 2> gpurun_out/scale_${name}_N$N.err | tail -1 > gpurun_out/scale_${name}_N$N.json
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/scale_${name}_N$N.json"))
    print("$name", {k: d.get(k) for k in ("metric", "value", "unit", "n_gpus", "ms_per_step", "scaling")}, "frac", round(d["roofline"]["frac"], 3), "e2e", d["e2e"]["value"])
except Exception as e:
    print("$name failed", e)
PY
  tail -2 gpurun_out/scale_${name}_N$N.err | cut -c1-300
done
