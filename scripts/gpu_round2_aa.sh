# batched table launches in multi-problem evaluations: parity (CL tests use > 16 problems) and the CL200 line with / without
mkdir -p gpurun_out/r2aa
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -2 | tee gpurun_out/r2aa/gputests.txt
for tq in 0 1; do
  IAK3D_TABLE_QUEUE=$tq python bench.py --workload cl --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2aa/cl_tq$tq.json 2> gpurun_out/r2aa/cl_tq$tq.err; tail -c 200 gpurun_out/r2aa/cl_tq$tq.err
  python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2aa/cl_tq$tq.json").read().splitlines() if l.startswith("{")][-1])
print("table queue $tq:", d["value"], d["stage_ms"], d["gpu_launches"], d["e2e"]["value"])
PY
done
IAK3D_TABLE_QUEUE=0 python scripts/stage_probe.py 4000 36 2>&1 | head -1
IAK3D_TABLE_QUEUE=1 python scripts/stage_probe.py 4000 36 2>&1 | head -1
