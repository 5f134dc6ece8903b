mkdir -p gpurun_out/r2ab
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -2 | tee gpurun_out/r2ab/gputests.txt
python bench.py --workload cl --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2ab/cl.json 2> gpurun_out/r2ab/cl.err; tail -c 200 gpurun_out/r2ab/cl.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2ab/cl.json").read().splitlines() if l.startswith("{")][-1])
print("cl200:", d["value"], d["stage_ms"], d["gpu_launches"], d["e2e"]["value"])
PY
python scripts/stage_probe.py 1200 1 2>&1 | sed -n 2p
