mkdir -p gpurun_out/r2ae
python bench.py > gpurun_out/r2ae/bench.json 2> gpurun_out/r2ae/bench.err; tail -c 400 gpurun_out/r2ae/bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2ae/bench.json').read().splitlines() if l.startswith('{')][-1])
print('S5', d['value'], d['ms_per_step'], d['stage_ms'], d['e2e'], d['roofline']['frac'], d['gpu_launches'])
for k in ('cl200','p1m'): print(k, d[k]['value'], d[k]['e2e']['value'])
PY
