# last run of the round: parity suite, smoke, default bench
mkdir -p gpurun_out/r2ad
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -2 | tee gpurun_out/r2ad/gputests.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/r2ad/bench.json 2> gpurun_out/r2ad/bench.err; tail -c 300 gpurun_out/r2ad/bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2ad/bench.json').read().splitlines() if l.startswith('{')][-1])
print('S5', d['value'], d['e2e']['value'], d['single_eval_ms'], d['roofline']['frac'], d['gpu_launches'])
print('edgeroi', d['edgeroi_sized']['nll_vector_call_wall_ms'], d['edgeroi_sized']['gradient_call_wall_ms'], d['edgeroi_sized']['device_ms'])
for k in ('cl200','p1m'): print(k, d[k]['value'], d[k]['roofline']['frac'], d[k]['e2e']['value'], d[k]['gpu_launches'], d[k].get('stage_ms'), d[k]['ms_per_step'])
print('cl gradient', d['cl200']['gradient'])
PY
