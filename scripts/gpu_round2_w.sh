# final 1-GPU records of the round: default bench line (S5 + cl200 + p1m), the reference arm, smoke, launch list
mkdir -p gpurun_out/r2w
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2w/smoke.txt 2>&1; tail -3 gpurun_out/r2w/smoke.txt
python bench.py > gpurun_out/r2w/bench.json 2> gpurun_out/r2w/bench.err; tail -c 300 gpurun_out/r2w/bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2w/bench.json').read().splitlines() if l.startswith('{')][-1])
print('S5', d['value'], d['e2e']['value'], d['single_eval_ms'], d['stage_ms'], d['roofline']['frac'], d['roofline_cov_build']['frac'], d['gpu_launches'])
print('edgeroi', d.get('edgeroi_sized'))
for k in ('cl200','p1m'): print(k, d[k]['value'], d[k]['roofline']['frac'], d[k]['e2e']['value'])
print('cpu', d.get('cpu_baseline'))
PY
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2w/bench_ref.json 2> gpurun_out/r2w/bench_ref.err; tail -c 600 gpurun_out/r2w/bench_ref.json
python examples/eg_edgeroi_shaped.py --impl gpu > gpurun_out/r2w/eg_gpu.json 2> gpurun_out/r2w/eg_gpu.err; python examples/eg_edgeroi_shaped.py --impl gpu > gpurun_out/r2w/eg_gpu2.json 2>> gpurun_out/r2w/eg_gpu.err; cut -c1-400 gpurun_out/r2w/eg_gpu2.json
