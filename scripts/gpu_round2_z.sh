# final checks of the round on one GPU: parity suite (default engine choice and pair engine everywhere), smoke, default bench, launch list
mkdir -p gpurun_out/r2z
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/r2z/gputests.txt
IAK3D_ENGINE=3 timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -2 | tee gpurun_out/r2z/gputests_engine3.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/r2z/bench.json 2> gpurun_out/r2z/bench.err; tail -c 300 gpurun_out/r2z/bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2z/bench.json').read().splitlines() if l.startswith('{')][-1])
print('S5', d['value'], d['e2e']['value'], d['single_eval_ms'], d['gradient_wall_ms'], d['roofline']['frac'])
print('edgeroi', d['edgeroi_sized']['nll_call_wall_ms'], d['edgeroi_sized']['device_ms'])
for k in ('cl200','p1m'): print(k, d[k]['value'], d[k]['roofline']['frac'], d[k]['e2e']['value'])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2z/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload s5 > gpurun_out/r2z/ncu_bench.log 2>&1; wc -l gpurun_out/r2z/launches.csv
python examples/eg_edgeroi_shaped.py --impl gpu > gpurun_out/r2z/eg_gpu.json 2> gpurun_out/r2z/eg_gpu.err; cut -c1-330 gpurun_out/r2z/eg_gpu.json
