mkdir -p gpurun_out/r2p
(IAK3D_PAIR_WARPS=4 timeout 120 python scripts/engine_compare.py 5000; IAK3D_PAIR_WARPS=8 timeout 120 python scripts/engine_compare.py 2049 5000 5100) > gpurun_out/r2p/cmp.txt 2>&1; cat gpurun_out/r2p/cmp.txt
IAK3D_PAIR_WARPS=8 IAK3D_ENGINE=3 timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2p/gputests_w8_e3.txt 2>&1; tail -3 gpurun_out/r2p/gputests_w8_e3.txt
IAK3D_PAIR_WARPS=8 timeout 900 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2p/bench_w8.json 2> gpurun_out/r2p/bench_w8.err; tail -c 300 gpurun_out/r2p/bench_w8.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2p/bench_w8.json').read().splitlines() if l.startswith('{')][-1])
print('S5', d['value'], d['stage_ms']['factorise'], d['roofline']['frac'])
for k in ('cl200','p1m'): print(k, d[k]['value'], d[k]['roofline']['frac'])
PY
