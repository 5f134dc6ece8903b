"""Kriging throughput probe: m supports against an n-interval kept fit (device-resident timing)."""
import sys, time, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import iak3d_b200 as iak
from iak3d_b200 import synth
from helpers import PARBNDS
ctx = iak.Context(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
ds = synth.make_dataset(n, synth.SEEDS["S5"])
P, modelx, types, cmeOpt = synth.default_pars("edgeroi")
sm = iak.setupIAK3D(ds["xData"], ds["dIData"], ctx=ctx)
fit = iak.nllIAK3D(None, ds["zData"], ds["XData"], modelx=modelx, nud=0.5, sdfdType_cd1=types[0], sdfdType_cxd0=types[1],
                   sdfdType_cxd1=types[2], cmeOpt=cmeOpt, setupMats=sm, parBnds=PARBNDS, rtnAll=True, attachBigMats=True, parsBTfmd=P)
side = int(np.ceil(np.sqrt(m)))
xMap = synth.make_grid(side)[:m]
dI = np.repeat(synth.GSM_INTERVALS[3:4], m, axis=0)
X = synth.grid_design(xMap, synth.GSM_INTERVALS[3], 3)
k = fit["fit"]
k.stage(xMap, dI, X)
for rep in range(3):
    ctx.sync(); ctx.timer_start()
    k.enqueue()
    ms = ctx.timer_stop()
    z, v = k.fetch()
    print(f"n={n} m={m}: {ms:.2f} ms  {m/ms*1e3:.0f} preds/s  useful {m*float(n)**2/ms/1e9:.2f} TFLOP/s (n^2 flops per prediction)")
t0 = time.perf_counter(); z, v = k.predict(xMap, dI, X); dt = time.perf_counter() - t0
print(f"end-to-end predict(): {dt*1e3:.1f} ms  {m/dt:.0f} preds/s")
