"""Wall-clock anatomy of one nllIAK3D call (the optimiser's fn) at Edgeroi size and at n = 5000: device stages vs host overhead."""
import sys, time, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import iak3d_b200 as iak
from iak3d_b200 import synth
from helpers import PARBNDS
ctx = iak.Context(0)
for n in (1200, 5000):
    ds = synth.make_dataset(n, synth.SEEDS["E"])
    sm = iak.setupIAK3D(ds["xData"], ds["dIData"], ctx=ctx)
    kw = dict(modelx="spherical", nud=0.5, sdfdType_cd1=-9, sdfdType_cxd0=0, sdfdType_cxd1=0, cmeOpt=1, prodSum=True, setupMats=sm, parBnds=PARBNDS, useReml=True)
    p = np.array([0.3, 0.1, -2.0, -1.2, 0.4, -3.0])
    for _ in range(5): iak.nllIAK3D(p, ds["zData"], ds["XData"], **kw)
    reps = 200
    t0 = time.perf_counter()
    for k in range(reps): iak.nllIAK3D(p + 1e-3 * np.sin(k), ds["zData"], ds["XData"], **kw)
    wall = (time.perf_counter() - t0) / reps * 1e3
    ms, fl, by = ctx.last_stage_ms()
    t0 = time.perf_counter()
    for k in range(reps): iak.gradnllIAK3D(p + 1e-3 * np.sin(k), ds["zData"], ds["XData"], **kw)
    gwall = (time.perf_counter() - t0) / reps * 1e3
    gms, _, _ = ctx.last_stage_ms()
    print(f"n={n}: nllIAK3D wall {wall:.3f} ms, device stages {np.round(ms, 3)} sum {sum(ms):.3f} ms; gradnllIAK3D (7 evals) wall {gwall:.3f} ms, device {sum(gms):.3f} ms")
