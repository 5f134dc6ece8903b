"""Where the wall time of the Edgeroi-shaped fit goes (examples/eg_edgeroi_shaped.py --impl gpu): cProfile of the optimiser loop,
and the wall / device time of one nllIAK3D call at the fitted parameters.  python scripts/profile_fit.py [n]"""
import cProfile
import io
import os
import pstats
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "examples"))
import eg_edgeroi_shaped as eg  # noqa: E402
import iak3d_b200 as iak  # noqa: E402
from iak3d_b200 import synth  # noqa: E402
from iak3d_b200.api import logitab  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1200
ds = synth.make_dataset(n + 250, synth.SEEDS["E"])
fitr = ds["prof"] < ds["prof"].max() - 59
x, dI, z, X = ds["xData"][fitr], ds["dIData"][fitr], ds["zData"][fitr], ds["XData"][fitr]
sm = iak.setupIAK3D(x, dI)
kw = dict(setupMats=sm, parBnds=eg.PARBNDS, useReml=True, **eg.MODEL)
fn = lambda p: iak.nllIAK3D(p, z, X, **kw)
gr = lambda p: iak.gradnllIAK3D(p, z, X, **kw)
B = eg.PARBNDS
p0 = np.array([float(logitab(30.0, *B["axBnds"])), float(logitab(0.3, *B["adBnds"])), np.log(0.2), np.log(0.2), float(logitab(0.5, *B["sxd1Bnds"])), np.log(0.1)])
lower = np.array([-20.0] * 6); upper = np.array([20.0, 20.0, np.inf, np.inf, 20.0, np.inf])
fn(p0); gr(p0)
t0 = time.perf_counter()
pfit, nll = eg.optimIt(p0, fn, gr, lower, upper)
print("fit seconds (no profiler):", time.perf_counter() - t0, "nll", nll)
for name, f in (("nllIAK3D", fn), ("gradnllIAK3D", gr)):
    t0 = time.perf_counter()
    for _ in range(200):
        f(pfit)
    w = (time.perf_counter() - t0) / 200
    ctx = sm.ctx
    ctx.sync(); ctx.timer_start()
    for _ in range(200):
        f(pfit)
    dev = ctx.timer_stop() / 200
    print(f"{name}: wall {1e3 * w:.3f} ms per call; stream-time between calls {dev:.3f} ms")
pr = cProfile.Profile()
pr.enable()
eg.optimIt(p0, fn, gr, lower, upper)
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28)
print(s.getvalue()[:6000])
