"""Forward-difference batch vs analytic gradient (iak3d_nll_grad): wall time per gradient and agreement."""
import sys, time, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import iak3d_b200 as iak
from iak3d_b200 import synth
from helpers import PARBNDS
ctx = iak.Context(0)
for n in [int(a) for a in sys.argv[1:]] or [1200, 5000]:
    ds = synth.make_dataset(n, synth.SEEDS["S5"])
    sm = iak.setupIAK3D(ds["xData"], ds["dIData"], ctx=ctx)
    kw = dict(modelx="matern", nud=0.5, cmeOpt=1, prodSum=True, setupMats=sm, parBnds=PARBNDS, useReml=True)
    p = np.array([0.2, -1.3, 0.1, -2.2, -1.4, -1.9, 0.4, -3.0])
    out = {}
    for name, an in (("fd", False), ("analytic", True)):
        for _ in range(2): g = iak.gradnllIAK3D(p, ds["zData"], ds["XData"], analytic=an, **kw)
        reps = 10 if n <= 5000 else 2
        ctx.sync(); t0 = time.perf_counter()
        for k in range(reps): g = iak.gradnllIAK3D(p + 1e-4 * np.sin(k), ds["zData"], ds["XData"], analytic=an, **kw)
        ctx.sync(); out[name] = ((time.perf_counter() - t0) / reps * 1e3, g)
    print(f"n={n}: forward-difference batch (9 evaluations) {out['fd'][0]:.2f} ms | analytic {out['analytic'][0]:.2f} ms | speed-up {out['fd'][0] / out['analytic'][0]:.2f}x | max |diff| / max |g| = "
          f"{np.max(np.abs(out['fd'][1] - out['analytic'][1])) / np.max(np.abs(out['analytic'][1])):.2e}", flush=True)
