"""Per-stage device times of one likelihood evaluation for several models / sizes (diagnostic)."""
import sys, ctypes as C, numpy as np
sys.path.insert(0, '.')
import iak3d_b200 as iak
from iak3d_b200 import synth
from iak3d_b200._lib import make_pars, Pars, dptr, iptr
ctx = iak.Context(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 1
ds = synth.make_dataset(n, synth.SEEDS["S5"])
W = np.column_stack([ds["XData"], ds["zData"]])
sm = iak.SetupMats(ds["xData"], ds["dIData"], W=W, ctx=ctx)
q = W.shape[1]
for kind, nonstat in (("matern0.5", False), ("edgeroi", False), ("matern0.8", False), ("matern0.5", True), ("wendland", False)):
    P, modelx, types, cmeOpt = synth.default_pars(kind, nonstat)
    arr = (Pars * nb)(*[make_pars(P, modelx, *types, cmeOpt)] * nb)
    dsarr = (C.c_void_p * nb)(*[sm.h] * nb)
    lnd = np.empty(nb); G = np.empty((nb, q * q)); st = np.empty(nb, np.int32)
    acc = np.zeros(4); reps = 5
    for r in range(reps + 2):
        ctx.check(ctx.lib.iak3d_nll_terms_batch(ctx.h, nb, dsarr, arr, dptr(lnd), dptr(G), iptr(st)))
        ms, fl, by = ctx.last_stage_ms()
        if r >= 2: acc += np.array(ms)
    acc /= reps
    print(f"{kind:10s} nonstat={nonstat} n={n} nb={nb} status={st[0]} tables={acc[0]:.3f} cov={acc[1]:.3f} chol={acc[2]:.3f} fin={acc[3]:.3f} ms"
          f"  chol TF={nb*n**3/3/acc[2]/1e9:.2f} cov GB/s={nb*8*n*(n+1)/2/acc[1]/1e6:.0f}")
