"""A/B of the two tile engines (IAK3D_ENGINE=1: one CTA per SM; 2: CTA pairs, chol_pair.cuh) on batched likelihood
evaluations: run twice, once per engine, and compare the printed values / times."""
import sys, os, ctypes as C, numpy as np
sys.path.insert(0, '.')
import iak3d_b200 as iak
from iak3d_b200 import synth
from iak3d_b200._lib import make_pars, Pars, dptr, iptr
ctx = iak.Context(0)
sizes = [int(a) for a in sys.argv[1:]] or [700, 2049, 5000]
for n in sizes:
    ds = synth.make_dataset(n, synth.SEEDS["S5"])
    W = np.column_stack([ds["XData"], ds["zData"]])
    sm = iak.SetupMats(ds["xData"], ds["dIData"], W=W, ctx=ctx)
    q = W.shape[1]
    nb = 9
    P, modelx, types, cmeOpt = synth.default_pars("matern0.8", False)
    plist = []
    for k in range(nb):
        Pk = dict(P); Pk["ax"] = P["ax"] * (1 + 0.01 * k)
        plist.append(make_pars(Pk, modelx, *types, cmeOpt))
    arr = (Pars * nb)(*plist)
    dsarr = (C.c_void_p * nb)(*[sm.h] * nb)
    lnd = np.empty(nb); G = np.empty((nb, q * q)); st = np.empty(nb, np.int32)
    acc = np.zeros(4); reps = 10
    for r in range(reps + 3):
        ctx.check(ctx.lib.iak3d_nll_terms_batch(ctx.h, nb, dsarr, arr, dptr(lnd), dptr(G), iptr(st)))
        ms, fl, by = ctx.last_stage_ms()
        if r >= 3: acc += np.array(ms)
    acc /= reps
    print(f"engine={os.environ.get('IAK3D_ENGINE','default')} n={n} nb={nb} status={st.tolist()} chol={acc[2]:.3f} ms TF={nb*n**3/3/acc[2]/1e9:.2f}"
          f" lnd0={lnd[0]!r} lnd8={lnd[8]!r} G0sum={G[0].sum()!r} G8sum={G[8].sum()!r}", flush=True)
ctx.close()
