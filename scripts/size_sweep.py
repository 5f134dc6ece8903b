"""REML terms against LAPACK on the host for awkward sizes (tile / unit boundaries +-1, primes): lndet, W'A^-1W, A^-1W."""
import sys, numpy as np, scipy.linalg as sla
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import iak3d_b200 as iak
from iak3d_b200 import synth
from helpers import PARBNDS
ctx = iak.Context(0)
worst = 0.0
for n in [int(a) for a in sys.argv[1:]] or [63, 64, 65, 127, 128, 129, 191, 192, 193, 255, 257, 1000, 1531, 4999, 5001, 8191]:
    for kind in ("matern0.5", "edgeroi"):
        ds = synth.make_dataset(n, 1000 + n)
        P, modelx, types, cmeOpt = synth.default_pars(kind)
        W = np.column_stack([ds["XData"], ds["zData"]])
        sm = iak.SetupMats(ds["xData"], ds["dIData"], W=W, ctx=ctx)
        t = iak.nllIAK3D(None, ds["zData"], ds["XData"], modelx=modelx, nud=0.5, sdfdType_cd1=types[0], sdfdType_cxd0=types[1],
                         sdfdType_cxd1=types[2], cmeOpt=cmeOpt, setupMats=sm, parBnds=PARBNDS, forCompLik=True, attachBigMats=True, parsBTfmd=P)
        A = t["A"]
        c = sla.cho_factor(A, lower=True, check_finite=False)
        lnd = 2.0 * np.sum(np.log(np.diag(c[0])))
        iAW = sla.cho_solve(c, W, check_finite=False)
        e = [abs(t["lndetA"] - lnd) / max(abs(lnd), 1.0), np.max(np.abs(t["WiAW"] - W.T @ iAW)) / np.max(np.abs(W.T @ iAW)),
             np.max(np.abs(t["iAW"] - iAW)) / np.max(np.abs(iAW))]
        worst = max(worst, max(e))
        print(f"n={n:5d} {kind:10s} lndet {e[0]:.1e}  WiAW {e[1]:.1e}  iAW {e[2]:.1e}", flush=True)
        sm.close()
print("worst", worst)
assert worst < 1e-9
