"""Explicit inverse (chol2inv path of lndetANDinvCb / lmmFit$iC) timing and accuracy at n = 5000."""
import sys, time, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import iak3d_b200 as iak
from iak3d_b200 import synth
from helpers import PARBNDS, case_pars
ctx = iak.Context(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
ds = synth.make_dataset(n, 8)
P, modelx, types, cmeOpt = case_pars("edgeroi", False, 0.5)
sm = iak.setupIAK3D(ds["xData"], ds["dIData"], ctx=ctx)
fit = iak.nllIAK3D(None, ds["zData"], ds["XData"], modelx=modelx, nud=0.5, sdfdType_cd1=types[0], sdfdType_cxd0=types[1], sdfdType_cxd1=types[2],
                   cmeOpt=cmeOpt, setupMats=sm, parBnds=PARBNDS, rtnAll=True, attachBigMats=True, parsBTfmd=P)
for rep in range(3):
    t0 = time.perf_counter(); Cm, iC = fit["fit"].get_big(want_C=False, want_iC=True); dt = time.perf_counter() - t0
    print(f"n={n}: iC (device inverse + {n * n * 8 / 1e6:.0f} MB to host) {dt * 1e3:.1f} ms")
Cm, _ = fit["fit"].get_big(want_C=True, want_iC=False)
E = Cm @ iC - np.eye(n)
print("max |C iC - I| =", np.abs(E).max(), " symmetric:", np.array_equal(iC, iC.T))
th = np.log([max(P[k], 1e-3) for k in ("cx0", "cx1", "cd1", "cxd0", "cxd1", "cme")])
W = np.column_stack([ds["XData"], ds["zData"]]); sm.set_W(W)
t0 = time.perf_counter(); g = iak.gradHessIAK3D(sm, P, modelx, types, th); print(f"gradHessIAK3D {time.perf_counter() - t0:.3f} s nll={g['nll']:.6f}")
