"""Smallest end-to-end pass over every kernel of the library (for compute-sanitizer)."""
import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import iak3d_b200 as iak
from iak3d_b200 import synth
from helpers import PARBNDS, case_pars
ctx = iak.Context(0)
ds = synth.make_dataset(300, 3)
for kind, nonstat in (("matern0.8", True), ("edgeroi", False)):
    P, modelx, types, cmeOpt = case_pars(kind, nonstat, 0.5)
    sm = iak.setupIAK3D(ds["xData"], ds["dIData"], ctx=ctx)
    kw = dict(modelx=modelx, nud=0.5, sdfdType_cd1=types[0], sdfdType_cxd0=types[1], sdfdType_cxd1=types[2], cmeOpt=cmeOpt,
              setupMats=sm, parBnds=PARBNDS, parsBTfmd=P)
    fit = iak.nllIAK3D(None, ds["zData"], ds["XData"], rtnAll=True, attachBigMats=True, **kw)
    t = iak.nllIAK3D(None, ds["zData"], ds["XData"], forCompLik=True, attachBigMats=True, **kw)
    xMap = synth.make_grid(12); dIMap = synth.GSM_INTERVALS[:2]
    pr = iak.predictIAK3D(xMap, dIMap, lambda i, idx: synth.grid_design(xMap[idx], dIMap[i], 1), fit)
    Cm, iC = fit["fit"].get_big()
    g = iak.gradHessIAK3D(sm, P, modelx, types, np.log([max(P[k], 1e-3) for k in ("cx0", "cx1", "cd1", "cxd0", "cxd1", "cme")]))
    sm2 = iak.setupIAK3D2(xMap[:50], np.repeat(dIMap[:1], 50, 0), ds["xData"], ds["dIData"], ctx=ctx)
    X2 = iak.setCIAK3D2(P, modelx, *types, cmeOpt, sm2)
    print(kind, fit["nll"], float(np.nanmean(pr["zMap"])), g["nll"], X2.shape, float(np.abs(Cm @ iC - np.eye(300)).max()))
A = np.cov(np.random.default_rng(0).standard_normal((90, 400))) + np.eye(90)
print(iak.lndetANDinvCb(A, np.ones((90, 3)), ctx=ctx)["lndetC"], iak.lndetANDinvCb(A, None, ctx=ctx)["invCb"].shape)
print("fp64 peaks skipped; launches", ctx.launch_count())
ctx.close()
