#!/usr/bin/env python
"""The flow of the reference's example driver (egEdgeroi.R: fitIAK3D -> predictIAK3D -> validation) on an Edgeroi-SHAPED
synthetic dataset (BASELINE.json configs[0]; the Edgeroi data themselves come from R packages that are not available
offline): ~290 profiles, ~1200 horizon intervals for fitting, 60 hold-out profiles (getEdgeroiData.R:87), the Edgeroi
covariance settings (egEdgeroi.R:83-97: modelx = 'spherical', sdfdType_cd1 = -9, prodSum, cmeOpt = 1, nud = 0.5).

    python examples/eg_edgeroi_shaped.py [--impl gpu|oracle] [--n 1200] [--grid 60]

The optimiser loop is optimIt (optim_functions.R:177-257): L-BFGS-B (maxit 70, gradient = the forward-difference batch
gradnllIAK3D) and Nelder-Mead (350 evaluations) in turn until a round improves the objective by <= 1e-4.  `--impl oracle`
runs the same loop on the CPU oracle (test infrastructure; slow) so that the two fits can be compared.  One JSON line."""
import argparse
import json
import os
import sys
import time

import numpy as np
from scipy.optimize import minimize

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from iak3d_b200 import synth  # noqa: E402

PARBNDS = dict(axBnds=(0.05, 60.0), nuxBnds=(0.05, 20.0), adBnds=(0.02, 1.2), tau2Bnds=(0.01, 20.0), sxd1Bnds=(1e-8, 1.0 - 1e-8), maxd=2.0)
MODEL = dict(modelx="spherical", nud=0.5, sdfdType_cd1=-9, sdfdType_cxd0=0, sdfdType_cxd1=0, cmeOpt=1, prodSum=True)


def optimIt(par, fn, gr, lower, upper, tolIt=1e-4, log=None):
    """optim_functions.R:177-257."""
    prevOF = fn(par)
    it, bounds = 1, list(zip(lower, upper))
    while True:
        if it % 2 == 1:
            res = minimize(fn, par, jac=gr, method="L-BFGS-B", bounds=bounds, options=dict(maxiter=70))
        else:
            res = minimize(fn, par, method="Nelder-Mead", bounds=bounds, options=dict(maxfev=350, xatol=1e-8, fatol=1e-8))
        if log is not None:
            log.append(dict(method="L-BFGS-B" if it % 2 == 1 else "Nelder-Mead", value=float(res.fun), nfev=int(res.nfev)))
        it += 1
        if res.fun < prevOF - tolIt:
            prevOF, par = float(res.fun), res.x
        else:
            return par, prevOF


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", default="gpu", choices=["gpu", "oracle"])
    ap.add_argument("--n", type=int, default=1200)
    ap.add_argument("--grid", type=int, default=60)
    ap.add_argument("--pars-out", default=None)
    args = ap.parse_args()
    ds = synth.make_dataset(args.n + 250, synth.SEEDS["E"])
    hold = ds["prof"] >= ds["prof"].max() - 59                      # the last 60 profiles are held out
    fitr = ~hold
    x, dI, z, X = ds["xData"][fitr], ds["dIData"][fitr], ds["zData"][fitr], ds["XData"][fitr]
    nev = dict(fn=0, gr=0)
    t_setup = time.perf_counter()
    if args.impl == "gpu":
        import iak3d_b200 as iak
        sm = iak.setupIAK3D(x, dI)                                  # creates the CUDA context on first use (~1 s on a fresh box)
        kw = dict(setupMats=sm, parBnds=PARBNDS, useReml=True, **MODEL)

        def fn(p):
            nev["fn"] += 1
            return iak.nllIAK3D(p, z, X, **{k: v for k, v in kw.items()})

        def gr(p):
            nev["gr"] += 1
            return iak.gradnllIAK3D(p, z, X, **kw)
    else:
        from oracle import iak3d_oracle as orc
        osm = orc.setupIAK3D(x, dI)
        a = (MODEL["modelx"], 0.5, -9, 0, 0, 1, True, osm, PARBNDS)

        def fn(p):
            nev["fn"] += 1
            return orc.nllIAK3D(p, z, X, *a)

        def gr(p):
            nev["gr"] += 1
            f0 = orc.nllIAK3D(p, z, X, *a)
            g = np.zeros(p.size)
            for i in range(p.size):
                v = p.copy(); v[i] += 1e-6
                g[i] = (orc.nllIAK3D(v, z, X, *a) - f0) / 1e-6
            g[~np.isfinite(g)] = 0.0
            return g
    # initial values in the spirit of setInitsIAK3D: ax a third of the extent, ad 0.3 m, equal variance shares
    from iak3d_b200.api import logitab
    p0 = np.array([float(logitab(30.0, *PARBNDS["axBnds"])), float(logitab(0.3, *PARBNDS["adBnds"])), np.log(0.2), np.log(0.2),
                   float(logitab(0.5, *PARBNDS["sxd1Bnds"])), np.log(0.1)])
    lower = np.array([-20.0] * 6); upper = np.array([20.0, 20.0, np.inf, np.inf, 20.0, np.inf])   # setInitsIAK3D.R:753-943
    log = []
    fn(p0)                                                          # first call: library load, task lists
    t0 = time.perf_counter()
    t_setup = t0 - t_setup
    pfit, nll = optimIt(p0, fn, gr, lower, upper, log=log)
    t_fit = time.perf_counter() - t0
    out = dict(impl=args.impl, setup_seconds=t_setup, n=int(fitr.sum()), nprofiles=int(len(np.unique(ds["prof"][fitr]))), nll=float(nll), pars=list(map(float, pfit)),
               fn_evals=nev["fn"], gr_evals=nev["gr"], fit_seconds=t_fit, rounds=log)
    # ---- prediction at the GlobalSoilMap depths on a grid and at the hold-out supports (egEdgeroi.R:396-430) ----
    t0 = time.perf_counter()
    xMap = synth.make_grid(args.grid); dIMap = synth.GSM_INTERVALS
    Xfn = lambda i, idx: synth.grid_design(xMap[idx], dIMap[i], synth.SEEDS["E"])
    if args.impl == "gpu":
        fit = iak.nllIAK3D(pfit, z, X, rtnAll=True, attachBigMats=True, **kw)
        pm = iak.predictIAK3D(xMap, dIMap, Xfn, fit)
        zv, vv = fit["fit"].predict(ds["xData"][hold], ds["dIData"][hold], ds["XData"][hold])
        zMap, vMap = pm["zMap"], pm["vMap"]
    else:
        fit = orc.nllIAK3D(pfit, z, X, *a, rtnAll=True)
        fit.update(xData=x, dIData=dI)
        types = (-9, 0, 0)
        zMap, vMap, _, _ = orc.predictIAK3D(xMap, dIMap, Xfn, fit, "spherical", types, 1)
        zv = np.empty(int(hold.sum())); vv = np.empty_like(zv)
        hx, hd, hX = ds["xData"][hold], ds["dIData"][hold], ds["XData"][hold]
        for k in range(zv.size):                                   # one support at a time keeps the oracle's memory small
            a_, b_, _, _ = orc.predictIAK3D(hx[k:k + 1], hd[k:k + 1], lambda i, idx: hX[k:k + 1], fit, "spherical", types, 1)
            zv[k], vv[k] = a_[0, 0], b_[0, 0]
    res = ds["zData"][hold] - zv
    out.update(predict_seconds=time.perf_counter() - t0, map_mean=float(np.mean(zMap)), map_var_mean=float(np.mean(vMap)),
               holdout_rmse=float(np.sqrt(np.mean(res ** 2))), holdout_bias=float(np.mean(res)),
               holdout_in_pi90=float(np.mean(np.abs(res) <= 1.645 * np.sqrt(vv))), cxdhat=float(fit["cxdhat"]),
               betahat=[float(b) for b in np.ravel(fit["betahat"])])
    if args.pars_out:
        np.save(args.pars_out, pfit)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
