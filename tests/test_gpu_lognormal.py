"""GPU parity: lognormal data (lnTfmdData = TRUE; nrUpdatesIAK3DlnN.R, fitIAK3D.R:762-781, predictIAK3D.R:229-284)
through the C ABI -- one bordered factorisation with W = [X, z, sigma2Vec - diag(C), V_ab] -- against the oracle's
literal restatement with the explicit inverse.  Tolerance 1e-8 (north_star)."""
import numpy as np
import pytest

import iak3d_b200 as iak
from iak3d_b200 import synth
from oracle import iak3d_oracle as orc
from helpers import PARBNDS, case_pars, lognormal_case, relerr, scaled_err

pytestmark = pytest.mark.gpu
TOL = 1e-8


def _both(ctx, n, seed, pU, kind, nonstat, nud, rtnAll=True):
    ds, z, iU, vXU = lognormal_case(n, seed, pU=max(pU, 1))
    if pU == 0:
        iU, vXU = [], None
    P, modelx, types, cmeOpt = case_pars(kind, nonstat, nud)
    sm = iak.setupIAK3D(ds["xData"], ds["dIData"], ctx=ctx)
    got = iak.nllIAK3D(None, z, ds["XData"], vXU=vXU, iU=iU, modelx=modelx, nud=nud, sdfdType_cd1=types[0], sdfdType_cxd0=types[1],
                       sdfdType_cxd1=types[2], cmeOpt=cmeOpt, setupMats=sm, parBnds=PARBNDS, useReml=False, lnTfmdData=True,
                       rtnAll=rtnAll, attachBigMats=True, parsBTfmd=P)
    want = orc.nllIAK3D_lnN(None, z, ds["XData"], vXU, iU, modelx, nud, *types, cmeOpt, True,
                            orc.setupIAK3D(ds["xData"], ds["dIData"]), PARBNDS, rtnAll=rtnAll, parsBTfmd=P)
    return ds, z, iU, vXU, got, want, modelx, types, cmeOpt


@pytest.mark.parametrize("pU", [0, 1, 3])
@pytest.mark.parametrize("kind,nonstat,nud", [("matern0.5", False, 0.5), ("edgeroi", True, 1.5), ("matern0.8", True, 2.5)])
def test_lognormal_nll_and_fit_lists(ctx, pU, kind, nonstat, nud):
    ds, z, iU, vXU, got, want, *_ = _both(ctx, 450, 12, pU, kind, nonstat, nud)
    assert abs(got["nll_unrounded"] - want["nll_unrounded"]) <= TOL * abs(want["nll_unrounded"])
    assert abs(got["nll"] - want["nll"]) <= 1.01e-5                       # both are rounded to the 1e-5 grid (:160-161)
    assert got["noits"] == want["noits"]
    assert scaled_err(got["betahat"], want["betahat"]) <= TOL
    assert scaled_err(got["vbetahat"], want["vbetahat"]) <= 1e-7
    assert got["cxdhat"] == 1.0
    assert relerr(got["sigma2Vec"], want["sigma2Vec"]) <= 1e-12
    assert relerr(got["diagC"], np.diag(want["C"])) <= 1e-12
    assert scaled_err(got["iCX"], want["iCX"]) <= 1e-7
    assert scaled_err(got["iCz_muhat"], want["iCz_muhat"]) <= 1e-7


def test_lognormal_scalar_return_and_repeatability(ctx):
    ds, z, iU, vXU, got, want, *_ = _both(ctx, 300, 3, 2, "matern0.5", False, 0.5, rtnAll=False)
    assert isinstance(got, float) and abs(got - want) <= 1.01e-5
    again = _both(ctx, 300, 3, 2, "matern0.5", False, 0.5, rtnAll=False)[4]
    assert again == got


def test_lognormal_bad_parameters_give_badnll(ctx):
    ds, z, iU, vXU = lognormal_case(120, 4, pU=2)
    P, modelx, types, cmeOpt = case_pars("matern0.5", False, 0.5)
    P = dict(P); P["nux"] = 99.0
    sm = iak.setupIAK3D(ds["xData"], ds["dIData"], ctx=ctx)
    assert iak.nllIAK3D(None, z, ds["XData"], vXU=vXU, iU=iU, modelx=modelx, nud=0.5, cmeOpt=cmeOpt, setupMats=sm, parBnds=PARBNDS,
                        useReml=False, lnTfmdData=True, parsBTfmd=P) == 9e99


@pytest.mark.parametrize("pU,backtransform", [(0, True), (2, True), (2, False)])
def test_lognormal_predictions(ctx, pU, backtransform):
    ds, z, iU, vXU, got, want, modelx, types, cmeOpt = _both(ctx, 500, 30, pU, "edgeroi", False, 0.5)
    want.update(xData=ds["xData"], dIData=ds["dIData"])
    rng = np.random.default_rng(2)
    xMap = np.vstack([rng.uniform(0, 100, (120, 2)), ds["xData"][::70]])
    dIMap = synth.GSM_INTERVALS
    Xfn = lambda i, idx: synth.grid_design(xMap[idx], dIMap[i], 3)

    def vfn(i, idx):
        r = np.random.default_rng(100 + i)
        B = r.standard_normal((xMap.shape[0], pU, pU)) * 0.1
        return np.einsum("nij,nkj->nik", B, B)[idx].reshape(-1, pU)
    pg = iak.predictIAK3D(xMap, dIMap, Xfn, got, vXUMap_fn=vfn if pU else None, rqrBTfmdPreds=backtransform)
    zw, vw, lw, uw = orc.predictIAK3D(xMap, dIMap, Xfn, want, modelx, types, cmeOpt, vXUMap_fn=vfn if pU else None,
                                      rqrBTfmdPreds=backtransform)
    assert scaled_err(pg["zMap"], zw) <= TOL and scaled_err(pg["vMap"], vw) <= 1e-7
    assert scaled_err(pg["pi90LMap"], lw) <= 1e-7 and scaled_err(pg["pi90UMap"], uw) <= 1e-7
    if backtransform:
        assert np.all(pg["zMap"] > 0)


def test_lognormal_forward_difference_gradient(ctx):
    """gradnllIAK3D with lnTfmdData = TRUE (fitIAK3D.R:1865-1886): nothing is profiled, every candidate is a full
    evaluation with its own Newton-Raphson; values (and therefore the quotient) live on the reference's 1e-5 grid."""
    ds, z, iU, vXU = lognormal_case(300, 9, pU=2)
    p0 = np.array([0.2, -1.3, 0.1, -1.5, -1.2, -1.9, -0.8, -1.0, -2.5])
    sm = iak.setupIAK3D(ds["xData"], ds["dIData"], ctx=ctx)
    kw = dict(vXU=vXU, iU=iU, modelx="matern", nud=0.5, cmeOpt=1, prodSum=True, setupMats=sm, parBnds=PARBNDS, useReml=False, lnTfmdData=True)
    g, f0 = iak.gradnllIAK3D(p0, z, ds["XData"], rtn_nll=True, **kw)
    osm = orc.setupIAK3D(ds["xData"], ds["dIData"])
    fw = lambda v: orc.nllIAK3D_lnN(v, z, ds["XData"], vXU, iU, "matern", 0.5, 0, 0, 0, 1, True, osm, PARBNDS)
    w0 = fw(p0)
    assert abs(f0 - w0) <= 1.01e-5
    gw = np.array([(fw(p0 + 1e-6 * np.eye(len(p0))[k]) - w0) / 1e-6 for k in range(len(p0))])
    assert np.all(np.isfinite(g))
    # both quotients are differences of values rounded to 1e-5, divided by 1e-6: they can differ by one grid step (10)
    assert np.max(np.abs(g - gw)) <= 10.0 + 1e-6
