"""GPU parity: spline sd functions (sdfdType > 0 -> setCApproxIAK3D / setCApproxIAK3D2, fitIAK3D.R:1919-2233) through
the C ABI against the oracle.  Tolerances as elsewhere: 1e-10 on covariance entries, 1e-8 on likelihood / kriging."""
import numpy as np
import pytest

import iak3d_b200 as iak
from iak3d_b200 import synth
from oracle import iak3d_oracle as orc
from helpers import PARBNDS, SPLINE_CASES, SPLINE_KNOTS, relerr, scaled_err, spline_case_pars

pytestmark = pytest.mark.gpu
TOL_COV, TOL = 1e-10, 1e-8


def _setups(ctx, ds, types):
    sm = iak.setupIAK3D(ds["xData"], ds["dIData"], ctx=ctx, sdfdType_cd1=types[0], sdfdType_cxd0=types[1], sdfdType_cxd1=types[2],
                        sdfdKnots=SPLINE_KNOTS)
    return sm, orc.setupIAK3D(ds["xData"], ds["dIData"], types, SPLINE_KNOTS)


@pytest.mark.parametrize("label,kind,types,coefs", SPLINE_CASES)
@pytest.mark.parametrize("n,nud", [(1, 0.5), (65, 1.5), (640, 2.5)])
def test_setCApprox_vs_oracle(ctx, label, kind, types, coefs, n, nud):
    ds = synth.make_dataset(n, 300 + n)
    P, modelx, types, cmeOpt = spline_case_pars(kind, types, coefs, nud)
    sm, osm = _setups(ctx, ds, types)
    got = iak.setCIAK3D(P, modelx, *types, cmeOpt, sm)
    Cw, s2w = orc.setCIAK3D(P, modelx, *types, cmeOpt, osm)
    assert relerr(got["C"], Cw) <= TOL_COV
    assert np.array_equal(got["C"], got["C"].T)
    assert relerr(got["sigma2Vec"], s2w) <= 1e-13


@pytest.mark.parametrize("label,kind,types,coefs", SPLINE_CASES)
def test_setCApprox2_vs_oracle(ctx, label, kind, types, coefs):
    ds = synth.make_dataset(400, 41)
    P, modelx, types, cmeOpt = spline_case_pars(kind, types, coefs, 1.5)
    rng = np.random.default_rng(3)
    x1 = np.vstack([ds["xData"][:30], rng.uniform(0, 100, (60, 2))])
    d1 = np.vstack([ds["dIData"][5:35], np.tile(synth.GSM_INTERVALS, (10, 1))])
    sm, _ = _setups(ctx, ds, types)
    sm2 = iak.setupIAK3D2(x1, d1, None, None, ctx=ctx, setup2=sm)
    got = iak.setCIAK3D2(P, modelx, *types, cmeOpt, sm2)
    want = orc.setCIAK3D2(P, modelx, *types, cmeOpt, orc.setupIAK3D2(x1, d1, ds["xData"], ds["dIData"], types, SPLINE_KNOTS))
    nz = np.abs(want) > 1e-200
    assert relerr(got[nz], want[nz]) <= TOL_COV
    assert np.all(got[~nz] == 0.0)


def test_negative_averaged_sd_gives_na(ctx):
    ds = synth.make_dataset(90, 4)
    P, modelx, types, cmeOpt = spline_case_pars("matern0.5", (2, 3, 1), dict(sdfdPars_cd1=[0.5, -3.0], sdfdPars_cxd0=[1, 1, 1], sdfdPars_cxd1=[1]))
    sm, osm = _setups(ctx, ds, types)
    assert orc.setCIAK3D(P, modelx, *types, cmeOpt, osm)[0] is None
    assert iak.setCIAK3D(P, modelx, *types, cmeOpt, sm)["C"] is None
    nll = iak.nllIAK3D(None, ds["zData"], ds["XData"], modelx=modelx, nud=0.5, sdfdType_cd1=2, sdfdType_cxd0=3, sdfdType_cxd1=1,
                       cmeOpt=cmeOpt, setupMats=sm, parBnds=PARBNDS, parsBTfmd=P)
    assert nll == 9e99                                   # badnll (fitIAK3D.R:685,885)


def test_missing_basis_is_an_error(ctx):
    ds = synth.make_dataset(50, 4)
    P, modelx, types, cmeOpt = spline_case_pars("matern0.5", (2, 3, 1), SPLINE_CASES[0][3])
    sm = iak.setupIAK3D(ds["xData"], ds["dIData"], ctx=ctx)        # no knots given: the library must refuse, not guess
    with pytest.raises(iak.Iak3dError):
        iak.setCIAK3D(P, modelx, *types, cmeOpt, sm)


@pytest.mark.parametrize("label,kind,types,coefs", SPLINE_CASES)
def test_fit_and_predict_with_spline_sd(ctx, label, kind, types, coefs):
    ds = synth.make_dataset(500, 19)
    P, modelx, types, cmeOpt = spline_case_pars(kind, types, coefs)
    sm, osm = _setups(ctx, ds, types)
    got = iak.nllIAK3D(None, ds["zData"], ds["XData"], modelx=modelx, nud=0.5, sdfdType_cd1=types[0], sdfdType_cxd0=types[1],
                       sdfdType_cxd1=types[2], cmeOpt=cmeOpt, setupMats=sm, parBnds=PARBNDS, rtnAll=True, attachBigMats=True, parsBTfmd=P)
    want = orc.nllIAK3D(None, ds["zData"], ds["XData"], modelx, 0.5, *types, cmeOpt, True, osm, PARBNDS, rtnAll=True, parsBTfmd=P)
    want.update(xData=ds["xData"], dIData=ds["dIData"], sdfdKnots=SPLINE_KNOTS)
    assert abs(got["nll"] - want["nll"]) <= TOL * abs(want["nll"])
    assert scaled_err(got["betahat"], want["betahat"]) <= TOL
    assert relerr(got["sigma2Vec"], want["sigma2Vec"]) <= TOL
    Cg, _ = got["fit"].get_big(want_C=True, want_iC=False)
    assert scaled_err(Cg, want["C"]) <= 1e-9
    rng = np.random.default_rng(2)
    xMap = np.vstack([rng.uniform(0, 100, (150, 2)), ds["xData"][::60]])
    dIMap = synth.GSM_INTERVALS
    Xfn = lambda i, idx: synth.grid_design(xMap[idx], dIMap[i], 3)
    pg = iak.predictIAK3D(xMap, dIMap, Xfn, got)
    zw, vw, _, _ = orc.predictIAK3D(xMap, dIMap, Xfn, want, modelx, types, cmeOpt)
    assert scaled_err(pg["zMap"], zw) <= TOL
    assert scaled_err(pg["vMap"], vw) <= TOL


def test_fd_gradient_batch_with_spline_sd(ctx):
    """The npar+1 evaluations of a forward-difference gradient share the stationary table; each has its own scaled tables."""
    ds = synth.make_dataset(450, 23)
    types = (2, 3, 1)
    sm, osm = _setups(ctx, ds, types)
    # ax, nux, ad, cx0, cx1, cd1, cxd1(logit), 2 + 3 + 1 spline coefficients, cme
    pars = np.array([0.3, -1.0, 0.1, -2.0, -1.2, -1.8, 0.4, 0.8, 0.55, 1.1, 0.7, 0.4, 0.6, -3.0])
    kw = dict(modelx="matern", nud=0.5, sdfdType_cd1=2, sdfdType_cxd0=3, sdfdType_cxd1=1, cmeOpt=1, prodSum=True, setupMats=sm, parBnds=PARBNDS)
    g, f0 = iak.gradnllIAK3D(pars, ds["zData"], ds["XData"], rtn_nll=True, **kw)
    okw = lambda v: orc.nllIAK3D(v, ds["zData"], ds["XData"], "matern", 0.5, 2, 3, 1, 1, True, osm, PARBNDS)
    w0 = okw(pars)
    assert abs(f0 - w0) <= TOL * abs(w0)
    for i in (0, 2, 7, 10, 12, 13):
        v = pars.copy(); v[i] += 1e-6
        assert abs(g[i] - (okw(v) - w0) / 1e-6) <= 2e-3 * max(1.0, abs(g[i]))     # differences of 1e-8-accurate values / 1e-6
    singles = [iak.nllIAK3D(v, ds["zData"], ds["XData"], **kw) for v in (pars,)]
    assert singles[0] == f0                                  # batched == single, bit for bit


@pytest.mark.parametrize("optn,useReml", [(1, True), (2, False)])
def test_composite_likelihood_with_spline_sd_functions(ctx, optn, useReml):
    """setupIAK3D_CL carries the sd-function types and knots to every unit (iaCovMatern.R:551-575); the block predictors
    (compositeLikIAK3D.R:471-798) rebuild their joined-support setups with them."""
    label, kind, types, coefs = SPLINE_CASES[0]
    ds = synth.make_dataset(900, 83)
    P, modelx, types, cmeOpt = spline_case_pars(kind, types, coefs, 0.5)
    blk = synth.make_blocks(ds["xData"], ds["prof"], 4, 83)
    blk["compLikOptn"] = optn
    cl = iak.setupIAK3D_CL(ds["xData"], ds["dIData"], ds["zData"], ds["XData"], blk, ctx=ctx, sdfdType_cd1=types[0], sdfdType_cxd0=types[1],
                           sdfdType_cxd1=types[2], sdfdKnots=SPLINE_KNOTS)
    got = iak.nllIAK3D_CL(None, ds["zData"], ds["XData"], modelx, 0.5, *types, cmeOpt, True, PARBNDS, useReml, cl, parsBTfmd=P, rtnAll=True)
    want = orc.nllIAK3D_CL(None, ds["zData"], ds["XData"], modelx, 0.5, *types, cmeOpt, True, PARBNDS, useReml, blk, ds["xData"],
                           ds["dIData"], parsBTfmd=P, rtnTerms=True, sdfdKnots=SPLINE_KNOTS)
    assert abs(got["nll"] - want["nll"]) <= TOL * abs(want["nll"])
    want["sdfdKnots"] = SPLINE_KNOTS
    rng = np.random.default_rng(6)
    xMap = rng.uniform(0, 100, (90, 2))
    dIMap = synth.GSM_INTERVALS[[0, 4]]
    Xfn = lambda i, idx: synth.grid_design(xMap[idx], dIMap[i], 3)
    pg = iak.predictIAK3D_CL(xMap, dIMap, Xfn, got)
    zw, vw, lw, uw = orc.predictIAK3D_CL(xMap, dIMap, Xfn, want, modelx, types, cmeOpt)
    assert scaled_err(pg["zMap"], zw) <= TOL
    assert scaled_err(pg["vMap"], vw) <= TOL
