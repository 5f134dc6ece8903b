"""GPU parity: kept fit + block kriging and composite likelihood through the C ABI against the oracle.
Tolerance (north_star): <= 1e-8 relative on predictions, kriging variances and log-likelihood."""
import numpy as np
import pytest

import iak3d_b200 as iak
from iak3d_b200 import synth
from oracle import iak3d_oracle as orc
from helpers import PARBNDS, case_pars, relerr, scaled_err

pytestmark = pytest.mark.gpu
TOL = 1e-8


def _fit_both(ctx, ds, kind, nonstat, nud, cme=None):
    P, modelx, types, cmeOpt = case_pars(kind, nonstat, nud)
    if cme is not None:
        P = dict(P); P["cme"] = cme; cmeOpt = 1 if cme > 0 else 0
    sm = iak.setupIAK3D(ds["xData"], ds["dIData"], ctx=ctx)
    got = iak.nllIAK3D(None, ds["zData"], ds["XData"], modelx=modelx, nud=nud, sdfdType_cd1=types[0], sdfdType_cxd0=types[1],
                       sdfdType_cxd1=types[2], cmeOpt=cmeOpt, setupMats=sm, parBnds=PARBNDS, rtnAll=True, attachBigMats=True,
                       parsBTfmd=P)
    want = orc.nllIAK3D(None, ds["zData"], ds["XData"], modelx, nud, *types, cmeOpt, True,
                        orc.setupIAK3D(ds["xData"], ds["dIData"]), PARBNDS, rtnAll=True, parsBTfmd=P)
    want["xData"], want["dIData"] = ds["xData"], ds["dIData"]
    return got, want, modelx, types, cmeOpt


@pytest.mark.parametrize("kind,nonstat,nud", [("edgeroi", False, 0.5), ("matern0.5", False, 0.5), ("matern0.8", True, 1.5),
                                              ("nugget", False, 0.5), ("wendland", False, 2.5)])
def test_rtnAll_lists_match(ctx, kind, nonstat, nud):
    ds = synth.make_dataset(600, 8)
    got, want, *_ = _fit_both(ctx, ds, kind, nonstat, nud)
    assert abs(got["nll"] - want["nll"]) <= TOL * abs(want["nll"])
    assert abs(got["cxdhat"] - want["cxdhat"]) <= TOL * abs(want["cxdhat"])
    assert scaled_err(got["betahat"], want["betahat"]) <= TOL
    assert scaled_err(got["vbetahat"], want["vbetahat"]) <= TOL
    assert scaled_err(got["iCX"], want["iCX"]) <= 1e-7          # explicit-inverse products of an ill-conditioned C
    assert scaled_err(got["iCz_muhat"], want["iCz_muhat"]) <= 1e-7
    assert relerr(got["sigma2Vec"], want["sigma2Vec"]) <= TOL
    for k in ("cx0", "cx1", "cd1", "cxd0", "cxd1", "cme"):
        assert abs(got["parsBTfmd"][k] - want["parsBTfmd"][k]) <= TOL * abs(want["parsBTfmd"][k]) + 1e-300
    Cg, _ = got["fit"].get_big(want_C=True, want_iC=False)
    assert scaled_err(Cg, want["C"]) <= 1e-9


@pytest.mark.parametrize("kind,nonstat,nud", [("edgeroi", False, 0.5), ("matern0.5", True, 0.5), ("matern0.8", False, 2.5),
                                              ("nugget", False, 1.5)])
def test_predict_vs_oracle(ctx, kind, nonstat, nud):
    ds = synth.make_dataset(700, 15)
    got, want, modelx, types, cmeOpt = _fit_both(ctx, ds, kind, nonstat, nud)
    rng = np.random.default_rng(2)
    # map: random cells + cells exactly at data locations + the "distant" point (fitIAK3D.R:606)
    xMap = np.vstack([rng.uniform(0, 100, (300, 2)), ds["xData"][::50], [[9e99, 9e99]]])
    dIMap = synth.GSM_INTERVALS
    Xfn = lambda i, idx: synth.grid_design(np.clip(xMap[idx], 0, 100), dIMap[i], 3)
    pg = iak.predictIAK3D(xMap, dIMap, Xfn, got)
    zw, vw, lw, uw = orc.predictIAK3D(xMap, dIMap, Xfn, want, modelx, types, cmeOpt)
    assert scaled_err(pg["zMap"], zw) <= TOL
    assert scaled_err(pg["vMap"], vw) <= TOL
    assert scaled_err(pg["pi90LMap"], lw) <= TOL and scaled_err(pg["pi90UMap"], uw) <= TOL


def test_kriging_is_exact_at_data_supports_without_measurement_error(ctx):
    ds = synth.make_dataset(500, 21)
    got, want, modelx, types, cmeOpt = _fit_both(ctx, ds, "matern0.5", False, 0.5, cme=0.0)
    sel = np.arange(0, 500, 7)
    z, v = got["fit"].predict(ds["xData"][sel], ds["dIData"][sel], ds["XData"][sel])
    assert np.max(np.abs(z - ds["zData"][sel])) <= 1e-7 * np.max(np.abs(ds["zData"]))
    assert np.max(np.abs(v)) <= 1e-6 * got["cxdhat"]


def test_predict_chunking_and_nan_rows(ctx):
    """More supports than two panels (2 x 148 row blocks of 128 = 37888 rows each on a B200; the third reuses the first
    panel's buffer while the second is being solved) and NaN design rows (predictIAK3D.R:160-163)."""
    ds = synth.make_dataset(300, 5)
    got, want, modelx, types, cmeOpt = _fit_both(ctx, ds, "edgeroi", False, 0.5)
    xMap = synth.make_grid(300)[:80000]
    dIMap = synth.GSM_INTERVALS[2:3]

    def Xfn(i, idx):
        X = synth.grid_design(xMap[idx], dIMap[i], 3)
        X[idx % 1000 == 7, 2] = np.nan
        return X
    pg = iak.predictIAK3D(xMap, dIMap, Xfn, got)
    sub = np.r_[0:50, 16350:16450, 37850:37950, 75700:75850, 79950:80000]
    zw, vw, _, _ = orc.predictIAK3D(xMap[sub], dIMap, lambda i, idx: Xfn(i, sub[idx]), want, modelx, types, cmeOpt)
    okw = ~np.isnan(zw[0])
    assert np.array_equal(np.isnan(pg["zMap"][0, sub]), ~okw)
    assert scaled_err(pg["zMap"][0, sub][okw], zw[0][okw]) <= TOL
    assert scaled_err(pg["vMap"][0, sub][okw], vw[0][okw]) <= TOL
    assert np.isnan(pg["zMap"][0, 7]) and np.isnan(pg["vMap"][0, 1007])


@pytest.mark.parametrize("optn,useReml", [(1, True), (2, False), (3, False), (4, True), (4, False)])
def test_composite_likelihood_vs_oracle(ctx, optn, useReml):
    ds = synth.make_dataset(1500, 61)
    P, modelx, types, cmeOpt = case_pars("edgeroi", False, 0.5)
    blk = synth.make_blocks(ds["xData"], ds["prof"], 7, 61)
    blk["compLikOptn"] = optn
    cl = iak.setupIAK3D_CL(ds["xData"], ds["dIData"], ds["zData"], ds["XData"], blk, ctx=ctx)
    got = iak.nllIAK3D_CL(None, ds["zData"], ds["XData"], modelx, 0.5, *types, cmeOpt, True, PARBNDS, useReml, cl, parsBTfmd=P,
                          rtnTerms=True)
    want = orc.nllIAK3D_CL(None, ds["zData"], ds["XData"], modelx, 0.5, *types, cmeOpt, True, PARBNDS, useReml, blk, ds["xData"],
                           ds["dIData"], parsBTfmd=P, rtnTerms=True)
    assert abs(got["nll"] - want["nll"]) <= TOL * abs(want["nll"])
    assert scaled_err(got["XziAXz_SUM"], want["XziAXz_SUM"]) <= TOL
    assert abs(got["lndetA_SUM"] - want["lndetA_SUM"]) <= TOL * abs(want["lndetA_SUM"])


def test_cl_single_block_equals_exact(ctx):
    ds = synth.make_dataset(640, 9)
    P, modelx, types, cmeOpt = case_pars("matern0.5", False, 0.5)
    sm = iak.setupIAK3D(ds["xData"], ds["dIData"], ctx=ctx)
    exact = iak.nllIAK3D(None, ds["zData"], ds["XData"], modelx=modelx, nud=0.5, cmeOpt=cmeOpt, setupMats=sm, parBnds=PARBNDS, parsBTfmd=P)
    blk = dict(compLikOptn=4, listBlocks=[np.arange(640)], subsetPairsAdj=np.zeros((0, 2), int), subsetPairsNonadj=np.zeros((0, 2), int))
    cl = iak.setupIAK3D_CL(ds["xData"], ds["dIData"], ds["zData"], ds["XData"], blk, ctx=ctx)
    got = iak.nllIAK3D_CL(None, ds["zData"], ds["XData"], modelx, 0.5, *types, cmeOpt, True, PARBNDS, True, cl, parsBTfmd=P)
    assert got == exact


@pytest.mark.parametrize("optn,useReml,kind,nonstat", [(1, True, "edgeroi", False), (4, True, "matern0.5", True), (1, False, "matern0.8", False)])
def test_cl_block_prediction_vs_oracle(ctx, optn, useReml, kind, nonstat):
    """predictIAK3D for a composite-likelihood fit (predictIAK3D.R:211-226 -> predMatsIAK3D_CL_ByBlock): every adjacent
    pair is one kept factorisation; the non-stationary case exercises the square-build cd1 quirk inside Ckh."""
    ds = synth.make_dataset(1300, 71)
    P, modelx, types, cmeOpt = case_pars(kind, nonstat, 0.5)
    blk = synth.make_blocks(ds["xData"], ds["prof"], 6, 71)
    blk["compLikOptn"] = optn
    cl = iak.setupIAK3D_CL(ds["xData"], ds["dIData"], ds["zData"], ds["XData"], blk, ctx=ctx)
    if optn == 4:      # the fit uses single blocks only; prediction still needs the adjacent pairs (compositeLikIAK3D.R:503-507)
        pr = dict(blk); pr["compLikOptn"] = 1
        clp = iak.setupIAK3D_CL(ds["xData"], ds["dIData"], ds["zData"], ds["XData"], pr, ctx=ctx)
    got = iak.nllIAK3D_CL(None, ds["zData"], ds["XData"], modelx, 0.5, *types, cmeOpt, True, PARBNDS, useReml, cl, parsBTfmd=P, rtnAll=True)
    want = orc.nllIAK3D_CL(None, ds["zData"], ds["XData"], modelx, 0.5, *types, cmeOpt, True, PARBNDS, useReml, blk, ds["xData"],
                           ds["dIData"], parsBTfmd=P, rtnTerms=True)
    assert abs(got["nll"] - want["nll"]) <= TOL * abs(want["nll"])
    assert scaled_err(got["vbetahat"], want["vbetahat"]) <= TOL
    if optn == 4:
        got["compLikMats"] = dict(got["compLikMats"]); got["compLikMats"]["units"] = clp["units"]
    rng = np.random.default_rng(4)
    xMap = np.vstack([rng.uniform(0, 100, (140, 2)), ds["xData"][::160]])
    dIMap = synth.GSM_INTERVALS[[0, 3, 5]]
    Xfn = lambda i, idx: synth.grid_design(xMap[idx], dIMap[i], 3)
    pg = iak.predictIAK3D_CL(xMap, dIMap, Xfn, got)
    zw, vw, lw, uw = orc.predictIAK3D_CL(xMap, dIMap, Xfn, want, modelx, types, cmeOpt)
    assert scaled_err(pg["zMap"], zw) <= TOL
    assert scaled_err(pg["vMap"], vw) <= TOL
    assert scaled_err(pg["pi90LMap"], lw) <= TOL


@pytest.mark.parametrize("optn,kind,nonstat,nblocks", [(2, "edgeroi", False, 6), (3, "edgeroi", False, 6), (2, "matern0.5", True, 4),
                                                       (3, "matern0.8", False, 2)])
def test_clev_block_prediction_vs_oracle(ctx, optn, kind, nonstat, nblocks):
    """predictIAK3D for a composite-likelihood fit with the Eidsvik options (predictIAK3D.R:211-217 ->
    predMatsIAK3D_CLEV_ByBlock, compositeLikIAK3D.R:609-798): joint prediction of each block's map supports from the
    pairwise conditionals, all matrices on the device."""
    ds = synth.make_dataset(1300, 73)
    P, modelx, types, cmeOpt = case_pars(kind, nonstat, 0.5)
    blk = synth.make_blocks(ds["xData"], ds["prof"], nblocks, 73)
    blk["compLikOptn"] = optn
    cl = iak.setupIAK3D_CL(ds["xData"], ds["dIData"], ds["zData"], ds["XData"], blk, ctx=ctx)
    got = iak.nllIAK3D_CL(None, ds["zData"], ds["XData"], modelx, 0.5, *types, cmeOpt, True, PARBNDS, False, cl, parsBTfmd=P, rtnAll=True)
    want = orc.nllIAK3D_CL(None, ds["zData"], ds["XData"], modelx, 0.5, *types, cmeOpt, True, PARBNDS, False, blk, ds["xData"],
                           ds["dIData"], parsBTfmd=P, rtnTerms=True)
    assert abs(got["nll"] - want["nll"]) <= TOL * abs(want["nll"])
    rng = np.random.default_rng(5)
    xMap = np.vstack([rng.uniform(0, 100, (170, 2)), ds["xData"][::200]])
    dIMap = synth.GSM_INTERVALS[[1, 4]]
    Xfn = lambda i, idx: synth.grid_design(xMap[idx], dIMap[i], 3)
    pg = iak.predictIAK3D_CL(xMap, dIMap, Xfn, got)
    zw, vw, lw, uw = orc.predictIAK3D_CL(xMap, dIMap, Xfn, want, modelx, types, cmeOpt)
    assert np.all(vw > 0)
    assert scaled_err(pg["zMap"], zw) <= TOL
    assert scaled_err(pg["vMap"], vw) <= TOL
    assert scaled_err(pg["pi90UMap"], uw) <= TOL


def test_cl_forward_difference_gradient_vs_oracle(ctx):
    """gradnllIAK3D_CL (fitIAK3D.R:1819-1911): npar + 1 composite-likelihood evaluations, NA -> 0."""
    ds = synth.make_dataset(700, 67)
    blk = synth.make_blocks(ds["xData"], ds["prof"], 4, 67)
    blk["compLikOptn"] = 1
    modelx, types, cmeOpt = "matern", (0, 0, 0), 1
    cl = iak.setupIAK3D_CL(ds["xData"], ds["dIData"], ds["zData"], ds["XData"], blk, ctx=ctx)
    p0 = np.array([0.2, -1.3, 0.1, -2.2, -1.4, -1.9, 0.4, -3.0])
    g, f0 = iak.gradnllIAK3D_CL(p0, ds["zData"], ds["XData"], modelx, 0.5, *types, cmeOpt, True, PARBNDS, True, cl, rtn_nll=True)
    fw = lambda v: orc.nllIAK3D_CL(v, ds["zData"], ds["XData"], modelx, 0.5, *types, cmeOpt, True, PARBNDS, True, blk, ds["xData"],
                                   ds["dIData"])
    w0 = fw(p0)
    assert abs(f0 - w0) <= TOL * abs(w0)
    gw = np.array([(fw(p0 + 1e-6 * np.eye(len(p0))[k]) - w0) / 1e-6 for k in range(len(p0))])
    # forward differences of values that agree to ~1e-12 relative: compare at the noise floor of the difference
    assert np.max(np.abs(g - gw)) <= 2e-2 * max(1.0, np.max(np.abs(gw)))


def test_profilePredict_vs_oracle(ctx):
    """profilePredictIAK3D (predictIAK3D.R:292-504): one location, a full profile of depth intervals, incl. Ckh and iC."""
    ds = synth.make_dataset(450, 27)
    got, want, modelx, types, cmeOpt = _fit_both(ctx, ds, "matern0.5", True, 0.5)
    x1 = np.array([[41.3, 57.9]])
    dI = np.array([[0.0, 0.05], [0.05, 0.13], [0.13, 0.27], [0.27, 0.61], [0.61, 1.04], [1.04, 1.5], [1.5, 2.0]])
    X = synth.grid_design(np.repeat(x1, len(dI), axis=0), dI[0], 3)
    for k in range(len(dI)):
        X[k] = synth.grid_design(x1, dI[k], 3)[0]
    pg = iak.profilePredictIAK3D(x1, dI, X, got, rtnBigMats=True)
    zw, vw, lw, uw = orc.predictIAK3D(x1, dI, lambda i, idx: X[i:i + 1], want, modelx, types, cmeOpt)
    assert scaled_err(pg["zMap"], zw[:, 0]) <= TOL and scaled_err(pg["vMap"], vw[:, 0]) <= TOL
    assert scaled_err(pg["pi90UMap"], uw[:, 0]) <= TOL
    sm2 = orc.setupIAK3D2(np.repeat(x1, len(dI), axis=0), dI, ds["xData"], ds["dIData"])
    assert scaled_err(pg["Ckh"], orc.setCIAK3D2(want["parsBTfmd"], modelx, *types, cmeOpt, sm2)) <= 1e-10
    assert scaled_err(pg["iC"], want["iC"]) <= 1e-7
    assert scaled_err(pg["muhatMap"], (X @ want["betahat"]).ravel()) <= TOL
