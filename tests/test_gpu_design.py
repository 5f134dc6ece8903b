"""GPU parity: makeXvX on the device (csrc/design.cu through iak3d_design_* / iak3d_predict_stage_grid) against the oracle
restatement (oracle/makexvx_oracle.py).  Tolerance: 1e-12 relative to the column's largest entry (VERDICT task 8)."""
import numpy as np
import pytest

import iak3d_b200 as iak
from iak3d_b200 import synth
from oracle import iak3d_oracle as orc
from oracle import makexvx_oracle as mx
from helpers import PARBNDS, case_pars, scaled_err
from design_cases import ALLKNOTS, cubist_data, cubist_model, intervals, mlr_case

pytestmark = pytest.mark.gpu
TOLX = 1e-12


def col_err(A, B):
    A = np.asarray(A, float); B = np.asarray(B, float)
    assert A.shape == B.shape
    assert np.array_equal(np.isnan(A), np.isnan(B))
    den = np.maximum(np.nanmax(np.abs(B), axis=0), 1e-300)
    return float(np.nanmax(np.abs(A - B) / den[None, :]))


@pytest.mark.parametrize("modelX", [0, 0.5, 1, 1.5, 2, 2.5, 3])
def test_mlr_numeric_models(ctx, modelX):
    cov, levels, dI = mlr_case(777, 21)
    want = mx.makeXvX(cov, dI, modelX, covLevels=levels)
    got = iak.makeXvX(cov, dI, modelX, covLevels=levels, ctx=ctx)
    assert got["namesX"] == want["namesX"]
    assert col_err(got["X"], want["X"]) <= TOLX
    assert np.allclose(got["XLims"], want["XLims"], rtol=1e-12, atol=0, equal_nan=True)


def test_mlr_finite_limits_follow_the_reference_recycling(ctx):
    cov, levels, dI = mlr_case(501, 22, factor=False)
    free = mx.makeXvX(cov, dI, 2)
    p = len(free["namesX"])
    rng = np.random.default_rng(4)
    L = np.vstack([np.quantile(free["X"], 0.2, axis=0) - rng.uniform(0, 1, p), np.quantile(free["X"], 0.8, axis=0) + rng.uniform(0, 1, p)])
    want = mx.makeXvX(cov, dI, 2, XLims=L)
    got = iak.makeXvX(cov, dI, 2, XLims=L, ctx=ctx)
    assert not np.array_equal(want["X"], free["X"])
    assert col_err(got["X"], want["X"]) <= TOLX
    # one -Inf / +Inf pair anywhere disables every limit (makeXvX.R:190-194)
    L2 = L.copy(); L2[0, 1] = -np.inf; L2[1, 2] = np.inf
    assert col_err(iak.makeXvX(cov, dI, 2, XLims=L2, ctx=ctx)["X"], free["X"]) <= TOLX


@pytest.mark.parametrize("nDiscPts", [10, 100])
def test_mlr_bspline_of_depth_columns(ctx, nDiscPts):
    cov, levels, dI = mlr_case(300, 23, factor=False)
    names = ["const", "elev", "d.twi"] + [f"dSpline.{k + 1}" for k in range(6)]
    want = mx.makeXvX(cov, dI, names, allKnotsd=ALLKNOTS, nDiscPts=nDiscPts)
    got = iak.makeXvX(cov, dI, names, allKnotsd=ALLKNOTS, nDiscPts=nDiscPts, ctx=ctx)
    assert col_err(got["X"], want["X"]) <= TOLX
    L = np.vstack([np.full(9, -1e9), np.full(9, 1e9)]); L[0, 3:] = 0.01; L[1, 3:] = 0.6
    want = mx.makeXvX(cov, dI, names, allKnotsd=ALLKNOTS, nDiscPts=nDiscPts, XLims=L)
    got = iak.makeXvX(cov, dI, names, allKnotsd=ALLKNOTS, nDiscPts=nDiscPts, XLims=L, ctx=ctx)
    assert col_err(got["X"], want["X"]) <= TOLX


def test_supports_beyond_the_boundary_knots_are_nan(ctx):
    cov, levels, dI = mlr_case(5, 24, factor=False)
    dI[2] = [7.0, 7.5]
    names = ["const"] + [f"dSpline.{k + 1}" for k in range(6)]
    got = iak.makeXvX(cov, dI, names, allKnotsd=ALLKNOTS, nDiscPts=10, ctx=ctx)["X"]
    assert np.isnan(got[2, 1:]).all() and np.isfinite(np.delete(got, 2, axis=0)).all()


@pytest.mark.parametrize("spline,opt", [(False, 0), (True, 0), (True, 1)])
def test_cubist_rule_design(ctx, spline, opt):
    cm = cubist_model(spline, opt)
    D, dI = cubist_data(900, 31)
    dI[:40] = intervals(40, 5) * 0 + np.array([[0.3, 0.5]])          # supports that start / end ON rule breaks
    ak = ALLKNOTS if spline else ()
    want = mx.makeXvX(D, dI, cm, allKnotsd=ak, opt_dSpline=opt, nDiscPts=10)
    got = iak.makeXvX(D, dI, cm, allKnotsd=ak, opt_dSpline=opt, nDiscPts=10, ctx=ctx)
    assert col_err(got["X"], want["X"]) <= TOLX
    assert np.allclose(got["XLims"], want["XLims"], rtol=1e-12, atol=0)
    L = np.vstack([np.quantile(want["X"], 0.1, axis=0) - 0.1, np.quantile(want["X"], 0.9, axis=0) + 0.1])
    want = mx.makeXvX(D, dI, cm, allKnotsd=ak, opt_dSpline=opt, nDiscPts=10, XLims=L)
    got = iak.makeXvX(D, dI, cm, allKnotsd=ak, opt_dSpline=opt, nDiscPts=10, XLims=L, ctx=ctx)
    assert col_err(got["X"], want["X"]) <= TOLX


def test_cubist_nan_covariate_blanks_the_row(ctx):
    cm = cubist_model(False)
    D, dI = cubist_data(50, 33)
    D[7, 0] = np.nan
    want = mx.makeXvX(D, dI, cm)["X"]
    got = iak.makeXvX(D, dI, cm, ctx=ctx)["X"]
    assert np.isnan(want[7]).all() and col_err(got, want) <= TOLX


def test_grid_kriging_with_the_design_made_on_the_device(ctx):
    """predictIAK3D's loops (predictIAK3D.R:144-208) with makeXvX inside them, against one grid staging."""
    ds = synth.make_dataset(700, 15)
    rng = np.random.default_rng(3)
    # covariates of the data rows: smooth functions of location, so that map cells have them too
    covf = lambda xy: {"elev": 300 + 40 * np.cos(0.05 * xy[:, 0] + 0.3), "twi": np.sin(0.07 * xy[:, 1] + 0.1 * xy[:, 0])}
    XData = mx.makeXvX(covf(ds["xData"]), ds["dIData"], 1)["X"]
    beta = rng.normal(0, 1, XData.shape[1]) * np.array([1, 0.01, 1, 1, 0.01, 1])
    z = ds["zData"] - ds["XData"] @ np.zeros(ds["XData"].shape[1]) + XData @ beta
    P, modelx, types, cmeOpt = case_pars("matern0.5", True, 0.5)
    sm = iak.setupIAK3D(ds["xData"], ds["dIData"], ctx=ctx)
    got = iak.nllIAK3D(None, z, XData, modelx=modelx, nud=0.5, sdfdType_cd1=types[0], sdfdType_cxd0=types[1], sdfdType_cxd1=types[2],
                       cmeOpt=cmeOpt, setupMats=sm, parBnds=PARBNDS, rtnAll=True, attachBigMats=True, parsBTfmd=P)
    want = orc.nllIAK3D(None, z, XData, modelx, 0.5, *types, cmeOpt, True, orc.setupIAK3D(ds["xData"], ds["dIData"]), PARBNDS,
                        rtnAll=True, parsBTfmd=P)
    want["xData"], want["dIData"] = ds["xData"], ds["dIData"]
    xMap = np.vstack([rng.uniform(0, 100, (450, 2)), ds["xData"][::70]])
    dIMap = synth.GSM_INTERVALS
    covMap = covf(xMap)
    p = XData.shape[1]
    L = np.vstack([np.full(p, -1e9), np.full(p, 1e9)]); L[1, 0] = 0.4       # finite limits: the batch-dependent recycling is live
    nxPerBatch = 200
    def Xfn(i, idx):
        sub = {k: v[idx] for k, v in covMap.items()}
        return mx.makeXvX(sub, np.repeat(dIMap[i:i + 1], len(idx), axis=0), 1, nDiscPts=10, XLims=L)["X"]
    zw, vw, _, _ = orc.predictIAK3D(xMap, dIMap, Xfn, want, modelx, types, cmeOpt, nxPerBatch=nxPerBatch)
    dm = iak.DesignModel(1, ["elev", "twi"], None, (), 0, 10, L, ctx=ctx)
    fit = got["fit"]
    fit.stage_grid(dm, xMap, np.column_stack([covMap["elev"], covMap["twi"]]), dIMap, nxPerBatch=nxPerBatch)
    Xg = fit.fetch_X()
    Xw = np.vstack([np.vstack([Xfn(i, np.arange(b, min(b + nxPerBatch, len(xMap)))) for b in range(0, len(xMap), nxPerBatch)])
                    for i in range(len(dIMap))])
    assert col_err(Xg, Xw) <= TOLX
    fit.enqueue()
    zg, vg = fit.fetch()
    assert scaled_err(zg.reshape(len(dIMap), -1), zw) <= 1e-8
    assert scaled_err(vg.reshape(len(dIMap), -1), vw) <= 1e-8
    dm.close()


@pytest.mark.parametrize("modelX,nDisc", [(1.5, 100), (3, 10)])
def test_mlr_lognormal_branch(ctx, modelX, nDisc):
    """lnTfmdData = TRUE (makeXvX.R:246-290): means and var() of the point designs; vXU on the depth-dependent columns."""
    cov, levels, dI = mlr_case(260, 41)
    want = mx.makeXvX_lnN(cov, dI, modelX, covLevels=levels, nDiscPts=nDisc)
    got = iak.makeXvX(cov, dI, modelX, covLevels=levels, nDiscPts=nDisc, lnTfmdData=True, ctx=ctx)
    assert got["iU"] == want["iU"] and got["namesX"] == want["namesX"]
    assert col_err(got["X"], want["X"]) <= TOLX
    assert got["vXU"].shape == want["vXU"].shape
    assert np.max(np.abs(got["vXU"] - want["vXU"])) <= 1e-11 * np.max(np.abs(want["vXU"]))
    assert np.allclose(got["XLims"], want["XLims"], rtol=1e-12, atol=0)
    # given limits clamp the POINT values
    L = np.vstack([np.quantile(want["X"], 0.2, axis=0) - 0.5, np.quantile(want["X"], 0.8, axis=0) + 0.5])
    want = mx.makeXvX_lnN(cov, dI, modelX, covLevels=levels, nDiscPts=nDisc, XLims=L)
    got = iak.makeXvX(cov, dI, modelX, covLevels=levels, nDiscPts=nDisc, lnTfmdData=True, XLims=L, ctx=ctx)
    assert col_err(got["X"], want["X"]) <= TOLX
    assert np.max(np.abs(got["vXU"] - want["vXU"])) <= 1e-11 * max(np.max(np.abs(want["vXU"])), 1e-300)


def test_mlr_lognormal_with_bspline_columns(ctx):
    cov, levels, dI = mlr_case(120, 42, factor=False)
    names = ["const", "elev", "d.twi"] + [f"dSpline.{k + 1}" for k in range(6)]
    want = mx.makeXvX_lnN(cov, dI, names, allKnotsd=ALLKNOTS, nDiscPts=20)
    got = iak.makeXvX(cov, dI, names, allKnotsd=ALLKNOTS, nDiscPts=20, lnTfmdData=True, ctx=ctx)
    assert got["iU"] == want["iU"] == [2, 3, 4, 5, 6, 7, 8]
    assert col_err(got["X"], want["X"]) <= TOLX
    assert np.max(np.abs(got["vXU"] - want["vXU"])) <= 1e-11 * np.max(np.abs(want["vXU"]))


@pytest.mark.parametrize("opt", [0, 1])
def test_cubist_lognormal_branch(ctx, opt):
    """makeXvX.R:437-600: the rule design at nDiscPts depths per support; iU = every column."""
    cm = cubist_model(True, opt)
    D, dI = cubist_data(150, 43)
    want = mx.makeXvX_lnN(D, dI, cm, allKnotsd=ALLKNOTS, opt_dSpline=opt, nDiscPts=10)
    got = iak.makeXvX(D, dI, cm, allKnotsd=ALLKNOTS, opt_dSpline=opt, nDiscPts=10, lnTfmdData=True, ctx=ctx)
    assert got["iU"] == list(range(len(cm["namesX"])))
    assert col_err(got["X"], want["X"]) <= TOLX
    assert np.max(np.abs(got["vXU"] - want["vXU"])) <= 1e-11 * np.max(np.abs(want["vXU"]))
    assert np.allclose(got["XLims"], want["XLims"], rtol=1e-12, atol=0)
