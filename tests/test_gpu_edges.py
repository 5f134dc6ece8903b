"""GPU parity on the edges of the input space: degenerate unique sets, the largest number of fixed effects the bordered
rows hold, empty prediction sets, non-positive-definite and invalid inputs (status codes, never a crash)."""
import numpy as np
import pytest

import iak3d_b200 as iak
from iak3d_b200 import synth
from oracle import iak3d_oracle as orc
from helpers import PARBNDS, case_pars, relerr, scaled_err

pytestmark = pytest.mark.gpu


def _nll_both(ctx, x, dI, z, X, kind="matern0.5", nonstat=False, **over):
    P, modelx, types, cmeOpt = case_pars(kind, nonstat, 0.5)
    P = dict(P); P.update(over)
    sm = iak.setupIAK3D(x, dI, ctx=ctx)
    got = iak.nllIAK3D(None, z, X, modelx=modelx, nud=0.5, sdfdType_cd1=types[0], sdfdType_cxd0=types[1], sdfdType_cxd1=types[2], cmeOpt=cmeOpt,
                       setupMats=sm, parBnds=PARBNDS, parsBTfmd=P)
    want = orc.nllIAK3D(None, z, X, modelx, 0.5, *types, cmeOpt, True, orc.setupIAK3D(x, dI), PARBNDS, parsBTfmd=P)
    return got, want


def test_single_location_many_intervals(ctx):
    """Every observation at ONE location (nxU = 1): the horizontal table is a single 1."""
    rng = np.random.default_rng(1)
    n = 150
    top = np.round(np.sort(rng.uniform(0, 1.8, n)), 2)
    dI = np.column_stack([top, np.round(top + rng.choice([0.05, 0.1, 0.2], n), 2)])
    x = np.tile([[12.5, 40.0]], (n, 1))
    X = np.column_stack([np.ones(n), dI.mean(axis=1)]); z = rng.standard_normal(n) + 3.0
    got, want = _nll_both(ctx, x, dI, z, X)
    assert abs(got - want) <= 1e-8 * abs(want)


def test_single_interval_many_locations(ctx):
    """Every observation on the SAME depth interval (ndIU = 1): all depth tables are 1 x 1."""
    rng = np.random.default_rng(2)
    n = 300
    x = np.round(rng.uniform(0, 100, (n, 2)), 3); dI = np.tile([[0.0, 0.1]], (n, 1))
    X = np.column_stack([np.ones(n), rng.standard_normal(n)]); z = rng.standard_normal(n)
    for kind in ("matern0.5", "edgeroi", "nugget"):
        got, want = _nll_both(ctx, x, dI, z, X, kind)
        assert abs(got - want) <= 1e-8 * abs(want), kind


def test_replicated_supports_need_the_measurement_error(ctx):
    """Two observations on the same support (a collision): singular without cme -> badnll exactly where the reference's
    chol fails; fine with cme."""
    ds = synth.make_dataset(120, 5)
    x = np.vstack([ds["xData"], ds["xData"][:7]]); dI = np.vstack([ds["dIData"], ds["dIData"][:7]])
    z = np.concatenate([ds["zData"], ds["zData"][:7] + 0.1]); X = np.vstack([ds["XData"], ds["XData"][:7]])
    got, want = _nll_both(ctx, x, dI, z, X)
    assert abs(got - want) <= 1e-8 * abs(want)
    got0, want0 = _nll_both(ctx, x, dI, z, X, cme=0.0)
    assert got0 == 9e99 and want0 == 9e99


@pytest.mark.parametrize("p", [1, 31, 63])
def test_widest_design_matrix(ctx, p):
    """p = 63 fixed effects + z fill the 64 bordered rows (the ABI's limit); p = 64 is refused."""
    ds = synth.make_dataset(400, 6)
    rng = np.random.default_rng(p)
    X = np.column_stack([np.ones(400), rng.standard_normal((400, p - 1))]) if p > 1 else np.ones((400, 1))
    got, want = _nll_both(ctx, ds["xData"], ds["dIData"], ds["zData"], X)
    assert abs(got - want) <= 1e-8 * abs(want)
    if p == 63:
        sm = iak.setupIAK3D(ds["xData"], ds["dIData"], ctx=ctx)
        with pytest.raises(iak.Iak3dError):
            sm.set_W(np.column_stack([X, np.ones(400), ds["zData"]]))


def test_empty_and_tiny_prediction_sets(ctx):
    ds = synth.make_dataset(200, 7)
    P, modelx, types, cmeOpt = case_pars("edgeroi", False, 0.5)
    sm = iak.setupIAK3D(ds["xData"], ds["dIData"], ctx=ctx)
    fit = iak.nllIAK3D(None, ds["zData"], ds["XData"], modelx=modelx, nud=0.5, sdfdType_cd1=types[0], sdfdType_cxd0=types[1],
                       sdfdType_cxd1=types[2], cmeOpt=cmeOpt, setupMats=sm, parBnds=PARBNDS, rtnAll=True, attachBigMats=True, parsBTfmd=P)
    z, v = fit["fit"].predict(np.zeros((0, 2)), np.zeros((0, 2)), np.zeros((0, ds["XData"].shape[1])))
    assert z.size == 0 and v.size == 0
    z1, v1 = fit["fit"].predict(ds["xData"][:1], ds["dIData"][:1], ds["XData"][:1])
    z5, v5 = fit["fit"].predict(ds["xData"][:5], ds["dIData"][:5], ds["XData"][:5])
    assert z1[0] == z5[0] and v1[0] == v5[0]                      # a support's prediction does not depend on its batch
    # all design rows NaN: everything stays NaN (predictIAK3D.R:160-163)
    xMap = synth.make_grid(4)
    pm = iak.predictIAK3D(xMap, synth.GSM_INTERVALS[:1], lambda i, idx: np.full((idx.size, ds["XData"].shape[1]), np.nan), fit)
    assert np.all(np.isnan(pm["zMap"]))


def test_extreme_but_valid_parameters(ctx):
    """Parameters at the edges of the optimiser's box: tiny / huge ranges, a vanishing component, nux at its bounds."""
    ds = synth.make_dataset(260, 9)
    for over in (dict(ax=0.05), dict(ax=60.0), dict(ad=0.02), dict(ad=1.2), dict(cx1=1e-8, cd1=1e-8), dict(nux=0.05), dict(nux=20.0),
                 dict(cme=1e-9)):
        kind = "matern0.8" if "nux" in over else "matern0.5"
        got, want = _nll_both(ctx, ds["xData"], ds["dIData"], ds["zData"], ds["XData"], kind, **over)
        assert (got == want == 9e99) or abs(got - want) <= 1e-8 * abs(want), over


def test_lndetANDinvCb_rejects_indefinite_and_handles_many_rhs(ctx):
    rng = np.random.default_rng(3)
    n = 200
    B = rng.standard_normal((n, n)); A = B @ B.T + n * np.eye(n)
    b = rng.standard_normal((n, 130))                              # more right-hand sides than one pass of bordered rows holds
    got = iak.lndetANDinvCb(A, b, ctx=ctx)
    assert scaled_err(A @ got["invCb"], b) <= 1e-10
    assert abs(got["lndetC"] - np.linalg.slogdet(A)[1]) <= 1e-10 * abs(got["lndetC"])
    A[5, 5] = -1.0
    bad = iak.lndetANDinvCb(A, b[:, :2], ctx=ctx)
    assert np.isnan(bad["lndetC"]) and bad["invCb"] is None
