"""GPU parity: explicit inverse (chol2inv / solve(C)) and the Newton-Raphson terms of nrUpdatesIAK3D.R against the
oracle.  Tolerance: <= 1e-8 relative on nll; gradient / Hessian / FIM entries to 1e-7 of their scale (they are sums
of n^2 products of an explicit inverse)."""
import numpy as np
import pytest

import iak3d_b200 as iak
from iak3d_b200 import synth
from oracle import iak3d_oracle as orc
from helpers import PARBNDS, case_pars, scaled_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [5, 64, 130, 700])
def test_chol2inv_vs_lapack(ctx, n):
    rng = np.random.default_rng(n)
    Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    A = (Q * np.geomspace(1.0, 1e3, n)) @ Q.T; A = 0.5 * (A + A.T)
    got = iak.lndetANDinvCb(A, None, ctx=ctx)
    want_lnd, want = orc.lndetANDinvCb(A, None)
    assert abs(got["lndetC"] - want_lnd) <= 1e-8 * max(1.0, abs(want_lnd))
    assert scaled_err(got["invCb"], want) <= 1e-8
    assert scaled_err(A @ got["invCb"], np.eye(n)) <= 1e-8


def test_kept_fit_returns_iC(ctx):
    ds = synth.make_dataset(500, 8)
    P, modelx, types, cmeOpt = case_pars("edgeroi", False, 0.5)
    sm = iak.setupIAK3D(ds["xData"], ds["dIData"], ctx=ctx)
    fit = iak.nllIAK3D(None, ds["zData"], ds["XData"], modelx=modelx, nud=0.5, sdfdType_cd1=types[0], sdfdType_cxd0=types[1],
                       sdfdType_cxd1=types[2], cmeOpt=cmeOpt, setupMats=sm, parBnds=PARBNDS, rtnAll=True, attachBigMats=True, parsBTfmd=P)
    Cm, iC = fit["fit"].get_big()
    assert scaled_err(Cm @ iC, np.eye(500)) <= 1e-7
    want = orc.nllIAK3D(None, ds["zData"], ds["XData"], modelx, 0.5, *types, cmeOpt, True, orc.setupIAK3D(ds["xData"], ds["dIData"]),
                        PARBNDS, rtnAll=True, parsBTfmd=P)
    assert scaled_err(iC, want["iC"]) <= 1e-7


@pytest.mark.parametrize("kind,nonstat,ncomp,n", [("matern0.5", False, 6, 400), ("matern0.5", False, 5, 257), ("edgeroi", False, 6, 640),
                                                  ("matern1.5", True, 6, 300)])
def test_gradhess_vs_oracle(ctx, kind, nonstat, ncomp, n):
    ds = synth.make_dataset(n, 4)
    P, modelx, types, cmeOpt = case_pars(kind, nonstat, 0.5)
    W = np.column_stack([ds["XData"], ds["zData"]])
    sm = iak.SetupMats(ds["xData"], ds["dIData"], W=W, ctx=ctx)
    keys = ["cx0", "cx1", "cd1", "cxd0", "cxd1", "cme"][:ncomp]
    th = np.log([max(P[k], 1e-3) for k in keys]) + 0.1 * np.arange(ncomp)
    got = iak.gradHessIAK3D(sm, P, modelx, types, th)
    osm = orc.setupIAK3D(ds["xData"], ds["dIData"])
    dA = orc.component_matrices(P, modelx, types, osm)[:ncomp]
    want = orc.gradHessIAK3D(dA, th, W)
    assert abs(got["nll"] - want["nll"]) <= 1e-8 * abs(want["nll"])
    assert scaled_err(got["grad"], want["grad"]) <= 1e-7
    assert scaled_err(got["hess"], want["hess"]) <= 1e-7
    assert scaled_err(got["fim"], want["fim"]) <= 1e-7
    assert scaled_err(got["betahat"], want["betahat"]) <= 1e-8
    assert scaled_err(got["vbetahat"], want["vbetahat"]) <= 1e-8
    assert np.array_equal(got["hess"], got["hess"].T)


def test_gradhess_not_spd(ctx):
    ds = synth.make_dataset(100, 4)
    P, modelx, types, cmeOpt = case_pars("matern0.5", False, 0.5)
    W = np.column_stack([ds["XData"], ds["zData"]])
    sm = iak.SetupMats(ds["xData"], ds["dIData"], W=W, ctx=ctx)
    got = iak.gradHessIAK3D(sm, P, modelx, types, np.array([-800.0] * 6))     # all variances underflow to 0
    assert np.isnan(got["nll"])


@pytest.mark.parametrize("kind,ncomp,n", [("matern0.5", 6, 400), ("edgeroi", 5, 300)])
def test_nrUpdates_iteration_vs_oracle(ctx, kind, ncomp, n):
    """The whole Newton-Raphson iteration (nrUpdatesIAK3D.R:1-116): same number of updates, same optimum, small variances
    snapped to -20.  The Edgeroi model has no cd1 component, so its dA[3] is a zero matrix and the FIM is singular: both
    sides must leave at once with errFlag 1 (:74-78)."""
    ds = synth.make_dataset(n, 6)
    P, modelx, types, cmeOpt = case_pars(kind, False, 0.5)
    W = np.column_stack([ds["XData"], ds["zData"]])
    sm = iak.SetupMats(ds["xData"], ds["dIData"], W=W, ctx=ctx)
    osm = orc.setupIAK3D(ds["xData"], ds["dIData"])
    cxdhat = orc.nllIAK3D(None, ds["zData"], ds["XData"], modelx, 0.5, *types, cmeOpt, True, osm, PARBNDS, rtnAll=True, parsBTfmd=P)["cxdhat"]
    keys = ["cx0", "cx1", "cd1", "cxd0", "cxd1", "cme"][:ncomp]
    th0 = np.log([max(cxdhat * P[k], 1e-3) for k in keys]) + 0.2 * np.array([1, -1, 1, -1, 1, -1])[:ncomp]
    got = iak.nrUpdatesIAK3D(sm, P, modelx, types, th0)
    dA = orc.component_matrices(P, modelx, types, osm)[:ncomp]
    want = orc.nrUpdatesIAK3D(dA, th0, W)
    assert got["noits"] == want["noits"] and got["errFlag"] == want["errFlag"]
    if want["errFlag"] == 1:
        assert kind == "edgeroi" and want["noits"] == 0
        return
    assert want["noits"] >= 4
    assert abs(got["nll_unrounded"] - want["nll_unrounded"]) <= 1e-8 * abs(want["nll_unrounded"])
    assert abs(got["nll"] - want["nll"]) <= 1.01e-5
    live = want["lnvPars"] > -19.0                              # snapped-to--20 variances sit on a flat likelihood
    assert np.array_equal(live, got["lnvPars"] > -19.0)
    assert np.max(np.abs(got["lnvPars"][live] - want["lnvPars"][live])) <= 1e-5
    assert scaled_err(got["betahat"], want["betahat"]) <= 1e-7


class _FakeComm:
    """One process playing every rank in turn: rank r's call returns its own share; the 'all-reduce' adds the stored shares
    of the other ranks (computed beforehand with the same inputs)."""
    def __init__(self, rank, world, others):
        self.rank, self.world, self.others = rank, world, others
    def allreduce_sum(self, buf):
        self.mine = np.array(buf, float)
        return self.mine + sum(self.others)


@pytest.mark.parametrize("world,n", [(2, 400), (3, 700), (8, 1100)])
def test_gradhess_shares_add_up(ctx, world, n):
    """the trace terms of nrUpdatesIAK3D.R:197-210 cut over `world` ranks by tile-pair ownership."""
    ds = synth.make_dataset(n, 6)
    P, modelx, types, cmeOpt = case_pars("matern0.5", False, 0.5)
    W = np.column_stack([ds["XData"], ds["zData"]])
    sm = iak.SetupMats(ds["xData"], ds["dIData"], W=W, ctx=ctx)
    th = np.log([max(P[k], 1e-3) for k in ["cx0", "cx1", "cd1", "cxd0", "cxd1", "cme"]])
    whole = iak.gradHessIAK3D(sm, P, modelx, types, th)
    shares = []
    for r in range(world):
        c = _FakeComm(r, world, [])
        iak.gradHessIAK3D(sm, P, modelx, types, th, comm=c)
        shares.append(c.mine)
    assert all(np.any(s[:6 + 36] != 0) for s in shares[:min(world, 3)])          # every rank really owns tiles
    for r in (0, world - 1):
        c = _FakeComm(r, world, [s for k, s in enumerate(shares) if k != r])
        got = iak.gradHessIAK3D(sm, P, modelx, types, th, comm=c)
        for k in ("grad", "hess", "fim"):
            assert scaled_err(got[k], whole[k]) <= 1e-11
        assert got["nll"] == whole["nll"]
