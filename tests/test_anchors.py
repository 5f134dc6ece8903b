"""The oracle against tests/golden/mpmath_anchors.json -- values NO repository code generated: 40-digit mpmath evaluation of
the model's definitions (quadrature of the defining double integral, besselk, dense Cholesky, the textbook universal-kriging
system; scripts/make_mpmath_anchors.py imports neither the oracle nor the package).  They anchor the stages the
reference holds no vectors for: setCIAK3D assembly (fitIAK3D.R:955-1112), the REML / ML tail (:849-878) and the kriging
algebra (predictIAK3D.R:202-249).  Tolerances are north_star's: 1e-10 on covariance entries, 1e-8 on the rest."""
import json
import os

import numpy as np
import pytest

from oracle import iak3d_oracle as orc
from helpers import PARBNDS

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "mpmath_anchors.json")) as f:
    ANCHORS = json.load(f)
CASES = ANCHORS["cases"]


def anchor_inputs(c):
    """(parsBTfmd, modelx, types, cmeOpt, xData, dIData, XData, zData) of an anchor case"""
    P = {k: (float("nan") if v is None else float(v)) for k, v in c["pars"].items()}
    P["nud"] = c["nud"]
    for comp in ("cd1", "cxd0", "cxd1"):
        t = c["tau"][comp]
        P["sdfdPars_" + comp] = np.zeros(0) if t is None else np.array(t, float)
    return (P, c["modelx"], tuple(c["types"]), 1, np.array(c["xData"]), np.array(c["dIData"]), np.array(c["XData"]),
            np.array(c["zData"]))


def lower(c):
    n = c["n"]
    A = np.zeros((n, n))
    for i, row in enumerate(c["A_lower"]):
        A[i, :i + 1] = row
    return A + np.tril(A, -1).T


def cov_err(C, A):
    """max relative error of covariance entries in units of the 1e-10 tolerance.  Exact zeros of the anchor (compact
    support: spherical beyond its range) must be exact zeros; entries the model forms by cancellation just inside the range
    get the absolute floor 2e-15 max|A|."""
    C = np.asarray(C, float); A = np.asarray(A, float)
    assert np.all(C[A == 0.0] == 0.0)
    nz = A != 0.0
    tol = np.maximum(1e-10 * np.abs(A[nz]), 2e-15 * np.max(np.abs(A)))
    return float(np.max(np.abs(C[nz] - A[nz]) / tol))


def test_anchor_file_is_not_repository_output():
    src = open(os.path.join(HERE, "..", "scripts", "make_mpmath_anchors.py")).read()
    body = src.split('"""', 2)[2]
    assert "import oracle" not in body and "from oracle" not in body and "iak3d_b200" not in body
    assert ANCHORS["quadrature_check"] < 1e-25


@pytest.mark.parametrize("c", CASES, ids=[c["name"] for c in CASES])
def test_oracle_setC_vs_mpmath(c):
    P, modelx, types, cmeOpt, x, dI, X, z = anchor_inputs(c)
    sm = orc.setupIAK3D(x, dI)
    C, s2 = orc.setCIAK3D(P, modelx, *types, cmeOpt, sm)
    A = lower(c)
    assert cov_err(C, A) <= 1.0
    assert np.max(np.abs(np.asarray(s2).reshape(-1) - np.array(c["sigma2Vec"])) / np.array(c["sigma2Vec"])) <= 1e-10


@pytest.mark.parametrize("c", CASES, ids=[c["name"] for c in CASES])
def test_oracle_nll_vs_mpmath(c):
    P, modelx, types, cmeOpt, x, dI, X, z = anchor_inputs(c)
    sm = orc.setupIAK3D(x, dI)
    t = orc.nllIAK3D(None, z, X, modelx, c["nud"], *types, cmeOpt, True, sm, PARBNDS, useReml=c["useReml"], forCompLik=True, parsBTfmd=P)
    assert abs(t["lndetA"] - c["lndetA"]) <= 1e-8 * abs(c["lndetA"])
    W = np.array(c["WiAW"])
    assert np.max(np.abs(np.asarray(t["WiAW"]) - W)) <= 1e-8 * np.max(np.abs(W))
    fit = orc.nllIAK3D(None, z, X, modelx, c["nud"], *types, cmeOpt, True, sm, PARBNDS, useReml=c["useReml"], rtnAll=True, parsBTfmd=P)
    assert abs(fit["nll"] - c["nll"]) <= 1e-8 * abs(c["nll"])
    assert abs(fit["cxdhat"] - c["cxdhat"]) <= 1e-8 * c["cxdhat"]
    assert np.max(np.abs(np.asarray(fit["betahat"]).reshape(-1) - np.array(c["betahat"]))) <= 1e-8 * np.max(np.abs(c["betahat"]))


@pytest.mark.parametrize("c", [c for c in CASES if c["predict"]], ids=[c["name"] for c in CASES if c["predict"]])
def test_oracle_kriging_vs_mpmath(c):
    P, modelx, types, cmeOpt, x, dI, X, z = anchor_inputs(c)
    sm = orc.setupIAK3D(x, dI)
    fit = orc.nllIAK3D(None, z, X, modelx, c["nud"], *types, cmeOpt, True, sm, PARBNDS, useReml=c["useReml"], rtnAll=True, parsBTfmd=P)
    fit["xData"], fit["dIData"] = x, dI
    xm, dm, Xm = np.array(c["xMap"]), np.array(c["dIMap"]), np.array(c["XMap"])
    for s in range(len(xm)):      # one support at a time: predictIAK3D crosses every location with every interval
        zs, vs, _, _ = orc.predictIAK3D(xm[s:s + 1], dm[s:s + 1], lambda i, idx: Xm[s:s + 1], fit, modelx, types, cmeOpt)
        assert abs(float(np.asarray(zs).reshape(-1)[0]) - c["zMap"][s]) <= 1e-8 * abs(c["zMap"][s])
        assert abs(float(np.asarray(vs).reshape(-1)[0]) - c["vMap"][s]) <= 1e-8 * abs(c["vMap"][s])
