"""CPU: the library's host tail of a likelihood evaluation (iak3d_reml_tail, csrc/tail.cu; fitIAK3D.R:821-880) against the oracle's
REML / ML value and a direct NumPy evaluation of the formulas -- no GPU involved (the routine is host arithmetic)."""
import math

import numpy as np
import pytest

from iak3d_b200 import api
from oracle import iak3d_oracle as orc


def _terms(n, p, seed):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, p)); X[:, 0] = 1.0
    z = X @ rng.standard_normal(p) + rng.standard_normal(n)
    d = np.abs(np.subtract.outer(np.arange(n), np.arange(n)))
    A = 1.5 * np.eye(n) + 0.4 * np.exp(-d / 7.0)
    W = np.column_stack([X, z])
    return W.T @ np.linalg.solve(A, W), float(np.linalg.slogdet(A)[1])


@pytest.mark.parametrize("useReml", [True, False])
@pytest.mark.parametrize("n,p", [(40, 1), (300, 8), (500, 23)])
def test_tail_matches_direct_formulas(useReml, n, p):
    G, lnd = _terms(n, p, 11 * n + p)
    t = api.reml_tail(G, lnd, n, p, useReml)
    XiAX, XiAz, ziAz = G[:p, :p], G[:p, p], G[p, p]
    b = np.linalg.solve(XiAX, XiAz)
    res = ziAz - 2.0 * b @ XiAz + b @ XiAX @ b
    c = res / (n - p) if useReml else res / n
    if useReml:
        want = 0.5 * ((n - p) * math.log(2 * math.pi) + lnd + n * math.log(c) + np.linalg.slogdet(XiAX)[1] - p * math.log(c) + res / c)
    else:
        want = 0.5 * (n * math.log(2 * math.pi) + lnd + n * math.log(c) + res / c)
    assert abs(t["nll"] - want) <= 1e-12 * abs(want)
    assert np.max(np.abs(np.ravel(t["betahat"]) - b)) <= 1e-10 * np.max(np.abs(b))
    assert abs(t["cxdhat"] - c) <= 1e-12 * c and abs(t["resiAres"] - res) <= 1e-10 * abs(res)


def test_tail_reads_the_upper_triangle_only_and_flags_failures():
    G, lnd = _terms(120, 5, 3)
    good = api.reml_tail(G, lnd, 120, 5, True)
    Gl = G.copy(); Gl[np.tril_indices(6, -1)] = 1e9            # forceSymmetric keeps the upper triangle (fitIAK3D.R:811)
    assert api.reml_tail(Gl, lnd, 120, 5, True)["nll"] == good["nll"]
    bad = G.copy(); bad[2, 2] = -1.0                           # XiAX not positive definite -> badnll (fitIAK3D.R:885)
    assert api.reml_tail(bad, lnd, 120, 5, True)["nll"] == api.BADNLL
    nan = G.copy(); nan[0, 5] = np.nan
    assert api.reml_tail(nan, lnd, 120, 5, True)["nll"] == api.BADNLL
    assert api.reml_tail(G, float("nan"), 120, 5, True)["nll"] == api.BADNLL
    vals = api.reml_tail_batch(np.stack([G, bad, G]), np.array([lnd, lnd, lnd]), 120, 5, True)
    assert vals[0] == good["nll"] == vals[2] and vals[1] == api.BADNLL


def test_tail_composite_likelihood_counts():
    """n_for_cxd / n_for_lndet: the composite likelihood's own divisors (compositeLikIAK3D.R:206-240)."""
    G, lnd = _terms(200, 4, 9)
    t = api.reml_tail(G, lnd, 200, 4, False, n_for_cxd=777.0, n_for_lndet=555.0)
    b = np.linalg.solve(G[:4, :4], G[:4, 4])
    res = G[4, 4] - 2.0 * b @ G[:4, 4] + b @ G[:4, :4] @ b
    c = res / 777.0
    want = 0.5 * (200 * math.log(2 * math.pi) + lnd + 555.0 * math.log(c) + res / c)
    assert abs(t["nll"] - want) <= 1e-12 * abs(want)
