"""CPU: the lognormal path.  The product runs nrUpdatesIAK3DlnN on the q x q Gram matrix F' iC F of ONE bordered
factorisation (api.lnN_design / nrUpdatesIAK3DlnN_gram); the oracle restates the reference literally with the explicit
n x n inverse and TK (nrUpdatesIAK3DlnN.R).  Here the Gram matrix comes from NumPy, so the host algebra is checked
without a GPU; the GPU parity test is tests/test_gpu_lognormal.py."""
import numpy as np
import pytest

from iak3d_b200 import api
from iak3d_b200 import synth
from oracle import iak3d_oracle as orc
from helpers import lognormal_case


def _gram(ds, z, iU, vXU, P, modelx, types, cmeOpt):
    sm = orc.setupIAK3D(ds["xData"], ds["dIData"])
    A, s2 = orc.setCIAK3D(P, modelx, *types, cmeOpt, sm)
    F, pairs = api.lnN_design(ds["XData"], z, vXU, iU)
    F[:, ds["XData"].shape[1] + 1] = s2 - np.diag(A)                      # the column the device fills
    iC = np.linalg.inv(A)
    G = F.T @ iC @ F
    lndet = np.linalg.slogdet(A)[1]
    return A, s2, G, lndet, pairs


@pytest.mark.parametrize("pU", [0, 1, 3])
def test_gram_domain_newton_equals_explicit(pU):
    ds, z, iU, vXU = lognormal_case(260, 5, pU=max(pU, 1))
    if pU == 0:
        iU, vXU = [], None
    P, modelx, types, cmeOpt = synth.default_pars("matern0.5", False)
    A, s2, G, lndet, pairs = _gram(ds, z, iU, vXU, P, modelx, types, cmeOpt)
    n, p = ds["XData"].shape
    got = api.nrUpdatesIAK3DlnN_gram(G, lndet, n, p, iU, pairs)
    want = orc.nrUpdatesIAK3DlnN(z, ds["XData"], vXU, iU, A, s2)
    assert got["noits"] == want["noits"] and got["errFlag"] == want["errFlag"] == 0
    assert abs(got["nll"] - want["nll"]) <= 1e-9 * abs(want["nll"])
    assert np.allclose(got["betahat"], want["betahat"], rtol=1e-8, atol=1e-10)
    if pU:
        assert np.allclose(got["fim"], want["fim"], rtol=1e-8) and np.allclose(got["hess"], want["hess"], rtol=1e-7, atol=1e-9)
        assert np.max(np.abs(want["grad"])) < 1e-6 * np.max(np.abs(want["fim"]))     # converged


def test_newton_finds_the_minimum_of_the_profile_likelihood():
    """Independent check of the oracle: at the returned betahat the lognormal nll (nrUpdatesIAK3DlnN.R:213 with betaK
    profiled out) is stationary and not larger than at perturbed betaU."""
    ds, z, iU, vXU = lognormal_case(200, 8, pU=2)
    P, modelx, types, cmeOpt = synth.default_pars("edgeroi", False)
    sm = orc.setupIAK3D(ds["xData"], ds["dIData"])
    A, s2 = orc.setCIAK3D(P, modelx, *types, cmeOpt, sm)
    t = orc.nrUpdatesIAK3DlnN(z, ds["XData"], vXU, iU, A, s2)
    n, p = ds["XData"].shape
    iC = np.linalg.inv(A); lndet = np.linalg.slogdet(A)[1]

    def full_nll(beta):
        mu = orc.setmuIAK3D(ds["XData"], vXU, iU, beta, np.diag(A), s2)
        r = z.reshape(-1, 1) - mu
        return 0.5 * (n * np.log(2 * np.pi) + lndet + (r.T @ iC @ r)[0, 0])
    f0 = full_nll(t["betahat"])
    assert abs(f0 - t["nll_unrounded"]) <= 1e-9 * abs(f0)
    rng = np.random.default_rng(0)
    for _ in range(20):
        assert full_nll(t["betahat"] + 1e-3 * rng.standard_normal((p, 1))) >= f0 - 1e-9


def test_setmu_and_betavXbeta():
    rng = np.random.default_rng(1)
    n, pU = 7, 3
    v = rng.standard_normal((n, pU, pU)); v = v + np.transpose(v, (0, 2, 1))
    b = rng.standard_normal(pU)
    want = np.array([b @ v[i] @ b for i in range(n)]).reshape(n, 1)
    assert np.allclose(api.calcbetavXbeta(v.reshape(n * pU, pU), b), want)
    assert np.allclose(orc.calcbetavXbeta(v.reshape(n * pU, pU), b), want)
    X = rng.standard_normal((n, 5)); beta = rng.standard_normal(5); iU = [1, 3, 4]
    m1 = api.setmuIAK3D(X, v.reshape(n * pU, pU), iU, beta, np.ones(n), 2 * np.ones(n))
    m2 = orc.setmuIAK3D(X, v.reshape(n * pU, pU), iU, beta, np.ones(n), 2 * np.ones(n))
    assert np.allclose(m1, m2) and np.allclose(m1, X @ beta.reshape(-1, 1) + 0.5 * (1 + np.array([beta[iU] @ v[i] @ beta[iU] for i in range(n)]).reshape(n, 1)))
