"""CPU: the spline basis of the sd functions (splineBasisFns.R:158-325) and the oracle's setCApproxIAK3D[2]
(fitIAK3D.R:1919-2233) pinned by their defining properties; host mirror == oracle restatement."""
import numpy as np
import pytest

from iak3d_b200 import api
from iak3d_b200 import synth
from oracle import iak3d_oracle as orc
from helpers import SPLINE_CASES, SPLINE_KNOTS, spline_case_pars

BK = (0.1, 1.5)
IK = np.array([0.3, 0.6, 1.1])


def test_basis_point_properties():
    """paramVs = 1: (a, g.., w) with a = value at x = 0, w = value at the upper boundary knot, zero gradient beyond it,
    linear below the lower boundary knot (natural spline)."""
    par = np.array([0.7, -0.4, 0.25, 1.9])
    f = lambda x: orc.makeXnsClampedUG0(np.asarray(x, float), BK, IK, paramVs=1) @ par
    assert abs(f([0.0])[0] - par[0]) < 1e-14
    assert abs(f([BK[1]])[0] - par[-1]) < 1e-12
    assert np.allclose(f([1.5, 1.7, 2.0, 5.0]), par[-1], atol=1e-11)           # clamped: constant beyond xU
    lo = f([0.0, 0.03, 0.06, 0.09])
    assert np.allclose(np.diff(lo, 2), 0.0, atol=1e-13)                          # linear below xL
    x = np.linspace(0, 1.6, 4001)
    y = f(x)
    assert np.max(np.abs(np.diff(y, 3))) < 5e-8                                  # C2: third differences of a cubic stay O(h^3)


@pytest.mark.parametrize("paramVs", [0, 1])
def test_basis_interval_average_is_the_average(paramVs):
    rng = np.random.default_rng(5)
    a = np.round(rng.uniform(0, 1.7, 40), 2); b = np.round(a + rng.choice([0.05, 0.1, 0.3, 0.5], 40), 2)
    dI = np.column_stack([a, b])
    Xia = orc.makeXnsClampedUG0(dI, BK, IK, paramVs=paramVs)
    # Gauss-Legendre between consecutive knots inside each interval (piecewise cubic => exact)
    gx, gw = np.polynomial.legendre.leggauss(6)
    for i in range(dI.shape[0]):
        brk = np.unique(np.concatenate([[dI[i, 0], dI[i, 1]], [k for k in (*BK, *IK) if dI[i, 0] < k < dI[i, 1]]]))
        acc = np.zeros(Xia.shape[1])
        for lo, hi in zip(brk[:-1], brk[1:]):
            xs = 0.5 * (hi - lo) * gx + 0.5 * (hi + lo)
            acc += 0.5 * (hi - lo) * (gw @ orc.makeXnsClampedUG0(xs, BK, IK, paramVs=paramVs))
        assert np.allclose(Xia[i], acc / (dI[i, 1] - dI[i, 0]), rtol=1e-12, atol=1e-13)


@pytest.mark.parametrize("paramVs", [0, 1])
def test_host_mirror_basis_equals_oracle(paramVs):
    rng = np.random.default_rng(9)
    a = np.round(rng.uniform(0, 1.7, 60), 2); dI = np.column_stack([a, np.round(a + 0.2, 2)])
    for x in (dI, a):
        for ik in (IK, IK[:1], []):
            A = api.makeXnsClampedUG0(x, BK, ik, paramVs=paramVs); B = orc.makeXnsClampedUG0(x, BK, ik, paramVs=paramVs)
            assert A.shape == B.shape and np.allclose(A, B, rtol=1e-14, atol=1e-15)


def test_iasdfd_exponential_is_interval_average():
    dI = np.array([[0.0, 0.1], [0.3, 0.6], [1.0, 2.0]])
    got = orc.iasdfd(dI, (0.6, 1.7), -1)
    for (a, b), g in zip(dI, got):
        xs = np.linspace(a, b, 20001)
        assert abs(g - np.trapezoid((1 - 0.6) + 0.6 * np.exp(-1.7 * xs), xs) / (b - a)) < 1e-9


@pytest.mark.parametrize("label,kind,types,coefs", SPLINE_CASES)
def test_approx_consistency_square_vs_cross(label, kind, types, coefs):
    """The reference's own (dead) check compositeLikIAK3D.R:382 carried to the approximation:
    setCApproxIAK3D2(S, S) + cme I == setCApproxIAK3D(S) (without the cd1 quirk)."""
    ds = synth.make_dataset(220, 77)
    P, modelx, types, cmeOpt = spline_case_pars(kind, types, coefs)
    sm = orc.setupIAK3D(ds["xData"], ds["dIData"], types, SPLINE_KNOTS)
    C, s2 = orc.setCIAK3D(P, modelx, *types, cmeOpt, sm, quirk_cd1_unscaled=False)
    sm2 = orc.setupIAK3D2(ds["xData"], ds["dIData"], ds["xData"], ds["dIData"], types, SPLINE_KNOTS)
    C2 = orc.setCIAK3D2(P, modelx, *types, cmeOpt, sm2)
    assert np.allclose(C2 + P["cme"] * np.eye(220), C, rtol=1e-13, atol=1e-15)
    assert np.all(s2 >= np.diag(C) - P["cme"] - 1e-12)      # point variance >= variance of an interval average
    assert np.all(np.linalg.eigvalsh(C) > 0)


def test_approx_equals_exact_when_sd_is_constant():
    """With all spline coefficients equal to 1 the sd function is 1 everywhere (a = 1, w = 1 and the g's of a
    constant): the approximation must reproduce the stationary model."""
    ds = synth.make_dataset(150, 3)
    P, modelx, _, cmeOpt = synth.default_pars("matern0.5", False)
    sm = orc.setupIAK3D(ds["xData"], ds["dIData"], (1, 1, 1), dict(bdryKnots=(0.0, 1.6), intKnots_cd1=[0.4], intKnots_cxd0=[0.4], intKnots_cxd1=[0.4]))
    Pc = dict(P); Pc.update(sdfdPars_cd1=np.array([1.0]), sdfdPars_cxd0=np.array([1.0]), sdfdPars_cxd1=np.array([1.0]))
    Ca, _ = orc.setCIAK3D(Pc, modelx, 1, 1, 1, cmeOpt, sm, quirk_cd1_unscaled=False)
    Cs, _ = orc.setCIAK3D(P, modelx, 0, 0, 0, cmeOpt, orc.setupIAK3D(ds["xData"], ds["dIData"]))
    assert np.allclose(Ca, Cs, rtol=1e-12, atol=1e-14)


def test_negative_sd_gives_na():
    ds = synth.make_dataset(60, 4)
    P, modelx, types, cmeOpt = spline_case_pars("matern0.5", (2, 3, 1), dict(sdfdPars_cd1=[0.5, -3.0], sdfdPars_cxd0=[1, 1, 1], sdfdPars_cxd1=[1]))
    sm = orc.setupIAK3D(ds["xData"], ds["dIData"], types, SPLINE_KNOTS)
    assert orc.setCIAK3D(P, modelx, *types, cmeOpt, sm)[0] is None
