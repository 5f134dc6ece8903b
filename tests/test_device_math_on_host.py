"""The scalar arithmetic the table kernels run (iak3d_b200/csrc/iak3d_math.h), compiled for the host as the
test-only tests/_hostmath.so, against the oracle.  No GPU needed; the product never loads this library."""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import pytest
import scipy.special as ssp

from oracle import iak3d_oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def hm():
    so = os.path.join(HERE, "_hostmath.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-s", "tests/_hostmath.so"], cwd=ROOT)
    lib = C.CDLL(so)
    d = C.c_double
    lib.hm_avcov.argtypes = [d, d, d, d, d, d, C.c_int, d, d, C.POINTER(d)]
    lib.hm_avvar.argtypes = [d, d, d, d, C.c_int, d, d, C.POINTER(d)]
    lib.hm_besselk.argtypes = [d, d]; lib.hm_besselk.restype = d
    lib.hm_besselk_rt.argtypes = [d, d]; lib.hm_besselk_rt.restype = d
    lib.hm_besselk_fit.argtypes = [d, d]; lib.hm_besselk_fit.restype = d
    lib.hm_phix.argtypes = [d, C.c_int, d, d]; lib.hm_phix.restype = d
    return lib


def _avcov(hm, I1, I2, ad, nud, st, sp):
    out = C.c_double()
    assert hm.hm_avcov(I1[0], I1[1], I2[0], I2[1], ad, nud, st, sp[0], sp[1], C.byref(out)) == 0
    return out.value


def test_known_answers(hm):
    with open(os.path.join(HERE, "golden", "iacov_kat.json")) as f:
        kat = json.load(f)
    for case in kat["cases"]:
        for (nud, st), want in zip(kat["columns"], case["values"]):
            got = _avcov(hm, case["I1"], case["I2"], kat["ad"], nud, st, kat["sdfdPars"])
            assert abs(got - want) <= 1e-14 * abs(want), (case["case"], nud, st)


@pytest.mark.parametrize("nud", [0.5, 1.5, 2.5])
@pytest.mark.parametrize("st", [0, -1])
def test_random_intervals_match_oracle(hm, nud, st):
    rng = np.random.default_rng(7)
    sp = (0.6, 1.7)
    worst = 0.0
    for ad in (0.05, 0.35, 1.1):
        tops = np.round(rng.uniform(0, 1.8, 60), 2)
        th = rng.choice([0.01, 0.05, 0.1, 0.2, 0.5], 60)
        dI = np.column_stack([tops, np.round(tops + th, 2)])
        want = orc.iaCovMatern2(dI, dI, ad, nud, sp, st)
        for i in range(60):
            for j in range(60):
                got = _avcov(hm, dI[i], dI[j], ad, nud, st, sp)
                worst = max(worst, abs(got - want[i, j]) / abs(want[i, j]))
    assert worst <= 1e-10      # north_star: covariance entries to 1e-10 relative


def test_invalid_nud_and_type(hm):
    out = C.c_double()
    assert hm.hm_avcov(0, 1, 0, 1, 0.3, 1.0, 0, 0, 0, C.byref(out)) == 2
    assert hm.hm_avcov(0, 1, 0, 1, 0.3, 0.5, 3, 0, 0, C.byref(out)) == 2


def test_psi_tau2_clamp(hm):
    # psi - tau2 inside (-1e-3, 1e-3) is clamped (iaCovMatern.R:233-237): ad = 1/1.7 makes psi == tau2 for nud = 0.5
    ad = 1.0 / 1.7
    for sp2 in (1.7, 1.7004, 1.6996):
        got = _avcov(hm, (0.0, 0.3), (0.2, 0.7), ad, 0.5, -1, (0.6, sp2))
        want = orc.iaCovMatern2([[0.0, 0.3]], [[0.2, 0.7]], ad, 0.5, (0.6, sp2), -1)[0, 0]
        assert abs(got - want) <= 1e-10 * abs(want)


def test_avvar(hm):
    out = C.c_double()
    dI = np.array([[0.0, 0.05], [0.3, 0.6], [1.0, 2.0]])
    for st, sp in ((0, (0.0, 0.0)), (-1, (0.6, 1.7))):
        prm = orc.nsMaternParams(0.35, 1.5, sp, st)
        want = orc.avVarVec(dI, prm)
        for k in range(3):
            assert hm.hm_avvar(dI[k, 0], dI[k, 1], 0.35, 1.5, st, sp[0], sp[1], C.byref(out)) == 0
            assert abs(out.value - want[k]) <= 1e-14 * abs(want[k])


def test_besselk_against_scipy(hm):
    worst = 0.0
    for nu in (0.05, 0.3, 0.5, 0.8, 1.0, 1.5, 2.2, 3.7, 7.25, 12.0, 20.0):
        for x in np.concatenate([np.logspace(-6, np.log10(2.0), 40), np.linspace(2.0001, 700.0, 60)]):
            want = ssp.kv(nu, x)
            if not np.isfinite(want) or want == 0.0 or want < 1e-290:
                continue
            got = hm.hm_besselk(nu, float(x))
            worst = max(worst, abs(got - want) / want)
            got = hm.hm_besselk_rt(nu, float(x))          # reciprocal-table variant used on the device
            worst = max(worst, abs(got - want) / want)
    assert worst <= 5e-13
    assert hm.hm_besselk(0.8, 800.0) == 0.0


def test_besselk_chebyshev_fit_matches_continued_fraction(hm):
    """The table kernels evaluate K_nu for x >= 2 from a per-order Chebyshev fit of the scaled continued-fraction values
    (iak3d_math.h: BkFit): it must reproduce the continued fraction itself to a few ulp over the whole order range the
    reference accepts (0.05 <= nu <= 20, iaCovMatern.R:354) and agree with SciPy like the original does."""
    rng = np.random.default_rng(11)
    for nu in (0.05, 0.31, 0.5000001, 0.8, 1.0, 1.49, 1.7, 2.2, 3.3, 5.0, 7.3, 12.1, 19.99, 20.0):
        xs = np.concatenate([np.exp(rng.uniform(np.log(2.0), np.log(700.0), 1500)),
                             [2.0, 2.0000001, 4.0, 8.0, 16.0, 15.9999, 16.0001, 700.0, 705.3]])
        fit = np.array([hm.hm_besselk_fit(nu, x) for x in xs]); cf = np.array([hm.hm_besselk(nu, x) for x in xs])
        m = cf > 1e-300
        assert np.max(np.abs(fit[m] / cf[m] - 1.0)) <= 2e-14, nu
        assert np.max(np.abs(fit[m] / ssp.kv(nu, xs[m]) - 1.0)) <= 5e-13, nu
        assert hm.hm_besselk_fit(nu, 706.0) == 0.0                    # R's xmax_BESS_K cut-off is kept
    # below x = 2 nothing changes (Temme's series)
    for x in (1e-3, 0.5, 1.9999):
        assert hm.hm_besselk_fit(0.8, x) == hm.hm_besselk_rt(0.8, x)


def test_phix_against_oracle(hm):
    D = np.concatenate([[0.0, 1e-9, 1e-3], np.linspace(0.01, 90.0, 200), [1.3e100]])
    for nu in (0.5, 0.8, 1.5, 2.5, 3.3, 11.0):
        want = orc.spatialCovIAK3D(D, [1.0, 10.0, nu], "matern")
        got = np.array([hm.hm_phix(float(d), 2, 10.0, nu) for d in D])
        ok = want > 1e-280
        assert np.max(np.abs(got[ok] - want[ok]) / want[ok]) <= 1e-10
        assert np.all(got[~ok] <= 1e-280) and not np.any(np.isnan(got))
    want = orc.spatialCovIAK3D(D, [1.0, 25.0, None], "spherical")
    got = np.array([hm.hm_phix(float(d), 1, 25.0, 0.0) for d in D])
    assert np.max(np.abs(got - want)) <= 1e-15 and got[-1] == 0.0
