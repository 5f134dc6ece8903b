"""CPU: the library's readParsIAK3D (iak3d_read_pars, csrc/tail.cu; fitIAK3D.R:1332-1417) against the host mirror
(api.readParsIAK3D + make_pars) and against the oracle's restatement, over every model switch that changes the consumption
order of the vector."""
import ctypes as C
import itertools

import numpy as np
import pytest

from iak3d_b200 import api
from iak3d_b200._lib import MAX_SPL, Model, Pars, load, make_model, make_pars
from oracle import iak3d_oracle as orc
from helpers import PARBNDS


def _npar(modelx, types, cmeOpt, prodSum, lnT):
    n = 0
    if modelx != "nugget":
        n += 1 + (modelx != "spherical")
    n += 1
    if prodSum:
        n += 1 + (modelx != "nugget")
    n += types[0] != -9
    n += (0 if modelx == "nugget" else 1) if not lnT else (1 + (modelx != "nugget"))
    for t in types:
        n += 2 if t == -1 else (t if t > 0 else 0)
    return n + (cmeOpt == 1)


CASES = [(m, t, c, ps, ln) for m in ("matern", "spherical", "wendland", "nugget")
         for t in ((0, 0, 0), (-9, 0, 0), (-1, 0, -1), (0, -1, 0), (3, 0, 2), (-9, -1, 4))
         for c in (0, 1) for ps in (True, False) for ln in (False, True)]


@pytest.mark.parametrize("modelx,types,cmeOpt,prodSum,lnT", CASES[::3])
def test_read_pars_matches_host_mirror_and_oracle(modelx, types, cmeOpt, prodSum, lnT):
    lib = load()
    rng = np.random.default_rng(1000 * CASES.index((modelx, types, cmeOpt, prodSum, lnT)) + 7)
    for extra in (0, 2):                                  # trailing entries are the site-specific random-effect variances
        v = rng.standard_normal(_npar(modelx, types, cmeOpt, prodSum, lnT) + extra) * 1.5
        M = make_model(modelx, 0.5, *types, cmeOpt, prodSum, PARBNDS, lnT)
        got = Pars()
        assert lib.iak3d_read_pars(C.byref(M), v.ctypes.data_as(C.POINTER(C.c_double)), v.size, C.byref(got)) == 0
        pb = api.readParsIAK3D(v, modelx, 0.5, *types, cmeOpt, prodSum, PARBNDS, lnTfmdData=lnT)
        want = make_pars(pb, modelx, *types, cmeOpt)
        ob = orc.readParsIAK3D(v, modelx, 0.5, *types, cmeOpt, prodSum, PARBNDS, lnTfmdData=lnT)
        for f in ("ax", "nux", "ad", "nud", "cx0", "cx1", "cd1", "cxd0", "cxd1", "cme"):
            a, b = getattr(got, f), getattr(want, f)
            assert (np.isnan(a) and np.isnan(b)) or abs(a - b) <= 4e-16 * max(abs(b), 1e-300), (f, a, b)
            o = float(np.ravel(ob[f])[0]) if np.size(ob[f]) else float("nan")
            assert (np.isnan(a) and np.isnan(o)) or abs(a - o) <= 4e-16 * max(abs(o), 1e-300), (f, a, o)
        assert (got.modelx, got.cmeOpt, list(got.sdfdType), got.quirk_cd1_unscaled, got.quirk_scale, got.nssre) == \
               (want.modelx, want.cmeOpt, list(want.sdfdType), want.quirk_cd1_unscaled, want.quirk_scale, want.nssre)
        for c in range(3):
            # (1 - exp(v0)) / (1 - exp(-tau2 maxd)): the two exp() implementations (libm / NumPy) differ by an ulp, the
            # differences of 1 amplify it
            assert np.allclose(list(got.sdfdPars[c]), list(want.sdfdPars[c]), rtol=1e-13, atol=0)
            assert np.allclose(list(got.sdfdSpl[c]), list(want.sdfdSpl[c]), rtol=0, atol=0)
        assert np.allclose(list(got.cssre), list(want.cssre), rtol=4e-16, atol=0)


def test_read_pars_rejects_what_the_reference_stops_on():
    lib = load()
    M = make_model("matern", 0.5, 0, 0, 0, 1, True, PARBNDS)
    v = np.zeros(3)                                        # too short
    assert lib.iak3d_read_pars(C.byref(M), v.ctypes.data_as(C.POINTER(C.c_double)), v.size, C.byref(Pars())) != 0
    M = make_model("matern", 0.5, -5, 0, 0, 1, True, PARBNDS)     # "sdfd option not yet programmed"
    v = np.zeros(9)
    assert lib.iak3d_read_pars(C.byref(M), v.ctypes.data_as(C.POINTER(C.c_double)), v.size, C.byref(Pars())) != 0
    M = make_model("matern", 0.5, MAX_SPL + 1, 0, 0, 1, True, PARBNDS)
    v = np.zeros(40)
    assert lib.iak3d_read_pars(C.byref(M), v.ctypes.data_as(C.POINTER(C.c_double)), v.size, C.byref(Pars())) != 0


def test_read_pars_overflow_goes_to_infinity_like_R():
    """exp() of a huge log-variance is Inf in R (then parsOK = FALSE -> badnll); never an exception."""
    lib = load()
    M = make_model("spherical", 0.5, -9, 0, 0, 1, True, PARBNDS)
    v = np.array([0.3, -0.2, 800.0, -0.7, 0.4, -2.0])
    got = Pars()
    assert lib.iak3d_read_pars(C.byref(M), v.ctypes.data_as(C.POINTER(C.c_double)), v.size, C.byref(got)) == 0
    assert np.isinf(got.cx0) and np.isfinite(got.cx1)
    v = np.array([-900.0, 900.0, 0.0, 0.0, -900.0, 0.0])      # logits at the ends of their ranges
    assert lib.iak3d_read_pars(C.byref(M), v.ctypes.data_as(C.POINTER(C.c_double)), v.size, C.byref(got)) == 0
    assert got.ax == PARBNDS["axBnds"][0] and got.ad == PARBNDS["adBnds"][1] and got.cxd1 == PARBNDS["sxd1Bnds"][0]
