"""GPU parity: covariance kernels through the C ABI against the oracle (bit-for-bit inputs, seeded).
Tolerance (BASELINE.json north_star): max relative error <= 1e-10 on covariance entries."""
import json
import os

import numpy as np
import pytest

import iak3d_b200 as iak
from iak3d_b200 import synth
from oracle import iak3d_oracle as orc
from helpers import MODEL_CASES, case_pars, relerr, scaled_err


def cov_err(got, want):
    """Relative error with an ulp-level absolute floor: |got - want| / (|want| + 1e-5 max|want|) would hide errors,
    so instead entries must satisfy |got - want| <= TOL |want| + 2e-15 max|want|.  The floor covers the entries the
    reference itself forms by cancellation of O(1) terms (spherical / Wendland just inside the range, where
    1 - 1.5 r + 0.5 r^3 ~ 1e-6 carries 1e-10 relative noise from the last bit of r^3).  Returns the worst ratio of
    |got - want| to that bound (<= 1 passes)."""
    got = np.asarray(got, float); want = np.asarray(want, float)
    if want.size == 0:
        return 0.0
    bound = TOL_COV * np.abs(want) + 2e-15 * np.max(np.abs(want))
    return float(np.max(np.abs(got - want) / bound))

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
TOL_COV = 1e-10


def test_iacov_known_answers(ctx):
    with open(os.path.join(HERE, "golden", "iacov_kat.json")) as f:
        kat = json.load(f)
    for case in kat["cases"]:
        for (nud, st), want in zip(kat["columns"], case["values"]):
            got, _ = iak.iaCovMatern2([case["I1"]], [case["I2"]], kat["ad"], nud, kat["sdfdPars"], st, ctx=ctx)
            assert abs(got[0, 0] - want) <= 1e-13 * abs(want), (case["case"], nud, st)


@pytest.mark.parametrize("nud", [0.5, 1.5, 2.5])
@pytest.mark.parametrize("st", [0, -1])
def test_iacov_rectangular_vs_oracle(ctx, nud, st):
    rng = np.random.default_rng(17)
    for ad in (0.05, 0.35, 1.1):
        t1 = np.round(rng.uniform(0, 1.8, 157), 2); t2 = np.round(rng.uniform(0, 1.8, 93), 2)
        d1 = np.column_stack([t1, np.round(t1 + rng.choice([0.01, 0.05, 0.1, 0.3], 157), 2)])
        d2 = np.column_stack([t2, np.round(t2 + rng.choice([0.02, 0.1, 0.5], 93), 2)])
        got, av = iak.iaCovMatern2(d1, d2, ad, nud, (0.6, 1.7), st, ctx=ctx)
        want = orc.iaCovMatern2(d1, d2, ad, nud, (0.6, 1.7), st)
        assert relerr(got, want) <= TOL_COV
        assert relerr(av, orc.avVarVec(d1, orc.nsMaternParams(ad, nud, (0.6, 1.7), st))) <= 1e-13


def test_iacov_invalid_nud(ctx):
    with pytest.raises(ValueError):
        iak.iaCovMatern2([[0, 1]], [[0, 1]], 0.3, 1.0, (0, 0), 0, ctx=ctx)


@pytest.mark.parametrize("model,a,nu", [("spherical", 25.0, None), ("matern", 10.0, 0.5), ("matern", 10.0, 1.5), ("matern", 10.0, 2.5),
                                        ("matern", 10.0, 0.8), ("matern", 3.0, 0.05), ("matern", 40.0, 7.3), ("matern", 10.0, 20.0),
                                        ("wendland", 30.0, 2.4), ("wendland", 30.0, 1.0), ("wendland", 12.0, 4.7)])
def test_spatial_cov_vs_oracle(ctx, model, a, nu):
    rng = np.random.default_rng(23)
    D = np.concatenate([[0.0, 1e-9, 1e-3, a, a * (1 + 1e-12), 1.3e100], rng.uniform(0, 140.0, 5000)])
    want = orc.spatialCovIAK3D(D, [1.0, a, nu], model)
    got = iak.spatialCovIAK3D(D, [1.0, a, nu], model, ctx=ctx)
    assert not np.any(np.isnan(got))
    big = want > 1e-250
    assert cov_err(got[big], want[big]) <= 1.0
    assert np.all(np.abs(got[~big]) <= 1e-250)
    assert got[-1 - 5000] == 0.0            # the "distant" point: exactly 0


def test_spatial_cov_invalid_returns_na(ctx):
    assert iak.spatialCovIAK3D(np.ones(3), [1.0, -2.0, 0.5], "matern", ctx=ctx) is None
    assert iak.spatialCovIAK3D(np.ones(3), [1.0, 2.0, 21.0], "matern", ctx=ctx) is None
    assert iak.spatialCovIAK3D(np.ones(3), [1.0, 2.0, 0.5], "wendland", ctx=ctx) is None


def test_setup_ids_follow_first_occurrence(ctx):
    ds = synth.make_dataset(700, 2)
    sm = iak.setupIAK3D(ds["xData"], ds["dIData"], ctx=ctx)
    o = orc.setupIAK3D(ds["xData"], ds["dIData"])
    xid, did = sm.ids()
    assert np.array_equal(xid, o["xid"]) and np.array_equal(did, o["did"])
    assert sm.nxU == o["xU"].shape[0] and sm.ndIU == o["dIU"].shape[0]


@pytest.mark.parametrize("label,kind,nonstat,nud", MODEL_CASES)
@pytest.mark.parametrize("n", [1, 63, 700])
def test_setC_vs_oracle(ctx, label, kind, nonstat, nud, n):
    ds = synth.make_dataset(n, 100 + n)
    P, modelx, types, cmeOpt = case_pars(kind, nonstat, nud)
    sm = iak.setupIAK3D(ds["xData"], ds["dIData"], ctx=ctx)
    got = iak.setCIAK3D(P, modelx, *types, cmeOpt, sm)
    Cw, s2w = orc.setCIAK3D(P, modelx, *types, cmeOpt, orc.setupIAK3D(ds["xData"], ds["dIData"]))
    nz = np.abs(Cw) > 1e-200
    assert cov_err(got["C"][nz], Cw[nz]) <= 1.0
    assert np.all(got["C"][~nz] == 0.0)
    assert np.array_equal(got["C"], got["C"].T)
    assert relerr(got["sigma2Vec"], s2w) <= 1e-13


@pytest.mark.parametrize("label,kind,nonstat,nud", MODEL_CASES)
def test_setC2_vs_oracle(ctx, label, kind, nonstat, nud):
    ds = synth.make_dataset(500, 31)
    P, modelx, types, cmeOpt = case_pars(kind, nonstat, nud)
    # set 1: some rows are at data locations / data intervals (exact-equality nugget indicator, fitIAK3D.R:1138)
    rng = np.random.default_rng(1)
    x1 = np.vstack([ds["xData"][:40], rng.uniform(0, 100, (90, 2))])
    d1 = np.vstack([ds["dIData"][5:45], np.tile(synth.GSM_INTERVALS, (15, 1))])
    sm2 = iak.setupIAK3D2(x1, d1, ds["xData"], ds["dIData"], ctx=ctx)
    got = iak.setCIAK3D2(P, modelx, *types, cmeOpt, sm2)
    want = orc.setCIAK3D2(P, modelx, *types, cmeOpt, orc.setupIAK3D2(x1, d1, ds["xData"], ds["dIData"]))
    nz = np.abs(want) > 1e-200
    assert cov_err(got[nz], want[nz]) <= 1.0
    assert np.all(got[~nz] == 0.0)


def test_invalid_parameters_give_na(ctx):
    ds = synth.make_dataset(100, 3)
    sm = iak.setupIAK3D(ds["xData"], ds["dIData"], ctx=ctx)
    P, modelx, types, cmeOpt = case_pars("matern0.5", False, 0.5)
    bad = dict(P); bad["nux"] = 25.0
    assert iak.setCIAK3D(bad, modelx, *types, cmeOpt, sm)["C"] is None
    bad = dict(P); bad["cx1"] = float("inf")
    assert iak.setCIAK3D(bad, modelx, *types, cmeOpt, sm)["C"] is None
