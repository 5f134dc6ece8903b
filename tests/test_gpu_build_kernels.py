"""GPU parity of the HOT build kernels at north_star's 1e-10 on covariance entries: the matrix the likelihood factorises
(`cov_factor_kernel` / `cov_factor_kernel_one` into the blocked factor buffer -- interior fast path with its own term order
and explicit fma, both the one-table and the generic path, diagonal / border tiles, identity padding aside) and the kriging
panel (`phix_rect_kernel` + `cov_panel_kernel`, and the odd-row-count fallback), fetched back through the C ABI
(`iak3d_factor_build_fetch`, `iak3d_predict_panel_fetch`) and compared entry by entry with the oracle's setCIAK3D
(fitIAK3D.R:955-1112) and setCIAK3D2 (:1117-1252)."""
import numpy as np
import pytest

import iak3d_b200 as iak
from iak3d_b200 import api, synth
from oracle import iak3d_oracle as orc
from helpers import MODEL_CASES, PARBNDS, SPLINE_CASES, SPLINE_KNOTS, case_pars, spline_case_pars
from test_gpu_cov import cov_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [700, 2049])
@pytest.mark.parametrize("label,kind,nonstat,nud", MODEL_CASES, ids=[c[0] for c in MODEL_CASES])
def test_factor_build_vs_oracle(ctx, label, kind, nonstat, nud, n):
    ds = synth.make_dataset(n, 41 + n)
    P, modelx, types, cmeOpt = case_pars(kind, nonstat, nud)
    W = np.column_stack([ds["XData"], ds["zData"]])
    sm = iak.SetupMats(ds["xData"], ds["dIData"], W=W, ctx=ctx)
    want, _ = orc.setCIAK3D(P, modelx, *types, cmeOpt, orc.setupIAK3D(ds["xData"], ds["dIData"]))
    want = np.asarray(want)
    for batched in (False, True):          # cov_factor_kernel_one (job in the constant bank) and cov_factor_kernel (job array)
        got, Wr = api.factor_build_fetch(P, modelx, *types, cmeOpt, sm, batched=batched, rtnW=True)
        assert cov_err(got, want) <= 1.0, (label, n, batched)
        assert np.array_equal(Wr, W.T)     # the bordered rows are W' exactly
    sm.close()


@pytest.mark.parametrize("label,kind,types,coefs", SPLINE_CASES, ids=[c[0] for c in SPLINE_CASES])
def test_factor_build_spline_vs_oracle(ctx, label, kind, types, coefs):
    n = 900
    ds = synth.make_dataset(n, 77)
    P, modelx, types, cmeOpt = spline_case_pars(kind, types, coefs)
    sm = iak.SetupMats(ds["xData"], ds["dIData"], W=np.column_stack([ds["XData"], ds["zData"]]), ctx=ctx, sdfdType_cd1=types[0],
                       sdfdType_cxd0=types[1], sdfdType_cxd1=types[2], sdfdKnots=SPLINE_KNOTS)
    osm = orc.setupIAK3D(ds["xData"], ds["dIData"], sdfdTypes=types, sdfdKnots=SPLINE_KNOTS)
    want, _ = orc.setCIAK3D(P, modelx, *types, cmeOpt, osm)
    got = api.factor_build_fetch(P, modelx, *types, cmeOpt, sm)
    assert cov_err(got, np.asarray(want)) <= 1.0, label
    sm.close()


def test_factor_build_invalid_parameters(ctx):
    ds = synth.make_dataset(300, 5)
    P, modelx, types, cmeOpt = case_pars("matern0.5", False, 0.5)
    P = dict(P); P["cx1"] = float("inf")
    sm = iak.SetupMats(ds["xData"], ds["dIData"], ctx=ctx)
    assert api.factor_build_fetch(P, modelx, *types, cmeOpt, sm) is None
    sm.close()


@pytest.mark.parametrize("m", [640, 333])      # even -> cov_panel_kernel; odd -> the cov_rect_kernel fallback of launch_cov_panel
@pytest.mark.parametrize("kind,nonstat,nud", [("matern0.5", False, 0.5), ("matern0.8", True, 1.5), ("edgeroi", False, 0.5),
                                              ("edgeroi", True, 2.5), ("wendland", False, 0.5), ("nugget", False, 0.5)])
def test_kriging_panel_vs_oracle(ctx, kind, nonstat, nud, m):
    ds = synth.make_dataset(700, 15)
    P, modelx, types, cmeOpt = case_pars(kind, nonstat, nud)
    sm = iak.setupIAK3D(ds["xData"], ds["dIData"], ctx=ctx)
    fit = iak.nllIAK3D(None, ds["zData"], ds["XData"], modelx=modelx, nud=nud, sdfdType_cd1=types[0], sdfdType_cxd0=types[1],
                       sdfdType_cxd1=types[2], cmeOpt=cmeOpt, setupMats=sm, parBnds=PARBNDS, rtnAll=True, attachBigMats=True, parsBTfmd=P)
    rng = np.random.default_rng(3)
    xMap = np.vstack([rng.uniform(0, 100, (m - 40, 2)), ds["xData"][::18][:39], [[9e99, 9e99]]])   # + data locations + the distant point
    dIMap = np.vstack([np.repeat(synth.GSM_INTERVALS[1:2], m // 2, axis=0), np.repeat(synth.GSM_INTERVALS[4:5], m - m // 2, axis=0)])
    k = fit["fit"]
    k.stage(xMap, dIMap, np.zeros((m, ds["XData"].shape[1])))
    got = k.panel_fetch(0)
    osm2 = orc.setupIAK3D2(xMap, dIMap, ds["xData"], ds["dIData"])
    want = np.asarray(orc.setCIAK3D2(fit["parsBTfmd"], modelx, *types, cmeOpt, osm2))
    assert got.shape == want.shape
    assert not np.any(np.isnan(got))
    # the "distant" point (9E99, 9E99) of fitIAK3D.R:606: every horizontal correlation is exactly 0 there, so
    # the row is the cd1 term alone.  The oracle's own value at that separation can be NaN (Wendland: Inf * 0, as in R).
    Pz = dict(fit["parsBTfmd"]); Pz["cx1"] = 0.0; Pz["cxd1"] = 0.0
    far = np.asarray(orc.setCIAK3D2(Pz, "spherical" if modelx != "nugget" else modelx, *types, cmeOpt, dict(osm2))) if modelx != "nugget" else want
    assert cov_err(got[:-1], want[:-1]) <= 1.0
    far = np.asarray(far)[-1]
    assert np.array_equal(got[-1], far) if not np.any(far) else cov_err(got[-1], far) <= 1.0
