"""GPU parity: site-specific random effects (siteIDData given: listStructs4ssre, iaCovMatern.R:440-460, 645-663; added to the
covariances at fitIAK3D.R:1091-1101 and 1233-1240) through the C ABI against the oracle.  1e-10 on covariance entries, 1e-8 on
likelihood and kriging output."""
import numpy as np
import pytest

import iak3d_b200 as iak
from iak3d_b200 import synth
from oracle import iak3d_oracle as orc
from helpers import PARBNDS, case_pars, scaled_err

pytestmark = pytest.mark.gpu
COLS = [0, 5]                       # colnames4ssre: the intercept and one depth column


def _case(n, seed):
    ds = synth.make_dataset(n, seed)
    site = np.array([f"s{int(p) // 3}" for p in ds["prof"]])          # a site = three neighbouring profiles
    return ds, site


def _pars(kind="matern0.5", nonstat=False):
    P, modelx, types, cmeOpt = case_pars(kind, nonstat, 0.5)
    P = dict(P); P["cssre"] = np.array([0.07, 0.011])
    return P, modelx, types, cmeOpt


@pytest.mark.parametrize("kind,nonstat,n", [("matern0.5", False, 700), ("edgeroi", True, 1300)])
def test_setCIAK3D_with_ssre(ctx, kind, nonstat, n):
    ds, site = _case(n, 31)
    P, modelx, types, cmeOpt = _pars(kind, nonstat)
    sm = iak.setupIAK3D(ds["xData"], ds["dIData"], ctx=ctx, siteIDData=site, XData=ds["XData"], colnames4ssre=COLS)
    got = iak.setCIAK3D(P, modelx, *types, cmeOpt, sm)
    osm = orc.setupIAK3D(ds["xData"], ds["dIData"], siteIDData=site, XData=ds["XData"], cols4ssre=COLS)
    Cw, s2w = orc.setCIAK3D(P, modelx, *types, cmeOpt, osm)
    assert scaled_err(got["C"], Cw) <= 1e-10
    assert np.isnan(got["sigma2Vec"]).all() and np.isnan(s2w).all()                # fitIAK3D.R:1092
    # the term really is there: without it the matrices differ
    C0, _ = orc.setCIAK3D({k: v for k, v in P.items() if k != "cssre"}, modelx, *types, cmeOpt, orc.setupIAK3D(ds["xData"], ds["dIData"]))
    assert scaled_err(C0, Cw) > 1e-4
    # the matrix the likelihood factorises (cov_factor_kernel + the ssre pass)
    sm.set_W(np.column_stack([ds["XData"], ds["zData"]]))
    A = iak.api.factor_build_fetch(P, modelx, *types, cmeOpt, sm)
    assert scaled_err(A, Cw) <= 1e-10
    A2 = iak.api.factor_build_fetch(P, modelx, *types, cmeOpt, sm, batched=True)
    assert np.array_equal(A, A2)


def test_setCIAK3D2_with_ssre(ctx):
    ds, site = _case(600, 32)
    P, modelx, types, cmeOpt = _pars()
    rng = np.random.default_rng(1)
    x1 = np.vstack([rng.uniform(0, 100, (80, 2)), ds["xData"][:40]])
    d1 = np.repeat(synth.GSM_INTERVALS[2:3], 120, axis=0)
    X1 = np.column_stack([np.ones(120), rng.standard_normal((120, ds["XData"].shape[1] - 1))])
    site1 = np.concatenate([np.array([f"new{k % 7}" for k in range(80)]), site[:40]])     # unseen sites and sites of the data
    sm2 = iak.setupIAK3D2(x1, d1, ds["xData"], ds["dIData"], ctx=ctx, siteIDData=site1, XData=X1, siteIDData2=site, XData2=ds["XData"],
                          colnames4ssre=COLS)
    got = iak.setCIAK3D2(P, modelx, *types, cmeOpt, sm2)
    osm2 = orc.setupIAK3D2(x1, d1, ds["xData"], ds["dIData"], siteIDData=site1, XData=X1, siteIDData2=site, XData2=ds["XData"], cols4ssre=COLS)
    want = orc.setCIAK3D2(P, modelx, *types, cmeOpt, osm2)
    assert scaled_err(got, want) <= 1e-10


@pytest.mark.parametrize("useReml", [True, False])
def test_nll_gradient_and_kriging_with_ssre(ctx, useReml):
    ds, site = _case(900, 33)
    P, modelx, types, cmeOpt = _pars("edgeroi")
    sm = iak.setupIAK3D(ds["xData"], ds["dIData"], ctx=ctx, siteIDData=site, XData=ds["XData"], colnames4ssre=COLS)
    osm = orc.setupIAK3D(ds["xData"], ds["dIData"], siteIDData=site, XData=ds["XData"], cols4ssre=COLS)
    kw = dict(modelx=modelx, nud=0.5, sdfdType_cd1=types[0], sdfdType_cxd0=types[1], sdfdType_cxd1=types[2], cmeOpt=cmeOpt,
              setupMats=sm, parBnds=PARBNDS, useReml=useReml)
    got = iak.nllIAK3D(None, ds["zData"], ds["XData"], rtnAll=True, attachBigMats=True, parsBTfmd=dict(P), **kw)
    want = orc.nllIAK3D(None, ds["zData"], ds["XData"], modelx, 0.5, *types, cmeOpt, True, osm, PARBNDS, useReml=useReml, rtnAll=True,
                        parsBTfmd=dict(P))
    assert abs(got["nll"] - want["nll"]) <= 1e-8 * abs(want["nll"])
    assert np.allclose(got["parsBTfmd"]["cssre"], want["parsBTfmd"]["cssre"], rtol=1e-8)          # rescaled by cxdhat (:863)
    # the optimiser's vector: the site variances are its last entries (fitIAK3D.R:1413)
    from test_gpu_cl_sharded import _edgeroi_vector
    pars = np.concatenate([_edgeroi_vector(P), np.log(P["cssre"])])
    v = iak.nllIAK3D(pars, ds["zData"], ds["XData"], prodSum=True, **kw)
    vw = orc.nllIAK3D(pars, ds["zData"], ds["XData"], modelx, 0.5, *types, cmeOpt, True, osm, PARBNDS, useReml=useReml)
    assert abs(v - vw) <= 1e-8 * abs(vw)
    g = iak.gradnllIAK3D(pars, ds["zData"], ds["XData"], prodSum=True, analytic="auto", **kw)
    assert g.shape == (8,) and np.all(np.isfinite(g)) and np.any(g[-2:] != 0)
    # kriging: map supports at sites of the data and at unseen sites
    rng = np.random.default_rng(4)
    xMap = np.vstack([rng.uniform(0, 100, (150, 2)), ds["xData"][::30]])
    siteMap = np.concatenate([np.array([f"m{k % 11}" for k in range(150)]), site[::30]])
    dIMap = synth.GSM_INTERVALS[:3]
    Xfn = lambda i, idx: synth.grid_design(np.clip(xMap[idx], 0, 100), dIMap[i], 3)
    pg = iak.predictIAK3D(xMap, dIMap, Xfn, got, siteIDMap=siteMap)
    want.update(xData=ds["xData"], dIData=ds["dIData"], siteIDData=site, cols4ssre=COLS, XData=ds["XData"])
    zw, vw_, _, _ = orc.predictIAK3D(xMap, dIMap, Xfn, want, modelx, types, cmeOpt, siteIDMap=siteMap)
    assert scaled_err(pg["zMap"], zw) <= 1e-8
    assert scaled_err(pg["vMap"], vw_) <= 1e-8
    with pytest.raises(ValueError):
        iak.predictIAK3D(xMap, dIMap, Xfn, got)             # siteIDMap is required (predictIAK3D.R:11)
