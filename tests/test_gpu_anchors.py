"""The CUDA path through the C ABI against tests/golden/mpmath_anchors.json -- 40-digit mpmath evaluations of the model's
definitions that no repository code produced (scripts/make_mpmath_anchors.py; see tests/test_anchors.py).  Covers the
reference-facing covariance build, the factor build the likelihood uses, the likelihood terms + REML / ML tail, and the
kriging predictor.  Tolerances are north_star's: 1e-10 on covariance entries, 1e-8 on the rest."""
import numpy as np
import pytest

import iak3d_b200 as iak
from iak3d_b200 import api
from helpers import PARBNDS
from test_anchors import CASES, anchor_inputs, cov_err, lower

pytestmark = pytest.mark.gpu
IDS = [c["name"] for c in CASES]


@pytest.mark.parametrize("c", CASES, ids=IDS)
def test_cuda_setC_vs_mpmath(ctx, c):
    P, modelx, types, cmeOpt, x, dI, X, z = anchor_inputs(c)
    sm = iak.SetupMats(x, dI, W=np.column_stack([X, z]), ctx=ctx)
    A = lower(c)
    got = iak.setCIAK3D(P, modelx, *types, cmeOpt, sm)
    assert cov_err(got["C"], A) <= 1.0
    assert np.max(np.abs(got["sigma2Vec"] - np.array(c["sigma2Vec"])) / np.array(c["sigma2Vec"])) <= 1e-10
    for batched in (False, True):
        assert cov_err(api.factor_build_fetch(P, modelx, *types, cmeOpt, sm, batched=batched), A) <= 1.0
    sm.close()


@pytest.mark.parametrize("c", CASES, ids=IDS)
def test_cuda_nll_vs_mpmath(ctx, c):
    P, modelx, types, cmeOpt, x, dI, X, z = anchor_inputs(c)
    sm = iak.setupIAK3D(x, dI, ctx=ctx)
    kw = dict(modelx=modelx, nud=c["nud"], sdfdType_cd1=types[0], sdfdType_cxd0=types[1], sdfdType_cxd1=types[2], cmeOpt=cmeOpt,
              setupMats=sm, parBnds=PARBNDS, useReml=c["useReml"], parsBTfmd=P)
    t = iak.nllIAK3D(None, z, X, forCompLik=True, **kw)
    assert abs(t["lndetA"] - c["lndetA"]) <= 1e-8 * abs(c["lndetA"])
    W = np.array(c["WiAW"])
    assert np.max(np.abs(t["WiAW"] - W)) <= 1e-8 * np.max(np.abs(W))
    fit = iak.nllIAK3D(None, z, X, rtnAll=True, **kw)
    assert abs(fit["nll"] - c["nll"]) <= 1e-8 * abs(c["nll"])
    assert abs(fit["cxdhat"] - c["cxdhat"]) <= 1e-8 * c["cxdhat"]
    assert np.max(np.abs(np.asarray(fit["betahat"]).reshape(-1) - np.array(c["betahat"]))) <= 1e-8 * np.max(np.abs(c["betahat"]))
    assert abs(iak.nllIAK3D(None, z, X, **kw) - c["nll"]) <= 1e-8 * abs(c["nll"])
    sm.close()


@pytest.mark.parametrize("c", [c for c in CASES if c["predict"]], ids=[c["name"] for c in CASES if c["predict"]])
def test_cuda_kriging_vs_mpmath(ctx, c):
    P, modelx, types, cmeOpt, x, dI, X, z = anchor_inputs(c)
    sm = iak.setupIAK3D(x, dI, ctx=ctx)
    fit = iak.nllIAK3D(None, z, X, modelx=modelx, nud=c["nud"], sdfdType_cd1=types[0], sdfdType_cxd0=types[1], sdfdType_cxd1=types[2],
                       cmeOpt=cmeOpt, setupMats=sm, parBnds=PARBNDS, useReml=c["useReml"], rtnAll=True, attachBigMats=True, parsBTfmd=P)
    xm, dm, Xm = np.array(c["xMap"]), np.array(c["dIMap"]), np.array(c["XMap"])
    zs, vs = fit["fit"].predict(xm, dm, Xm)          # all supports in one panel (each row: its own location and interval)
    assert np.max(np.abs(zs - np.array(c["zMap"])) / np.abs(c["zMap"])) <= 1e-8
    assert np.max(np.abs(vs - np.array(c["vMap"])) / np.abs(c["vMap"])) <= 1e-8
    sm.close()
