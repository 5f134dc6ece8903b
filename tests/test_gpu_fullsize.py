"""GPU parity at BASELINE.json's sizes through size-independent properties (the oracle is too slow / too large there):
configs[1] (n = 5000 single block) and a 4000-row composite-likelihood pair of configs[3]; configs[2] (n = 40 000) is
checked against LAPACK by scripts/big_runs.py (profiles/s40_n40000_r01.log) because the host side alone takes a minute."""
import numpy as np
import pytest

import iak3d_b200 as iak
from iak3d_b200 import synth
from helpers import PARBNDS, case_pars, scaled_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def s5(ctx):
    ds = synth.make_dataset(5000, synth.SEEDS["S5"])
    sm = iak.setupIAK3D(ds["xData"], ds["dIData"], ctx=ctx)
    return ds, sm


def _kw(sm, kind="matern0.5", nonstat=False, **over):
    P, modelx, types, cmeOpt = case_pars(kind, nonstat, 0.5)
    P = dict(P); P.update(over)
    return dict(modelx=modelx, nud=0.5, sdfdType_cd1=types[0], sdfdType_cxd0=types[1], sdfdType_cxd1=types[2], cmeOpt=cmeOpt,
                setupMats=sm, parBnds=PARBNDS, parsBTfmd=P), P, modelx, types, cmeOpt


def test_s5_solve_identity_and_determinism(s5):
    """A (A^-1 W) == W with A rebuilt by the rectangular kernel (another code path than the factor build), WiAW == W' A^-1 W,
    repeated evaluations bit-identical, batched == single."""
    ds, sm = s5
    kw, P, modelx, types, cmeOpt = _kw(sm)
    t = iak.nllIAK3D(None, ds["zData"], ds["XData"], forCompLik=True, attachBigMats=True, **kw)
    W = np.column_stack([ds["XData"], ds["zData"]])
    assert scaled_err(t["A"] @ t["iAW"], W) <= 1e-9
    assert scaled_err(W.T @ t["iAW"], t["WiAW"]) <= 1e-10
    sign, lnd = np.linalg.slogdet(t["A"])
    assert sign > 0 and abs(lnd - t["lndetA"]) <= 1e-10 * abs(lnd)
    a = iak.nllIAK3D(None, ds["zData"], ds["XData"], **kw); b = iak.nllIAK3D(None, ds["zData"], ds["XData"], **kw)
    assert a == b
    pars = np.array([0.2, -1.5, 0.1, -2.3, -1.4, -1.9, 0.4, -3.0])
    g, f0 = iak.gradnllIAK3D(pars, ds["zData"], ds["XData"], modelx="matern", nud=0.5, cmeOpt=1, setupMats=sm, parBnds=PARBNDS, rtn_nll=True)
    assert f0 == iak.nllIAK3D(pars, ds["zData"], ds["XData"], modelx="matern", nud=0.5, cmeOpt=1, setupMats=sm, parBnds=PARBNDS)
    assert np.all(np.isfinite(g))


def test_s5_scaling_invariance_of_the_profiled_likelihood(s5):
    """REML with the variance profiled out: multiplying every variance parameter by c changes cxdhat by 1/c and nothing else
    (fitIAK3D.R:849-863)."""
    ds, sm = s5
    kw, P, *_ = _kw(sm)
    f1 = iak.nllIAK3D(None, ds["zData"], ds["XData"], rtnAll=True, **kw)
    P2 = dict(P)
    for k in ("cx0", "cx1", "cd1", "cxd0", "cxd1", "cme"):
        P2[k] = 3.7 * P[k]
    kw2 = dict(kw); kw2["parsBTfmd"] = P2
    f2 = iak.nllIAK3D(None, ds["zData"], ds["XData"], rtnAll=True, **kw2)
    assert abs(f1["nll"] - f2["nll"]) <= 1e-9 * abs(f1["nll"])
    assert abs(f1["cxdhat"] - 3.7 * f2["cxdhat"]) <= 1e-9 * f1["cxdhat"]
    assert scaled_err(f1["betahat"], f2["betahat"]) <= 1e-9


def test_s5_kriging_round_trip_and_linearity(s5):
    """Without measurement error kriging at the data supports returns the data with zero variance; predictions are linear in z
    (two fits with z and 2 z + X b: predictions transform the same way, variances scale by 4)."""
    ds, sm = s5
    kw, P, *_ = _kw(sm, cme=0.0)
    kw["cmeOpt"] = 0
    fit = iak.nllIAK3D(None, ds["zData"], ds["XData"], rtnAll=True, attachBigMats=True, **kw)
    sel = np.arange(0, 5000, 11)
    z, v = fit["fit"].predict(ds["xData"][sel], ds["dIData"][sel], ds["XData"][sel])
    assert np.max(np.abs(z - ds["zData"][sel])) <= 1e-6 * np.max(np.abs(ds["zData"]))
    assert np.max(np.abs(v)) <= 1e-6 * fit["cxdhat"]
    rng = np.random.default_rng(3)
    xMap = rng.uniform(0, 100, (4000, 2)); dI = np.tile(synth.GSM_INTERVALS, (667, 1))[:4000]
    XM = synth.grid_design(xMap, dI[0], 3)          # any design matrix will do for the linearity property
    z1, v1 = fit["fit"].predict(xMap, dI, XM)
    b = rng.standard_normal(ds["XData"].shape[1])
    z2data = 2.0 * ds["zData"] + ds["XData"] @ b
    fit2 = iak.nllIAK3D(None, z2data, ds["XData"], rtnAll=True, attachBigMats=True, **kw)
    z2, v2 = fit2["fit"].predict(xMap, dI, XM)
    assert scaled_err(z2, 2.0 * z1 + XM @ b) <= 1e-8
    assert scaled_err(v2, 4.0 * v1) <= 1e-8


def test_cl_pair_of_configs3_size_equals_exact(ctx):
    """One 4000-row unit of the composite likelihood (a block pair of configs[3]): compLikOptn = 4 with a single block is
    the exact likelihood, bit for bit, through the batched entry point."""
    ds = synth.make_dataset(4000, synth.SEEDS["CL200"])
    kw, P, modelx, types, cmeOpt = _kw(iak.setupIAK3D(ds["xData"], ds["dIData"], ctx=ctx), "edgeroi")
    exact = iak.nllIAK3D(None, ds["zData"], ds["XData"], **kw)
    blk = dict(compLikOptn=4, listBlocks=[np.arange(4000)], subsetPairsAdj=np.zeros((0, 2), int), subsetPairsNonadj=np.zeros((0, 2), int))
    cl = iak.setupIAK3D_CL(ds["xData"], ds["dIData"], ds["zData"], ds["XData"], blk, ctx=ctx)
    got = iak.nllIAK3D_CL(None, ds["zData"], ds["XData"], modelx, 0.5, *types, cmeOpt, True, PARBNDS, True, cl, parsBTfmd=P)
    assert got == exact


def test_edgeroi_shaped_example_driver_runs(ctx):
    """examples/eg_edgeroi_shaped.py (the flow of egEdgeroi.R: optimIt fit -> predictIAK3D -> hold-out validation) end to end
    on the GPU at a reduced size; the fitted objective must not be worse than the oracle's value at the same parameters."""
    import json
    import os
    import subprocess
    import sys
    from oracle import iak3d_oracle as orc
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "examples", "eg_edgeroi_shaped.py"), "--impl", "gpu", "--n", "350", "--grid", "8"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert np.isfinite(d["nll"]) and d["nll"] < 9e98 and d["fn_evals"] > 50 and d["gr_evals"] > 3
    assert 0.5 <= d["holdout_in_pi90"] <= 1.0 and d["map_var_mean"] > 0
    ds = synth.make_dataset(350 + 250, synth.SEEDS["E"])
    fitr = ds["prof"] < ds["prof"].max() - 59
    want = orc.nllIAK3D(np.array(d["pars"]), ds["zData"][fitr], ds["XData"][fitr], "spherical", 0.5, -9, 0, 0, 1, True,
                        orc.setupIAK3D(ds["xData"][fitr], ds["dIData"][fitr]), PARBNDS)
    assert abs(want - d["nll"]) <= 1e-8 * abs(want)
