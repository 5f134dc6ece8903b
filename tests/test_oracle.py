"""Pins the CPU oracle (oracle/iak3d_oracle.py): known-answer vectors, an independent quadrature of the defining
integral, the reference's own (dead) consistency checks, and LAPACK identities.  The reference ships no tests for
this path, so these are the anchors the oracle's header lists."""
import json
import math
import os

import numpy as np
import pytest
import scipy.integrate as sint
import scipy.special as ssp

from iak3d_b200 import synth
from oracle import iak3d_oracle as orc
from helpers import MODEL_CASES, PARBNDS, case_pars, relerr

HERE = os.path.dirname(os.path.abspath(__file__))


def _kat():
    with open(os.path.join(HERE, "golden", "iacov_kat.json")) as f:
        return json.load(f)


def test_iacov_known_answers():
    kat = _kat()
    for case in kat["cases"]:
        for (nud, st), want in zip(kat["columns"], case["values"]):
            got = orc.iaCovMatern2([case["I1"]], [case["I2"]], kat["ad"], nud, kat["sdfdPars"], st)[0, 0]
            assert abs(got - want) <= 2e-15 * abs(want) + 1e-16, (case["case"], nud, st, got, want)


def _point_cov(h, ad, nud):
    x = np.abs(h) / ad
    if nud == 0.5:
        return np.exp(-x)
    if nud == 1.5:
        return (1 + math.sqrt(3) * x) * np.exp(-math.sqrt(3) * x)
    return (1 + math.sqrt(5) * x + 5 * x * x / 3) * np.exp(-math.sqrt(5) * x)


@pytest.mark.parametrize("nud", [0.5, 1.5, 2.5])
@pytest.mark.parametrize("st", [0, -1])
def test_iacov_against_quadrature(nud, st):
    ad, sp = 0.35, (0.6, 1.7)
    s = (lambda d: 1.0) if st == 0 else (lambda d: (1 - sp[0]) + sp[0] * math.exp(-sp[1] * d))
    for (a, b, c, d) in [(0.0, 0.1, 0.3, 0.6), (0.0, 0.3, 0.15, 0.6), (0.0, 1.0, 0.2, 0.5), (0.05, 0.15, 0.05, 0.15)]:
        # split at the kink |x - y| = 0 by integrating x over [a,b] and y over the two sides
        f = lambda y, x: s(x) * s(y) * _point_cov(x - y, ad, nud)
        val = 0.0
        for (ylo, yhi) in [(c, d)]:
            val, _ = sint.dblquad(f, a, b, lambda x: ylo, lambda x: yhi, epsabs=1e-13, epsrel=1e-12)
        want = val / ((b - a) * (d - c))
        got = orc.iaCovMatern2([[a, b]], [[c, d]], ad, nud, sp, st)[0, 0]
        assert abs(got - want) <= 5e-8 * abs(want)   # dblquad across the kink is the limiting accuracy here


def test_iacov_square_equals_rectangular_and_symmetric():
    ds = synth.make_dataset(300, 11)
    sm = orc.setupIAK3D(ds["xData"], ds["dIData"])
    for nud in (0.5, 1.5, 2.5):
        for st, sp in ((0, (1, 0, 1)), (-1, (0.6, 1.7))):
            M, av = orc.iaCovMatern(sm["dIU"], 0.35, nud, sp, st)
            R = orc.iaCovMatern2(sm["dIU"], sm["dIU"], 0.35, nud, sp, st)
            assert relerr(M, R) <= 1e-13
            assert np.allclose(M, M.T, rtol=0, atol=0)
            # diag of the averaged covariance <= averaged variance (Cauchy-Schwarz) and both positive
            assert np.all(np.diag(M) > 0) and np.all(av > 0)


def test_iacov_discretisation_converges():
    """iaCovDsc (iaCovMatern.R:709-739), the reference's own cross-check, converges to the closed form."""
    dI = np.array([[0.0, 0.1], [0.1, 0.3], [0.3, 0.6], [0.05, 0.5]])
    M, _ = orc.iaCovMatern(dI, 0.35, 1.5, (0.6, 1.7), -1)
    e50 = np.max(np.abs(orc.iaCovDsc(dI, 0.35, 1.5, (0.6, 1.7), -1, 50) - M))
    e200 = np.max(np.abs(orc.iaCovDsc(dI, 0.35, 1.5, (0.6, 1.7), -1, 200) - M))
    assert e200 < e50 < 2e-3 and e200 < 2e-4


def test_matern_halfinteger_closed_forms():
    D = np.array([0.0, 1e-6, 0.3, 2.0, 17.0, 200.0])
    for nu, f in ((0.5, lambda x: np.exp(-x)), (1.5, lambda x: (1 + x) * np.exp(-x)),
                  (2.5, lambda x: (1 + x + x * x / 3) * np.exp(-x))):
        got = orc.spatialCovIAK3D(D, [1.0, 10.0, nu], "matern")
        want = f(math.sqrt(2 * nu) * D / 10.0)
        assert relerr(got[1:], want[1:]) < 5e-13 and got[0] == 1.0


def test_spatial_cov_edge_cases():
    # the "distant" prediction point (9E99, 9E99): exactly 0, never NaN
    D = np.array([math.sqrt(2.0) * 9e99])
    assert orc.spatialCovIAK3D(D, [1.0, 25.0, None], "spherical")[0] == 0.0
    assert orc.spatialCovIAK3D(D, [1.0, 25.0, 0.8], "matern")[0] == 0.0
    assert orc.spatialCovIAK3D(np.array([1.0]), [1.0, -1.0, 0.5], "matern") is None
    assert orc.spatialCovIAK3D(np.array([1.0]), [1.0, 1.0, 25.0], "matern") is None
    w = orc.spatialCovIAK3D(np.array([0.0, 15.0, 30.0, 31.0]), [1.0, 30.0, 2.4], "wendland")
    assert w[0] == pytest.approx(1.0, abs=1e-14) and w[2] == 0.0 and w[3] == 0.0 and 0 < w[1] < 1


@pytest.mark.parametrize("label,kind,nonstat,nud", MODEL_CASES)
def test_cross_equals_square_minus_nugget(label, kind, nonstat, nud):
    """setCIAK3D2(S,S) + cme I == setCIAK3D(S): the reference's own (dead) 1e-8 check, compositeLikIAK3D.R:382."""
    ds = synth.make_dataset(160, 5)
    P, modelx, types, cmeOpt = case_pars(kind, nonstat, nud)
    sm = orc.setupIAK3D(ds["xData"], ds["dIData"])
    Cs, s2 = orc.setCIAK3D(P, modelx, *types, cmeOpt, sm, quirk_cd1_unscaled=False)
    sm2 = orc.setupIAK3D2(ds["xData"], ds["dIData"], ds["xData"], ds["dIData"])
    Cx = orc.setCIAK3D2(P, modelx, *types, cmeOpt, sm2)
    assert np.max(np.abs(Cx + P["cme"] * np.eye(160) - Cs)) <= 1e-13
    assert np.all(s2 >= np.diag(Cs) - 1e-12)           # point variance >= increment-averaged variance


def test_cd1_quirk_is_reproduced():
    ds = synth.make_dataset(100, 5)
    P, modelx, types, cmeOpt = case_pars("matern0.5", True, 0.5)
    sm = orc.setupIAK3D(ds["xData"], ds["dIData"])
    Cq, _ = orc.setCIAK3D(P, modelx, *types, cmeOpt, sm, quirk_cd1_unscaled=True)
    Cn, _ = orc.setCIAK3D(P, modelx, *types, cmeOpt, sm, quirk_cd1_unscaled=False)
    Phi, _ = orc.iaCovMatern(sm["dIU"], P["ad"], P["nud"], P["sdfdPars_cd1"], -1)
    want = (1.0 - P["cd1"]) * Phi[sm["did"][:, None], sm["did"][None, :]]
    assert np.max(np.abs((Cq - Cn) - want)) < 1e-14


def test_likelihood_against_dense_formula():
    ds = synth.make_dataset(250, 3)
    P, modelx, types, cmeOpt = case_pars("matern0.5", False, 0.5)
    sm = orc.setupIAK3D(ds["xData"], ds["dIData"])
    A, _ = orc.setCIAK3D(P, modelx, *types, cmeOpt, sm)
    X, z = ds["XData"], ds["zData"]
    n, p = X.shape
    nll = orc.nllIAK3D(None, z, X, modelx, 0.5, *types, cmeOpt, True, sm, PARBNDS, useReml=True, parsBTfmd=P)
    iA = np.linalg.inv(A)
    XiAX = X.T @ iA @ X
    beta = np.linalg.solve(XiAX, X.T @ iA @ z)
    r = z - X @ beta
    s2 = float(r @ iA @ r) / (n - p)
    want = 0.5 * ((n - p) * math.log(2 * math.pi) + np.linalg.slogdet(s2 * A)[1] + np.linalg.slogdet(XiAX / s2)[1] + (n - p))
    assert abs(nll - want) <= 1e-9 * abs(want)


def test_cl_single_block_equals_exact():
    ds = synth.make_dataset(200, 9)
    P, modelx, types, cmeOpt = case_pars("edgeroi", False, 0.5)
    sm = orc.setupIAK3D(ds["xData"], ds["dIData"])
    exact = orc.nllIAK3D(None, ds["zData"], ds["XData"], modelx, 0.5, *types, cmeOpt, True, sm, PARBNDS, useReml=True, parsBTfmd=P)
    cl = dict(compLikOptn=4, listBlocks=[np.arange(200)], subsetPairsAdj=np.zeros((0, 2), int), subsetPairsNonadj=np.zeros((0, 2), int))
    got = orc.nllIAK3D_CL(None, ds["zData"], ds["XData"], modelx, 0.5, *types, cmeOpt, True, PARBNDS, True, cl, ds["xData"],
                          ds["dIData"], parsBTfmd=P)
    assert abs(got - exact) <= 1e-12 * abs(exact)


def test_kriging_reproduces_data_without_measurement_error():
    ds = synth.make_dataset(150, 21)
    P, modelx, types, _ = case_pars("matern0.5", False, 0.5)
    P = dict(P); P["cme"] = 0.0
    sm = orc.setupIAK3D(ds["xData"], ds["dIData"])
    fit = orc.nllIAK3D(None, ds["zData"], ds["XData"], modelx, 0.5, *types, 0, True, sm, PARBNDS, useReml=True, rtnAll=True, parsBTfmd=P)
    fit["xData"], fit["dIData"] = ds["xData"], ds["dIData"]
    sel = np.array([3, 40, 77])
    zs, vs = [], []
    for i in sel:
        z, v, _, _ = orc.predictIAK3D(ds["xData"][i:i + 1], ds["dIData"][i:i + 1], lambda k, idx: ds["XData"][i:i + 1], fit,
                                      modelx, types, 0)
        zs.append(z[0, 0]); vs.append(v[0, 0])
    assert np.allclose(zs, ds["zData"][sel], rtol=1e-8)
    assert np.max(np.abs(vs)) < 1e-7


def test_gradhess_matches_finite_differences():
    ds = synth.make_dataset(120, 4)
    P, modelx, types, cmeOpt = case_pars("matern0.5", False, 0.5)
    sm = orc.setupIAK3D(ds["xData"], ds["dIData"])
    dA = orc.component_matrices(P, modelx, types, sm)
    th = np.log([P[k] for k in ("cx0", "cx1", "cd1", "cxd0", "cxd1", "cme")])
    W = np.column_stack([ds["XData"], ds["zData"]])
    g = orc.gradHessIAK3D(dA, th, W)
    for i in range(6):
        e = np.zeros(6); e[i] = 1e-5
        fd = (orc.gradHessIAK3D(dA, th + e, W)["nll"] - orc.gradHessIAK3D(dA, th - e, W)["nll"]) / 2e-5
        assert abs(fd - g["grad"][i]) <= 1e-5 * max(1.0, abs(g["grad"][i]))


def test_bessel_reference_values():
    # scipy.special.kv (the oracle's Bessel) against mpmath-free closed forms for half-integer orders
    x = np.array([0.01, 0.5, 2.0, 9.0, 40.0])
    assert relerr(ssp.kv(0.5, x), np.sqrt(np.pi / (2 * x)) * np.exp(-x)) < 1e-14
    assert relerr(ssp.kv(1.5, x), np.sqrt(np.pi / (2 * x)) * np.exp(-x) * (1 + 1 / x)) < 1e-14


def test_analytic_gradient_formula_against_central_differences():
    """The weights of the analytic gradient (oracle.gradnllIAK3D_analytic; product: iak3d_nll_grad): REML and ML, against
    central differences of the oracle's own objective."""
    import numpy as np
    from iak3d_b200 import synth
    from oracle import iak3d_oracle as orc
    from helpers import PARBNDS
    ds = synth.make_dataset(260, 33)
    sm = orc.setupIAK3D(ds["xData"], ds["dIData"])
    pars = np.array([0.2, -1.3, 0.1, -2.2, -1.4, -1.9, 0.4, -3.0])       # matern: ax, nux, ad, cx0, cx1, cd1, cxd1, cme
    for useReml in (True, False):
        f = lambda v: orc.nllIAK3D(v, ds["zData"], ds["XData"], "matern", 0.5, 0, 0, 0, 1, True, sm, PARBNDS, useReml=useReml)
        g = orc.gradnllIAK3D_analytic(pars, ds["zData"], ds["XData"], "matern", 0.5, 0, 0, 0, 1, True, sm, PARBNDS, useReml=useReml)
        for k in range(pars.size):
            h = 1e-5
            vp = pars.copy(); vp[k] += h; vm = pars.copy(); vm[k] -= h
            cd = (f(vp) - f(vm)) / (2 * h)
            assert abs(g[k] - cd) <= 2e-5 * max(1.0, abs(cd)), (useReml, k, g[k], cd)


def test_oracle_matches_frozen_end_to_end_cases():
    """tests/golden/oracle_cases.json (scripts/make_golden.py): the oracle must not drift from the values the GPU path is
    also checked against."""
    import sys
    sys.path.insert(0, os.path.join(HERE, "..", "scripts"))
    import make_golden
    with open(os.path.join(HERE, "golden", "oracle_cases.json")) as f:
        gold = json.load(f)
    for case, want in zip(make_golden.CASES, gold["cases"]):
        got = make_golden.compute(case)
        assert got["name"] == want["name"]
        for k in ("lndetA", "nll", "cxdhat"):
            assert abs(got[k] - want[k]) <= 1e-11 * abs(want[k]), (case[0], k)
        for k in ("WiAW", "betahat", "zMap", "vMap", "sigma2Vec_head"):
            a, b = np.asarray(got[k], float), np.asarray(want[k], float)
            assert np.max(np.abs(a - b)) <= 1e-10 * np.max(np.abs(b)), (case[0], k)
        for (i, j, c), (_, _, cw) in zip(got["C_entries"], want["C_entries"]):
            assert abs(c - cw) <= 1e-13 * abs(cw)


def _two_block_case(optn):
    ds = synth.make_dataset(400, 61)
    P, modelx, types, cmeOpt = case_pars("edgeroi", False, 0.5)
    blk = synth.make_blocks(ds["xData"], ds["prof"], 2, 61)
    blk["compLikOptn"] = optn
    lmm = dict(compLikMats=blk, xData=ds["xData"], dIData=ds["dIData"], parsBTfmd=P)
    z = ds["zData"].reshape(-1) - ds["zData"].mean()
    xMap = synth.make_grid(6)[:30]
    dIMap = np.repeat(synth.GSM_INTERVALS[1:2], 30, axis=0)
    return ds, P, modelx, types, cmeOpt, blk, lmm, z, xMap, dIMap


def _exact_simple_kriging(ds, P, modelx, types, cmeOpt, rows, xm, dm, z):
    sm = orc.setupIAK3D(np.vstack([xm, ds["xData"][rows]]), np.vstack([dm, ds["dIData"][rows]]))
    Cj, _ = orc.setCIAK3D(P, modelx, *types, cmeOpt, sm)
    m = xm.shape[0]
    T = np.linalg.solve(Cj[m:, m:], Cj[m:, :m])
    return T.T @ z[rows], np.diag(Cj[:m, :m] - Cj[:m, m:] @ T)


def test_eidsvik_block_predictor_with_one_pair_is_exact_kriging():
    """Pins the restatement of predMatsIAK3D_CLEV_ByBlock (compositeLikIAK3D.R:609-798) independently of its own algebra:
    with two blocks there is one pair, the composite likelihood IS the full likelihood of the joined data, and the
    Eidsvik predictor (A0, b0, Bk0, J0 and the sandwich invA0 J0 invA0) must collapse to simple kriging from all data --
    prediction and conditional variance."""
    ds, P, modelx, types, cmeOpt, blk, lmm, z, xMap, dIMap = _two_block_case(2)
    zk, vk = orc.predMatsIAK3D_CLEV_ByBlock(z, ds["XData"], xMap, dIMap, lmm, modelx, types, cmeOpt)
    bm = np.argmin(orc.xyDist(xMap, np.asarray(blk["centroids"])), axis=1)
    for k in (0, 1):
        pts = np.nonzero(bm == k)[0]
        rows = np.concatenate([blk["listBlocks"][k], blk["listBlocks"][1 - k]])
        ze, ve = _exact_simple_kriging(ds, P, modelx, types, cmeOpt, rows, xMap[pts], dIMap[pts], z)
        assert np.max(np.abs(zk[pts] - ze)) <= 1e-11 * np.max(np.abs(ze))
        assert np.max(np.abs(vk[pts] - ve) / ve) <= 1e-11


def test_pairwise_block_predictor_with_one_pair_is_exact_kriging():
    """Same collapse for predMatsIAK3D_CL_ByBlock (compositeLikIAK3D.R:471-607): one neighbour, so the average over the
    neighbours is the kriging term of the single pair."""
    ds, P, modelx, types, cmeOpt, blk, lmm, z, xMap, dIMap = _two_block_case(1)
    t = orc.predMatsIAK3D_CL_ByBlock(z, ds["XData"], xMap, dIMap, lmm, modelx, types, cmeOpt)
    bm = np.argmin(orc.xyDist(xMap, np.asarray(blk["centroids"])), axis=1)
    for k in (0, 1):
        pts = np.nonzero(bm == k)[0]
        rows = np.concatenate([blk["listBlocks"][k], blk["listBlocks"][1 - k]])
        ze, ve = _exact_simple_kriging(ds, P, modelx, types, cmeOpt, rows, xMap[pts], dIMap[pts], z)
        assert np.max(np.abs(t["CkhiCz_muhat"][pts] - ze)) <= 1e-11 * np.max(np.abs(ze))
        assert np.max(np.abs((t["diagCkk"][pts] - t["diagCkhiCChk"][pts]) - ve) / ve) <= 1e-10


def test_ssre_structures_are_block_diagonal_outer_products():
    """listStructs4ssre (iaCovMatern.R:440-460): C += cssre_k * blockdiag_over_sites(X[site, k] X[site, k]')."""
    from iak3d_b200 import synth
    ds = synth.make_dataset(200, 7)
    site = np.array([int(p) // 2 for p in ds["prof"]])
    cols = [0, 6]
    sm = orc.setupIAK3D(ds["xData"], ds["dIData"], siteIDData=site, XData=ds["XData"], cols4ssre=cols)
    P, modelx, types, cmeOpt = synth.default_pars("matern0.5")
    P = dict(P); P["nud"] = 0.5; P["cssre"] = np.array([0.3, 0.02])
    C1, s2 = orc.setCIAK3D(P, modelx, *types, cmeOpt, sm)
    C0, _ = orc.setCIAK3D(P, modelx, *types, cmeOpt, orc.setupIAK3D(ds["xData"], ds["dIData"]))
    D = C1 - C0
    Z = np.zeros_like(D)
    for s_ in np.unique(site):
        m = site == s_
        for k, c in enumerate(cols):
            Z[np.ix_(m, m)] += P["cssre"][k] * np.outer(ds["XData"][m, c], ds["XData"][m, c])
    assert np.max(np.abs(D - Z)) <= 1e-13 * np.max(np.abs(Z)) and np.isnan(s2).all()
    # the rectangular structures of a set against itself are the square ones
    sm2 = orc.setupIAK3D2(ds["xData"], ds["dIData"], ds["xData"], ds["dIData"], siteIDData=site, XData=ds["XData"], siteIDData2=site,
                          XData2=ds["XData"], cols4ssre=cols)
    C2 = orc.setCIAK3D2(P, modelx, *types, cmeOpt, sm2)
    assert np.max(np.abs(C2 + P["cme"] * np.eye(200) - C1)) <= 1e-12 * np.max(np.abs(C1))       # compositeLikIAK3D.R:382's check
    # readParsIAK3D: whatever follows the model's own parameters are log site variances (fitIAK3D.R:1413)
    pb = orc.readParsIAK3D(np.array([0.1, 0.2, -1.0, -2.0, -1.5, -3.0, -4.0]), "spherical", 0.5, -9, 0, 0, 0, True, PARBNDS)
    assert np.allclose(pb["cssre"], np.exp([-3.0, -4.0]))
