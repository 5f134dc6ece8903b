"""CPU: the oracle restatement of makeXvX (oracle/makexvx_oracle.py) against constructions that do not share its code."""
import numpy as np
import pytest

from oracle import makexvx_oracle as mx
from oracle import iak3d_oracle as orc
from design_cases import ALLKNOTS, cubist_data, cubist_model, intervals, mlr_case


def test_bs_basis_properties():
    x = np.linspace(0.0, 7.13, 400)
    B = mx.bs_basis(x, ALLKNOTS[1:-1], (ALLKNOTS[0], ALLKNOTS[-1]))
    assert B.shape == (400, 6)
    # with its dropped first column a cubic B-spline basis sums to one; that column is (1 - x / k1)^3 on the first span
    first = np.where(x < 0.3, (1 - x / 0.3) ** 3, 0.0)
    assert np.max(np.abs(B.sum(axis=1) + first - 1.0)) < 1e-13
    assert np.all(B >= -1e-15) and B[-1, -1] == 1.0
    assert np.isnan(mx.bs_basis([7.2], ALLKNOTS[1:-1], (0.0, 7.13))).all()


def test_seq_by_matches_r_semantics():
    v = mx.seq_by(0.1, 0.4, 10)
    assert v.size == 10 and v[0] == 0.1 and v[-1] <= 0.4 and abs(v[-1] - 0.4) < 1e-15
    assert mx.seq_by(0.2, 0.5, 1)[0] == 0.35


def test_mlr_columns_against_direct_formulas():
    cov, levels, dI = mlr_case(60, 3)
    out = mx.makeXvX(cov, dI, 2.5, covLevels=levels)
    names = out["namesX"]
    assert names[:6] == ["const", "elev", "twi", "geol", "geol", "geol"] and names[6] == "d" and names[-1] == "d3"
    X = out["X"]; a, b = dI[:, 0], dI[:, 1]
    assert np.array_equal(X[:, 0], np.ones(60)) and np.array_equal(X[:, 1], cov["elev"])
    assert np.array_equal(X[:, 3], (cov["geol"] == 2).astype(float)) and np.array_equal(X[:, 5], (cov["geol"] == 4).astype(float))
    j = names.index("d.elev"); assert np.allclose(X[:, j], cov["elev"] * (a + b) / 2, rtol=1e-15)
    j = names.index("d2"); assert np.allclose(X[:, j], a * a + a * b + b * b, rtol=1e-15)          # the reference's moment, no 1/3
    j = names.index("d2.twi"); assert np.allclose(X[:, j], cov["twi"] * (a * a + a * b + b * b), rtol=1e-15)
    j = names.index("d3"); assert np.allclose(X[:, j], a ** 3 + a * a * b + a * b * b + b ** 3, rtol=1e-14)
    # limits were an output: min / max of the depth columns, infinite elsewhere
    jd = names.index("d"); assert out["XLims"][0, jd] == X[:, jd].min() and out["XLims"][1, 0] == -np.inf


def test_mlr_limits_are_recycled_from_the_first_column():
    cov, levels, dI = mlr_case(40, 5, factor=False)
    free = mx.makeXvX(cov, dI, 1)
    p = len(free["namesX"])                       # const elev twi d d.elev d.twi
    L = np.vstack([np.full(p, -np.inf), np.full(p, np.inf)])
    L[1, 0] = 0.5
    # one infinite lower and one infinite upper limit anywhere switch ALL clamping off (infBnds, makeXvX.R:190-194)
    assert np.array_equal(mx.makeXvX(cov, dI, 1, XLims=L)["X"], free["X"])
    L = np.vstack([np.full(p, -1e9), np.full(p, 1e9)])
    L[1, 0] = 0.5                                 # upper limit of 'const' ...
    got = mx.makeXvX(cov, dI, 1, XLims=L)
    jd = free["namesX"].index("d")
    # ... meets the depth columns through mapply's recycling over t(X[, id123]): entry (row i, k-th depth column) takes limit
    # column (3 i + k) mod 6, so 'const's limit clamps 'd' (k = 0) in the even rows only
    even = np.arange(40) % 2 == 0
    assert np.array_equal(got["X"][:, jd], np.where(even, np.minimum(free["X"][:, jd], 0.5), free["X"][:, jd]))
    jde = free["namesX"].index("d.elev")
    assert np.array_equal(got["X"][:, jde], free["X"][:, jde])


def test_mlr_spline_columns_are_point_averages():
    cov, levels, dI = mlr_case(25, 7, factor=False)
    names = ["const", "elev"] + [f"dSpline.{k + 1}" for k in range(6)]
    out = mx.makeXvX(cov, dI, names, allKnotsd=ALLKNOTS, nDiscPts=10)
    for i in (0, 7, 24):
        pts = np.linspace(dI[i, 0], dI[i, 1], 10)
        want = mx.bs_basis(pts, ALLKNOTS[1:-1], (0.0, 7.13)).mean(axis=0)
        assert np.allclose(out["X"][i, 2:], want, rtol=1e-12, atol=1e-15)


@pytest.mark.parametrize("opt", [0, 1])
def test_cubist_trapezoids_equal_dense_quadrature(opt):
    cm = cubist_model(True, opt)
    D, dI = cubist_data(30, 11)
    out = mx.makeXvX(D, dI, cm, allKnotsd=ALLKNOTS, opt_dSpline=opt, nDiscPts=10)
    X = out["X"]
    pc = len(cm["comms4XCubist"])
    cm2 = dict(cm); cm2["allKnotsd"] = ALLKNOTS; cm2["opt_dSpline"] = opt
    for i in range(30):
        # the rule columns are piecewise linear in depth with jumps at the breaks: integrate piece by piece, densely
        a, b = dI[i]
        if a in (0.3, 0.5, 0.9) or b in (0.3, 0.5, 0.9):
            continue      # an interval that starts / ends ON a break is averaged across it by the reference (only interior breaks split)
        edges = [a] + [v for v in mx.getAlldBreaks(cm2) if a < v < b] + [b]
        tot = np.zeros(pc)
        for lo, hi in zip(edges[:-1], edges[1:]):
            dd = np.linspace(lo + 1e-9, hi - 1e-9, 2001)
            Di = np.repeat(D[i:i + 1], dd.size, axis=0); Di[:, 3] = dd
            Xi, _ = mx.cubist2XGivenSetup(cm2, Di)
            f = Xi[:, :pc]
            tot += ((f[1:] + f[:-1]) * 0.5 * np.diff(dd)[:, None]).sum(axis=0)
        want = tot / (b - a)
        # the reference evaluates 1e-4 beside each break instead of at it: error <= slope * 1e-4
        assert np.allclose(X[i, :pc], want, rtol=0, atol=2e-3 * max(1.0, np.abs(want).max()))
        pts = a + (np.arange(10) / 9.0) * (b - a)
        assert np.allclose(X[i, pc:], mx.allKnotsd2X(pts, ALLKNOTS, opt).mean(axis=0), rtol=1e-12, atol=1e-14)


def test_cubist_committee_normalisation():
    cm = cubist_model(False)
    D, dI = cubist_data(200, 13)
    D[:, 3] = 0.2
    X, mat = mx.cubist2XGivenSetup(cm, D)
    # const columns of committee 1 (rules 1-3): 1 / (rules applying in the committee) / 2 committees
    cnt1 = mat[:, :4].sum(axis=1)
    j_const_r1 = 0
    m = mat[:, 0] == 1
    assert np.allclose(X[m, j_const_r1], 1.0 / cnt1[m] / 2.0)
    assert np.all(X[~m, j_const_r1] == 0.0)


def test_lognormal_branch_is_mean_and_sample_covariance_of_point_designs():
    cov, levels, dI = mlr_case(15, 9, factor=False)
    out = mx.makeXvX_lnN(cov, dI, 1, nDiscPts=10)
    names = out["namesX"]; iU = out["iU"]
    assert [names[j] for j in iU] == ["d", "d.elev", "d.twi"]
    for i in (0, 6, 14):
        dd = np.linspace(dI[i, 0], dI[i, 1], 10)
        pts = np.column_stack([np.ones(10), np.full(10, cov["elev"][i]), np.full(10, cov["twi"][i]), dd, dd * cov["elev"][i], dd * cov["twi"][i]])
        assert np.allclose(out["X"][i], pts.mean(axis=0), rtol=1e-13)
        assert np.allclose(out["vXU"][3 * i:3 * i + 3], np.cov(pts[:, iU].T), rtol=1e-10, atol=1e-18)
    # the interval mean of d agrees with the closed form of the other branch; d2 does not (that branch's moment lacks its 1/3)
    assert np.allclose(out["X"][:, 3], mx.makeXvX(cov, dI, 1)["X"][:, 3], rtol=1e-13)
