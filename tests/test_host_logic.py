"""Host-side logic of the package (parameter transforms, REML/ML tail, composite-likelihood bookkeeping, grid
tiling, synthetic generator) against the oracle.  CPU only."""
import numpy as np
import pytest

import iak3d_b200 as iak
from iak3d_b200 import api, parallel, synth
from oracle import iak3d_oracle as orc
from helpers import PARBNDS


@pytest.mark.parametrize("modelx,types,cmeOpt,prodSum", [
    ("matern", (0, 0, 0), 1, True), ("spherical", (-9, 0, 0), 1, True), ("matern", (-1, -1, -1), 0, True),
    ("nugget", (0, 0, 0), 1, True), ("wendland", (0, -1, 0), 1, False), ("spherical", (-9, 0, -1), 1, True)])
def test_readPars_matches_oracle(modelx, types, cmeOpt, prodSum):
    rng = np.random.default_rng(3)
    v = rng.uniform(-2.0, 2.0, 20)
    a = api.readParsIAK3D(v, modelx, 1.5, *types, cmeOpt, prodSum, PARBNDS)
    b = orc.readParsIAK3D(v, modelx, 1.5, *types, cmeOpt, prodSum, PARBNDS)
    for k in b:
        x, y = np.asarray(a[k], float), np.asarray(b[k], float)
        assert x.shape == y.shape, k
        assert np.array_equal(np.isnan(x), np.isnan(y)), k
        assert np.allclose(x[~np.isnan(x)], y[~np.isnan(y)], rtol=1e-15, atol=0), k


def test_tau_transform_round_trip():
    t = api.tauTfm([0.6, 1.7], PARBNDS)
    back = api.tauTfm(t, PARBNDS, invt=True)
    assert np.allclose(back, [0.6, 1.7], rtol=1e-13)
    assert np.allclose(api.logitab(api.logitab(0.3, 0.1, 2.0), 0.1, 2.0, invt=True), 0.3)


@pytest.mark.parametrize("useReml", [True, False])
def test_reml_tail_matches_oracle(useReml):
    rng = np.random.default_rng(5)
    n, p = 400, 6
    X = rng.standard_normal((n, p)); z = rng.standard_normal(n) + X[:, 0]
    A = np.cov(rng.standard_normal((n, 3 * n))) + np.eye(n)
    W = np.column_stack([X, z])
    WiAW = W.T @ np.linalg.solve(A, W); lnd = np.linalg.slogdet(A)[1]
    a = api.reml_tail(WiAW, lnd, n, p, useReml)
    t = orc.reml_tail(WiAW, lnd, n, p, useReml)
    assert abs(a["cxdhat"] - t["cxdhat"]) <= 1e-13 * abs(t["cxdhat"])
    assert np.allclose(a["betahat"], t["betahat"], rtol=1e-11)
    c = t["cxdhat"]
    if useReml:
        want = 0.5 * ((n - p) * np.log(2 * np.pi) + lnd + n * np.log(c) + t["lndetXiAX"] - p * np.log(c) + t["resiAres"] / c)
    else:
        want = 0.5 * (n * np.log(2 * np.pi) + lnd + n * np.log(c) + t["resiAres"] / c)
    assert abs(a["nll"] - want) <= 1e-12 * abs(want)


def test_reml_tail_singular_design_is_badnll():
    WiAW = np.zeros((4, 4)); WiAW[3, 3] = 1.0
    assert api.reml_tail(WiAW, 0.0, 10, 3)["nll"] == iak.BADNLL


def test_grid_tile_partitions_exactly():
    for n in (1, 127, 1000, 10 ** 6):
        for world in (1, 2, 4, 8):
            cover = []
            for r in range(world):
                lo, hi = parallel.grid_tile(n, r, world)
                assert 0 <= lo <= hi <= n
                cover += list(range(lo, hi)) if n <= 1000 else []
                if r + 1 < world and hi < n:
                    assert (hi - lo) % 128 == 0
            if n <= 1000:
                assert cover == list(range(n))
            else:
                assert sum(parallel.grid_tile(n, r, world)[1] - parallel.grid_tile(n, r, world)[0] for r in range(world)) == n


def test_synth_is_deterministic_and_shaped():
    a = synth.make_dataset(1200, synth.SEEDS["E"]); b = synth.make_dataset(1200, synth.SEEDS["E"])
    assert np.array_equal(a["zData"], b["zData"]) and a["XData"].shape == (1200, 8)
    dI = a["dIData"]
    assert np.all(dI[:, 1] > dI[:, 0]) and np.allclose(np.round(dI, 2), dI)
    nprof = int(a["prof"].max()) + 1
    assert 250 <= nprof <= 500
    blk = synth.make_blocks(a["xData"], a["prof"], 6, 1)
    assert sorted(np.concatenate(blk["listBlocks"]).tolist()) == list(range(1200))
    allp = {tuple(r) for r in blk["subsetPairsAdj"]} | {tuple(r) for r in blk["subsetPairsNonadj"]}
    assert len(allp) == 15


def test_make_pars_struct():
    P, modelx, types, cmeOpt = synth.default_pars("matern0.5", True)
    s = iak._lib.make_pars(P, modelx, *types, cmeOpt)
    assert s.modelx == 2 and list(s.sdfdType) == [-1, -1, -1] and s.quirk_cd1_unscaled == 1
    assert (s.ax, s.ad, s.cme) == (10.0, 0.35, 0.05) and s.sdfdPars[2][1] == 1.7
    P2, modelx2, types2, _ = synth.default_pars("edgeroi")
    s2 = iak._lib.make_pars(P2, modelx2, *types2, 1)
    assert s2.modelx == 1 and s2.sdfdType[0] == -9 and np.isnan(s2.nux)


def test_reml_tail_batch_equals_scalar_tail():
    import numpy as np
    from iak3d_b200 import api
    rng = np.random.default_rng(4)
    p, n, cnt = 5, 300, 6
    W = []
    for i in range(cnt):
        B = rng.standard_normal((p + 1, p + 1)); W.append(B @ B.T + (p + 1) * np.eye(p + 1))
    W = np.array(W); lnd = rng.uniform(-50, 50, cnt)
    W[3, 0, 0] = -1.0                                # not positive definite -> BADNLL
    for reml in (True, False):
        got = api.reml_tail_batch(W, lnd, n, p, reml)
        for i in range(cnt):
            want = api.reml_tail(W[i], lnd[i], n, p, reml)["nll"]
            assert got[i] == want or abs(got[i] - want) <= 1e-12 * abs(want)
        assert got[3] == api.BADNLL
