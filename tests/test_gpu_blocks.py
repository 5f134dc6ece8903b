"""GPU parity: composite-likelihood block construction (csrc/voronoi.cu + iak3d_b200/blocks.py) against the oracle
restatement.  Integer / index work: exact equality."""
import numpy as np
import pytest

import iak3d_b200 as iak
from iak3d_b200 import blocks
from oracle import voronoi_oracle as vo

pytestmark = pytest.mark.gpu


def pts(n, seed, dup=0, clustered=False):
    rng = np.random.default_rng(seed)
    x = np.round(rng.uniform(0, 100, (n, 2)), 3)
    if clustered:
        x = np.vstack([np.round(rng.normal(30, 6, (3 * n // 4, 2)), 3), x[: n // 4]])
    if dup:
        x = np.vstack([x, x[rng.integers(0, len(x), dup)]])
    return x


def same(a, b):
    assert np.array_equal(a["centroids"], b["centroids"], equal_nan=True)
    assert np.array_equal(a["subsetPairsAdj"], b["subsetPairsAdj"]) and np.array_equal(a["subsetPairsNonadj"], b["subsetPairsNonadj"])
    assert len(a["listBlocks"]) == len(b["listBlocks"])
    for r1, r2, u1, u2 in zip(a["listBlocks"], b["listBlocks"], a["xU_blocks"], b["xU_blocks"]):
        assert np.array_equal(r1, r2) and np.array_equal(u1, u2)


@pytest.mark.parametrize("n,nPer,seed", [(3000, 100, 1), (4100, 490, 2), (777, 50, 3)])
def test_setVoronoiBlocks(ctx, n, nPer, seed):
    x = pts(n, seed, dup=n // 3)
    same(iak.setVoronoiBlocks(x, nPer, ctx=ctx), vo.setVoronoiBlocks(x, nPer))


def test_nearest_centroid_ties_and_empty_centroids(ctx):
    x = np.array([[0.0, 0.0], [1.0, 0.0], [2.0, 0.0], [1.0, 5.0]])
    vc = np.array([[np.nan, np.nan], [2.0, 0.0], [0.0, 0.0], [2.0, 0.0]])
    p = blocks.DevicePoints(x, ctx=ctx)
    blk, cnt = p.nearest(vc, want_counts=True)
    assert blk.tolist() == [2, 1, 1, 1]                    # (1, 0) is equidistant: first minimum; the duplicate never wins
    assert cnt.tolist() == [0, 3, 1, 0]
    p.close()


def test_balanceNeighbours(ctx):
    x = pts(4000, 7, dup=800, clustered=True)
    want0 = vo.setVoronoiBlocks(x, 100)
    got0 = iak.setVoronoiBlocks(x, 100, ctx=ctx)
    same(got0, want0)
    want = vo.balanceNeighbours(x, 100, want0)
    got = iak.balanceNeighbours(x, 100, compLikMats=got0, ctx=ctx)
    same(got, want)
    same(iak.setVoronoiBlocksWrap(x, 100, optnBalance=1, ctx=ctx), want)


def test_balanceNeighbours2_on_the_same_draws(ctx):
    x = pts(3000, 9, clustered=True)
    want0 = vo.setVoronoiBlocks(x, 150)
    kw = dict(nReps=2, nItsOfNoChange4Opt=25, nEpochs=5)
    want = vo.balanceNeighbours2(x, 150, want0, np.random.default_rng(11).standard_normal, **kw)
    got = iak.balanceNeighbours2(x, 150, compLikMats=iak.setVoronoiBlocks(x, 150, ctx=ctx), draws=np.random.default_rng(11).standard_normal,
                                 ctx=ctx, **kw)
    same(got, want)
    assert min(len(u) for u in got["xU_blocks"]) >= min(len(u) for u in want0["xU_blocks"])


def test_getBlocksFromLocations_large(ctx):
    x = pts(5000, 13)
    clm = vo.setVoronoiBlocks(x, 100)
    rng = np.random.default_rng(2)
    xMap = rng.uniform(-5, 105, (200_000, 2))
    got = iak.getBlocksFromLocations(xMap, clm, ctx=ctx)
    want = np.concatenate([vo.getBlocksFromLocations(xMap[k:k + 20000], clm) for k in range(0, len(xMap), 20000)])
    assert np.array_equal(got, want)
