"""CPU: block construction of the composite likelihood.  The library's two host algorithms (greedy seeding, Bowyer-Watson
Dirichlet adjacency: no GPU involved) against the oracle restatement (NumPy + Qhull), and the oracle against construction
properties."""
import ctypes as C

import numpy as np
import pytest

import iak3d_b200 as iak
from iak3d_b200 import blocks
from oracle import voronoi_oracle as vo


def pts(n, seed, dup=0):
    rng = np.random.default_rng(seed)
    x = np.round(rng.uniform(0, 100, (n, 2)), 3)
    if dup:
        x = np.vstack([x, x[rng.integers(0, n, dup)]])
    return x


@pytest.mark.parametrize("n,nPer,seed", [(1000, 50, 1), (997, 31, 2), (5000, 490, 3), (120, 50, 4), (64, 64, 5)])
def test_seed_matches_oracle_bit_for_bit(n, nPer, seed):
    x = pts(n, seed)
    got = blocks.seed_centroids(x, nPer)
    want = vo.seed_centroids(x, nPer)
    assert got.shape == want.shape == (n // nPer, 2)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("n,seed", [(3, 1), (4, 2), (10, 3), (100, 4), (101, 5), (400, 6), (1000, 7)])
def test_dirichlet_adjacency_matches_qhull_restatement(n, seed):
    vc = pts(n, 100 + seed)
    got = blocks.dirichlet_adjacency(vc)
    want = vo.dirichlet_adjacency(vc)
    assert np.array_equal(got, want)


def test_adjacency_is_delaunay_minus_edges_outside_the_window():
    from scipy.spatial import Delaunay
    vc = pts(200, 42)
    tri = Delaunay(vc)
    dela = {tuple(sorted((int(s[a]), int(s[(a + 1) % 3])))) for s in tri.simplices for a in range(3)}
    got = {tuple(r) for r in blocks.dirichlet_adjacency(vc)}
    assert got <= dela and len(got) >= 0.9 * len(dela)
    # the dropped pairs are those whose Dirichlet edge lies wholly outside the window: both ends of it are outside
    lo, hi = vc.min(axis=0) - 0.1 * np.ptp(vc, axis=0), vc.max(axis=0) + 0.1 * np.ptp(vc, axis=0)
    def circ(s):
        a, b, c = vc[s]
        d = 2 * ((b[0] - a[0]) * (c[1] - a[1]) - (b[1] - a[1]) * (c[0] - a[0]))
        ux = ((c[1] - a[1]) * ((b - a) ** 2).sum() - (b[1] - a[1]) * ((c - a) ** 2).sum()) / d
        uy = ((b[0] - a[0]) * ((c - a) ** 2).sum() - (c[0] - a[0]) * ((b - a) ** 2).sum()) / d
        return a + np.array([ux, uy])
    cc = np.array([circ(s) for s in tri.simplices])
    inside = np.all((cc >= lo) & (cc <= hi), axis=1)
    for k, s in enumerate(tri.simplices):
        if inside[k]:
            for a in range(3):
                assert tuple(sorted((int(s[a]), int(s[(a + 1) % 3])))) in got


def test_adjacency_special_cases():
    assert np.array_equal(blocks.dirichlet_adjacency([[0.0, 0.0], [1.0, 2.0]]), [[0, 1]])
    # duplicated centroid (what compositeLikIAK3D.R:996 produces) and a block without a centroid
    vc = np.array([[0.0, 0.0], [4.0, 0.1], [2.0, 3.0], [4.0, 0.1], [np.nan, np.nan], [2.1, 1.0]])
    got = blocks.dirichlet_adjacency(vc)
    assert np.array_equal(got, vo.dirichlet_adjacency(vc))
    assert 3 not in got and 4 not in got
    col = np.array([[0.0, 0.0], [1.0, 1.0], [3.0, 3.0], [2.0, 2.0]])
    assert np.array_equal(blocks.dirichlet_adjacency(col), [[0, 1], [1, 3], [2, 3]])


def test_oracle_blocks_partition_the_locations():
    x = pts(3000, 9, dup=600)
    clm = vo.setVoronoiBlocks(x, 100)
    nB = len(clm["centroids"])
    assert nB == len(np.unique(x, axis=0)) // 100
    allU = np.vstack(clm["xU_blocks"])
    assert len(allU) == len(np.unique(x, axis=0)) == len(np.unique(allU, axis=0))
    # every row is in the block of its location (and, by the separate x / y matching of :899, possibly in others)
    lab = vo.nearest_centroid(x, clm["centroids"])
    for b in range(nB):
        assert set(np.where(lab == b)[0]) <= set(clm["listBlocks"][b])
    pairs = {tuple(r) for r in clm["subsetPairsAdj"]} | {tuple(r) for r in clm["subsetPairsNonadj"]}
    assert len(pairs) == nB * (nB - 1) // 2


def test_oracle_balancing_reduces_the_spread():
    rng = np.random.default_rng(5)
    x = np.vstack([rng.normal(30, 6, (1500, 2)), rng.uniform(0, 100, (500, 2))])      # clustered: unequal Voronoi blocks
    clm0 = vo.setVoronoiBlocks(x, 100)
    n0 = np.array([len(b) for b in clm0["xU_blocks"]])
    clm2 = vo.balanceNeighbours2(x, 100, clm0, np.random.default_rng(1).standard_normal, nReps=1, nItsOfNoChange4Opt=20, nEpochs=4)
    n2 = np.array([len(b) for b in clm2["xU_blocks"]])
    assert n2.min() >= n0.min() and n2.sum() == n0.sum()
