"""CPU restatement of the reference's composite-likelihood block construction -- TEST INFRASTRUCTURE ONLY (imported by tests/
and bench.py's CPU legs, never by iak3d_b200/).

  setVoronoiBlocks      compositeLikIAK3D.R:829-935   greedy centroid seeding, nearest-centroid (Voronoi) assignment of the
                                                      unique locations, rows of each block, adjacency of the blocks
  setVoronoiBlocksWrap  :938-952
  balanceNeighbours     :958-1028   (optnBalance = 1; deterministic)
  calcnPerCluster       :1033-1046
  balanceNeighbours2    :1048-1133  (optnBalance = 2; random search -- the N(0,1) draws are an INPUT here, `rnorm(k)` is
                                     `draws(k)`, consumed in the reference's order, so that two implementations can be
                                     compared on the same stream; R's own Mersenne-Twister stream is not reproduced)
  getBlocksFromLocations :803-812

Third-party arithmetic: `deldir::deldir(x, y)$dirsgs` (package deldir, version unpinned): the edges of the Dirichlet (Voronoi)
tiles clipped to deldir's default window = bounding box of the points extended by 10 % per side; ind1, ind2 = the two points
an edge separates.  Restated through scipy.spatial.Delaunay (Qhull): the Dirichlet edge of a Delaunay edge is the segment
between the circumcentres of its two triangles (a ray for hull edges); the pair is adjacent when that edge meets the window.
Parity status: unpinned against R; pinned by construction tests (tests/test_blocks.py).

Reference quirks kept: rows join a block when their x coordinate occurs among the block's x coordinates AND their y coordinate
among its y coordinates, separately (:899); in balanceNeighbours the branch D2i < D2j stores block i's NEW centroid as the
centroid of block j as well (:996), so j loses its locations to i at the next assignment.  Indices are 0-based here."""
from __future__ import annotations

import numpy as np
from scipy.spatial import Delaunay

from .iak3d_oracle import _unique_first, xyDist


def colMeans(a):
    """R's colMeans accumulates in long double."""
    a = np.asarray(a, float)
    return np.asarray(a.astype(np.longdouble).sum(axis=0) / np.longdouble(a.shape[0]), float)


def _clip_hits(p, d, tmax, rw):
    t0, t1 = 0.0, tmax
    for pp, qq in ((-d[0], p[0] - rw[0]), (d[0], rw[1] - p[0]), (-d[1], p[1] - rw[2]), (d[1], rw[3] - p[1])):
        if pp == 0.0:
            if qq < 0.0:
                return False
            continue
        r = qq / pp
        if pp < 0.0:
            if r > t1: return False
            t0 = max(t0, r)
        else:
            if r < t0: return False
            t1 = min(t1, r)
    return t0 <= t1


def dirichlet_adjacency(vc):
    """Pairs (i < j) of deldir(vc)$dirsgs, sorted."""
    vc = np.asarray(vc, float)
    ok = ~np.isnan(vc).any(axis=1)
    _, first = np.unique(vc[ok], axis=0, return_index=True)
    orig = np.where(ok)[0][np.sort(first)]
    P = vc[orig]
    n = len(P)
    if n < 2:
        return np.zeros((0, 2), np.int64)
    if n == 2:
        return np.array([[orig.min(), orig.max()]], np.int64)
    dx, dy = np.ptp(P[:, 0]), np.ptp(P[:, 1])
    rw = (P[:, 0].min() - 0.1 * dx, P[:, 0].max() + 0.1 * dx, P[:, 1].min() - 0.1 * dy, P[:, 1].max() + 0.1 * dy)
    tri = Delaunay(P)
    cc = []
    for s in tri.simplices:
        a, b, c = P[s]
        bx, by, cx, cy = b[0] - a[0], b[1] - a[1], c[0] - a[0], c[1] - a[1]
        dd = 2.0 * (bx * cy - by * cx)
        cc.append((a[0] + (cy * (bx * bx + by * by) - by * (cx * cx + cy * cy)) / dd,
                   a[1] + (bx * (cx * cx + cy * cy) - cx * (bx * bx + by * by)) / dd))
    edges = {}
    for k, s in enumerate(tri.simplices):
        for e in range(3):
            i, j = sorted((int(s[e]), int(s[(e + 1) % 3])))
            edges.setdefault((i, j), []).append(k)
    out = set()
    for (i, j), ts in edges.items():
        if len(ts) >= 2:
            c1, c2 = cc[ts[0]], cc[ts[1]]
            hit = _clip_hits(c1, (c2[0] - c1[0], c2[1] - c1[1]), 1.0, rw)
        else:
            s = tri.simplices[ts[0]]
            k3 = [int(v) for v in s if v not in (i, j)][0]
            nx, ny = P[j, 1] - P[i, 1], -(P[j, 0] - P[i, 0])
            m = 0.5 * (P[i] + P[j])
            if nx * (P[k3, 0] - m[0]) + ny * (P[k3, 1] - m[1]) > 0:
                nx, ny = -nx, -ny
            hit = _clip_hits(cc[ts[0]], (nx, ny), np.inf, rw)
        if hit:
            a, b = int(orig[i]), int(orig[j])
            out.add((min(a, b), max(a, b)))
    return np.array(sorted(out), np.int64).reshape(-1, 2)


def seed_centroids(xU, nPerBlock):
    """:840-884."""
    nU = len(xU)
    nBlocks = nU // nPerBlock
    nSmall = nBlocks * (nPerBlock + 1) - nU
    nBig = nBlocks - nSmall
    rem = np.arange(nU)
    vc = np.full((nBlocks, 2), np.nan)
    for i in range(nBlocks):
        X = xU[rem]
        if len(X) == 0:
            continue
        DNS, DEW = X[:, 1].max() - X[:, 1].min(), X[:, 0].max() - X[:, 0].min()
        it = int(np.argmax(X[:, 1])) if DNS > DEW else int(np.argmax(X[:, 0]))
        D = xyDist(X, X[it:it + 1]).reshape(-1)
        o = np.argsort(D, kind="stable")
        take = min(nPerBlock + 1, len(X)) if i < nBig else (min(nPerBlock, len(X)) if i < nBlocks - 1 else len(X))
        iThis = o[:take]
        vc[i] = colMeans(X[iThis])
        rem = np.delete(rem, iThis)
    return vc


def nearest_centroid(x, vc):
    D = xyDist(np.asarray(x, float), np.asarray(vc, float))
    D = np.where(np.isnan(D), np.inf, D)
    return D.argmin(axis=1)


def setVoronoiBlocks(x, nPerBlock=50, vcentres=None):
    x = np.asarray(x, float).reshape(-1, 2)
    xU, _ = _unique_first(x)
    vc = seed_centroids(xU, nPerBlock) if vcentres is None else np.array(vcentres, float)
    nB = len(vc)
    iBlock = nearest_centroid(xU, vc)
    blocks, xUb = [], []
    for i in range(nB):
        xb = xU[iBlock == i]
        xUb.append(xb)
        blocks.append(np.where(np.isin(x[:, 0], xb[:, 0]) & np.isin(x[:, 1], xb[:, 1]))[0])
    adj = dirichlet_adjacency(vc)
    aset = {tuple(r) for r in adj}
    non = np.array([(i, j) for i in range(nB) for j in range(i + 1, nB) if (i, j) not in aset], np.int64).reshape(-1, 2)
    return dict(listBlocks=blocks, xU_blocks=xUb, centroids=vc, subsetPairsAdj=adj, subsetPairsNonadj=non)


def balanceNeighbours(x, nPerBlock, clm):
    maxPermissDiff, maxnUpdates = nPerBlock, 10
    diffs = lambda c: np.array([len(c["xU_blocks"][i]) - len(c["xU_blocks"][j]) for i, j in c["subsetPairsAdj"]])
    adjDiffs = diffs(clm)
    vc = np.array(clm["centroids"], float)
    nUpdates, carryOn = 1, True
    while carryOn and np.max(np.abs(adjDiffs)) > maxPermissDiff:
        t = int(np.argmax(np.abs(adjDiffs)))
        iT, jT = (int(v) for v in clm["subsetPairsAdj"][t])
        X = np.vstack([clm["xU_blocks"][iT], clm["xU_blocks"][jT]])
        D = xyDist(X, X)
        iDistant = int(np.argmax(D.T.reshape(-1)) % len(X))      # which(D == max(D), arr.ind = TRUE)[1, 1]: column-major first hit
        o = np.argsort(D[iDistant], kind="stable")
        half = o[:int(np.ceil(len(o) / 2))]
        rest = np.setdiff1d(np.arange(len(X)), half)
        D2i = xyDist(X[iDistant:iDistant + 1], vc[iT:iT + 1])[0, 0]
        D2j = xyDist(X[iDistant:iDistant + 1], vc[jT:jT + 1])[0, 0]
        if D2i < D2j:
            vc[iT] = colMeans(X[half])
            vc[jT] = vc[iT]                                         # :996 as written
        else:
            vc[jT] = colMeans(X[half])
            vc[iT] = colMeans(X[rest])
        clm = setVoronoiBlocks(x, nPerBlock, vcentres=vc)
        if nUpdates > maxnUpdates:
            carryOn = False
        nUpdates += 1
        adjDiffs = diffs(clm)
    return clm


def calcnPerCluster(vc, xU):
    return np.bincount(nearest_centroid(xU, vc), minlength=len(vc))


def balanceNeighbours2(x, nPerBlock, clm, draws, nReps=5, nItsOfNoChange4Opt=200, nEpochs=15):
    x = np.asarray(x, float).reshape(-1, 2)
    xU, _ = _unique_first(x)
    nc = len(clm["centroids"])
    ofVec, vcList = [], []
    for _ in range(nReps):
        vc = np.array(clm["centroids"], float)
        sd = max(xU[:, 0].max() - xU[:, 0].min(), xU[:, 1].max() - x[:, 1].min()) / 10      # (sic) min over x, not xU (:1067)
        cur = calcnPerCluster(vc, xU)
        of = -cur.min()
        for _e in range(nEpochs):
            cnt = 0
            while cnt <= nItsOfNoChange4Opt:
                lo = np.argsort(cur, kind="stable")[:min(3, nc)]
                hi = np.argsort(-cur, kind="stable")[:min(3, nc)]
                iMove = list(dict.fromkeys(list(lo) + list(hi)))
                prop = vc.copy()
                prop[iMove] = prop[iMove] + sd * np.asarray(draws(2 * len(iMove)), float).reshape(2, len(iMove)).T
                pc = calcnPerCluster(prop, xU)
                ofp = -pc.min()
                if ofp < of:
                    vc, cur, of, cnt = prop, pc, ofp, 0
                else:
                    cnt += 1
            sd = sd / 2
        vcList.append(vc); ofVec.append(of)
    best = int(np.argmin(ofVec))
    return setVoronoiBlocks(x, nPerBlock, vcentres=vcList[best])


def setVoronoiBlocksWrap(x, nPerBlock=50, optnBalance=1, draws=None, **kw):
    clm = setVoronoiBlocks(x, nPerBlock)
    if optnBalance == 1:
        return balanceNeighbours(x, nPerBlock, clm)
    if optnBalance == 2:
        return balanceNeighbours2(x, nPerBlock, clm, draws, **kw)
    raise ValueError("enter valid option for balancing neighbours")


def getBlocksFromLocations(xMap, clm):
    return nearest_centroid(xMap, clm["centroids"])
