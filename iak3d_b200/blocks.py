"""Host mirror of the reference's composite-likelihood block construction (compositeLikIAK3D.R:803-1133): setVoronoiBlocks,
setVoronoiBlocksWrap, balanceNeighbours, balanceNeighbours2, calcnPerCluster, getBlocksFromLocations.  The distance /
nearest-centroid passes (every location against every centroid -- the whole cost of balanceNeighbours2's random search and of
placing 10^6 map cells) run on the GPU (csrc/voronoi.cu: iak3d_points_*); the greedy seeding and the Dirichlet-tile adjacency
(the reference's deldir call) are the library's host algorithms (iak3d_voronoi_seed, iak3d_dirichlet_adjacency).

compLikMats is the dict the rest of the package consumes: listBlocks (row indices per block, 0-based), xU_blocks, centroids,
subsetPairsAdj, subsetPairsNonadj."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import dptr, f64, iptr


def _unique_first(x):
    x = np.ascontiguousarray(np.asarray(x, float).reshape(len(x), -1))
    _, first = np.unique(x, axis=0, return_index=True)
    return x[np.sort(first)]


class DevicePoints:
    """Locations resident on the GPU for repeated nearest-centroid passes."""

    def __init__(self, xy, ctx=None):
        self.ctx = ctx or _lib.default_context()
        xy = f64(np.asarray(xy, float).reshape(-1, 2))
        self.m = xy.shape[0]
        h = C.c_void_p()
        self.ctx.check(self.ctx.lib.iak3d_points_create(self.ctx.h, self.m, dptr(xy), C.byref(h)))
        self.h = h
        self._close_order = 0
        self.ctx.adopt(self)

    def nearest(self, vcentres, want_blocks=True, want_counts=False):
        vc = f64(np.asarray(vcentres, float).reshape(-1, 2))
        nB = vc.shape[0]
        blk = np.empty(self.m, np.int32) if want_blocks else None
        cnt = np.zeros(nB, np.int64) if want_counts else None
        cp = cnt.ctypes.data_as(C.POINTER(C.c_int64)) if cnt is not None else None
        self.ctx.check(self.ctx.lib.iak3d_points_nearest(self.h, nB, dptr(vc), iptr(blk), cp))
        return blk, cnt

    def close(self):
        if getattr(self, "h", None) is not None and self.h.value:
            self.ctx.lib.iak3d_points_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def seed_centroids(xU, nPerBlock):
    lib = _lib.load()
    xU = f64(np.asarray(xU, float).reshape(-1, 2))
    nB = xU.shape[0] // int(nPerBlock)
    if nB < 1:
        raise ValueError("fewer unique locations than nPerBlock")
    vc = np.empty((nB, 2), order="F")
    got = C.c_int32()
    st = lib.iak3d_voronoi_seed(xU.shape[0], dptr(xU), int(nPerBlock), nB, C.byref(got), dptr(vc))
    if st != 0:
        raise _lib.Iak3dError(f"iak3d_voronoi_seed status {st}")
    return np.ascontiguousarray(vc)


def dirichlet_adjacency(vcentres):
    lib = _lib.load()
    vc = f64(np.asarray(vcentres, float).reshape(-1, 2))
    nB = vc.shape[0]
    cap = max(3 * nB, 1)
    pairs = np.empty((cap, 2), np.int32)
    npairs = C.c_int64()
    st = lib.iak3d_dirichlet_adjacency(nB, dptr(vc), iptr(pairs), cap, C.byref(npairs))
    if st != 0:
        raise _lib.Iak3dError(f"iak3d_dirichlet_adjacency status {st}")
    return pairs[:npairs.value].astype(np.int64)


def setVoronoiBlocks(x, nPerBlock=50, plotVor=False, vcentres=None, ctx=None, _pts=None):
    """compositeLikIAK3D.R:829-935."""
    x = np.asarray(x, float).reshape(-1, 2)
    xU = _unique_first(x)
    vc = seed_centroids(xU, nPerBlock) if vcentres is None else np.array(vcentres, float).reshape(-1, 2)
    nB = len(vc)
    pts = _pts or DevicePoints(xU, ctx=ctx)
    try:
        iBlock, _ = pts.nearest(vc)
    finally:
        if _pts is None:
            pts.close()
    blocks, xUb = [], []
    for i in range(nB):
        xb = xU[iBlock == i]
        xUb.append(xb)
        # rows whose x occurs among the block's x values and whose y among its y values -- separately (:899)
        blocks.append(np.where(np.isin(x[:, 0], xb[:, 0]) & np.isin(x[:, 1], xb[:, 1]))[0])
    adj = dirichlet_adjacency(vc)
    aset = {tuple(r) for r in adj}
    non = np.array([(i, j) for i in range(nB) for j in range(i + 1, nB) if (i, j) not in aset], np.int64).reshape(-1, 2)
    return dict(listBlocks=blocks, xU_blocks=xUb, centroids=vc, subsetPairsAdj=adj, subsetPairsNonadj=non)


def _colMeans(a):
    a = np.asarray(a, float)
    return np.asarray(a.astype(np.longdouble).sum(axis=0) / np.longdouble(a.shape[0]), float)


def _dist(a, b):
    a = np.asarray(a, float); b = np.asarray(b, float)
    return np.sqrt((a[:, None, 0] - b[None, :, 0]) ** 2 + (a[:, None, 1] - b[None, :, 1]) ** 2)


def balanceNeighbours(x, nPerBlock=50, plotVor=False, compLikMats=None, ctx=None):
    """compositeLikIAK3D.R:958-1028 (optnBalance = 1): the most unequal adjacent pair is re-split across its longest
    axis, at most eleven times."""
    clm = compLikMats
    x = np.asarray(x, float).reshape(-1, 2)
    pts = DevicePoints(_unique_first(x), ctx=ctx)
    try:
        diffs = lambda c: np.array([len(c["xU_blocks"][i]) - len(c["xU_blocks"][j]) for i, j in c["subsetPairsAdj"]])
        adjDiffs = diffs(clm)
        vc = np.array(clm["centroids"], float)
        nUpdates, carryOn = 1, True
        while carryOn and np.max(np.abs(adjDiffs)) > nPerBlock:
            t = int(np.argmax(np.abs(adjDiffs)))
            iT, jT = (int(v) for v in clm["subsetPairsAdj"][t])
            X = np.vstack([clm["xU_blocks"][iT], clm["xU_blocks"][jT]])
            D = _dist(X, X)
            iDistant = int(np.argmax(D.T.reshape(-1)) % len(X))
            o = np.argsort(D[iDistant], kind="stable")
            half = o[:int(np.ceil(len(o) / 2))]
            rest = np.setdiff1d(np.arange(len(X)), half)
            D2i = _dist(X[iDistant:iDistant + 1], vc[iT:iT + 1])[0, 0]
            D2j = _dist(X[iDistant:iDistant + 1], vc[jT:jT + 1])[0, 0]
            if D2i < D2j:
                vc[iT] = _colMeans(X[half])
                vc[jT] = vc[iT]               # compositeLikIAK3D.R:996 as written: block j inherits block i's new centroid
            else:
                vc[jT] = _colMeans(X[half])
                vc[iT] = _colMeans(X[rest])
            clm = setVoronoiBlocks(x, nPerBlock, vcentres=vc, ctx=ctx, _pts=pts)
            if nUpdates > 10:
                carryOn = False
            nUpdates += 1
            adjDiffs = diffs(clm)
    finally:
        pts.close()
    return clm


def calcnPerCluster(vcentres, xU=None, ctx=None, _pts=None):
    """compositeLikIAK3D.R:1033-1046: locations per nearest centroid."""
    pts = _pts or DevicePoints(xU, ctx=ctx)
    try:
        _, cnt = pts.nearest(vcentres, want_blocks=False, want_counts=True)
    finally:
        if _pts is None:
            pts.close()
    return cnt


def balanceNeighbours2(x, nPerBlock=50, plotVor=False, compLikMats=None, draws=None, nReps=5, nItsOfNoChange4Opt=200, nEpochs=15,
                       ctx=None):
    """compositeLikIAK3D.R:1048-1133 (optnBalance = 2): random search that raises the size of the smallest block; every
    proposal is one nearest-centroid pass over all unique locations on the GPU.  `draws(k)` supplies k N(0,1) values (the
    reference's rnorm(k)); default: numpy's generator."""
    if draws is None:
        rng = np.random.default_rng()
        draws = rng.standard_normal
    x = np.asarray(x, float).reshape(-1, 2)
    xU = _unique_first(x)
    nc = len(compLikMats["centroids"])
    pts = DevicePoints(xU, ctx=ctx)
    try:
        ofVec, vcList = [], []
        for _ in range(nReps):
            vc = np.array(compLikMats["centroids"], float)
            sd = max(xU[:, 0].max() - xU[:, 0].min(), xU[:, 1].max() - x[:, 1].min()) / 10     # (sic) :1067
            cur = calcnPerCluster(vc, _pts=pts)
            of = -cur.min()
            for _e in range(nEpochs):
                cnt = 0
                while cnt <= nItsOfNoChange4Opt:
                    lo = np.argsort(cur, kind="stable")[:min(3, nc)]
                    hi = np.argsort(-cur, kind="stable")[:min(3, nc)]
                    iMove = list(dict.fromkeys(list(lo) + list(hi)))
                    prop = vc.copy()
                    prop[iMove] = prop[iMove] + sd * np.asarray(draws(2 * len(iMove)), float).reshape(2, len(iMove)).T
                    pc = calcnPerCluster(prop, _pts=pts)
                    ofp = -pc.min()
                    if ofp < of:
                        vc, cur, of, cnt = prop, pc, ofp, 0
                    else:
                        cnt += 1
                sd = sd / 2
            vcList.append(vc); ofVec.append(of)
        best = int(np.argmin(ofVec))
        return setVoronoiBlocks(x, nPerBlock, vcentres=vcList[best], ctx=ctx, _pts=pts)
    finally:
        pts.close()


def setVoronoiBlocksWrap(x, nPerBlock=50, plotVor=False, optnBalance=1, draws=None, ctx=None, **kw):
    """compositeLikIAK3D.R:938-952."""
    clm = setVoronoiBlocks(x, nPerBlock, ctx=ctx)
    if optnBalance == 1:
        return balanceNeighbours(x, nPerBlock, compLikMats=clm, ctx=ctx)
    if optnBalance == 2:
        return balanceNeighbours2(x, nPerBlock, compLikMats=clm, draws=draws, ctx=ctx, **kw)
    raise ValueError("Error - enter valid option for balancing neighbours!")


def getBlocksFromLocations(xMap, compLikMats, ctx=None):
    """compositeLikIAK3D.R:803-812: block of every map location = nearest centroid, first on ties."""
    pts = DevicePoints(xMap, ctx=ctx)
    try:
        blk, _ = pts.nearest(compLikMats["centroids"])
    finally:
        pts.close()
    return blk.astype(np.int64)
