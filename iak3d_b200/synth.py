"""Deterministic synthetic soil-profile datasets of the shapes BASELINE.json names.

The Edgeroi data (R packages GSIF / ithir, getEdgeroiData.R:3,57) are not available offline, so every
configuration is generated: profiles uniform on a 100 km x 100 km square (coordinates in km,
getEdgeroiData.R:81), contiguous horizon intervals from the surface with thicknesses drawn from a small
set and depths rounded to the cm (fitIAK3D.R:127), an n x 8 design matrix and a response.  Host-side
NumPy only; no GPU, no oracle.
"""
from __future__ import annotations

import numpy as np

SEEDS = {"E": 20260001, "S5": 20260005, "S40": 20260040, "CL200": 20260200, "P1M": 20261000}
GSM_INTERVALS = np.array([[0.0, 0.05], [0.05, 0.15], [0.15, 0.30], [0.30, 0.60], [0.60, 1.00], [1.00, 2.00]])  # egEdgeroi.R:396
_THICK = np.array([0.05, 0.10, 0.15, 0.20, 0.30, 0.40, 0.50])


def make_profiles(n, seed, extent_km=100.0, min_sep_km=0.001):
    """Returns xData (n x 2, km), dIData (n x 2, m, rounded to 0.01), prof (n,) profile index."""
    rng = np.random.default_rng(seed)
    xs, ds, pid = [], [], []
    k = 0
    seen = set()
    while len(ds) < n:
        while True:
            xy = np.round(rng.uniform(0.0, extent_km, 2), 3)      # 1 m grid => >= 1 m separation
            key = (float(xy[0]), float(xy[1]))
            if key not in seen:
                seen.add(key)
                break
        top = 0.0
        nh = 0
        while top < 2.0 - 1e-9 and len(ds) < n and nh < 6:
            t = float(rng.choice(_THICK))
            bot = round(top + t, 2)
            xs.append(xy); ds.append((round(top, 2), bot)); pid.append(k)
            top = bot; nh += 1
            if rng.uniform() < 0.12:
                break
        k += 1
    return np.array(xs, float), np.array(ds, float), np.array(pid, np.int64)


def _depth_cols(dI):
    """Three interval-averaged smooth functions of depth (closed-form averages of d, d^2, exp(-2d))."""
    a, b = dI[:, 0], dI[:, 1]
    m1 = 0.5 * (a + b)
    m2 = (b ** 3 - a ** 3) / (3.0 * (b - a))
    m3 = -(np.exp(-2.0 * b) - np.exp(-2.0 * a)) / (2.0 * (b - a))
    return np.column_stack([m1 - 0.6, m2 - 0.6, m3 - 0.4])


def make_design(xData, dIData, prof, seed, p=8):
    """X = intercept + (p-4) per-profile N(0,1) covariates + 3 interval-averaged depth columns."""
    rng = np.random.default_rng(seed + 7)
    nprof = int(prof.max()) + 1
    cov = rng.standard_normal((nprof, p - 4))
    return np.column_stack([np.ones(len(prof)), cov[prof], _depth_cols(dIData)])


def make_response(xData, dIData, prof, X, seed):
    """z = X beta + smooth spatial field + per-profile effect + depth-correlated and white noise (O(n))."""
    rng = np.random.default_rng(seed + 13)
    p = X.shape[1]
    beta = np.concatenate([[30.0], rng.uniform(-2.0, 2.0, p - 1)])
    nprof = int(prof.max()) + 1
    k = rng.uniform(0.02, 0.3, (6, 2)); ph = rng.uniform(0, 2 * np.pi, 6); am = rng.uniform(0.5, 1.5, 6)
    field = sum(am[i] * np.cos(xData @ k[i] + ph[i]) for i in range(6))
    pe = 1.2 * rng.standard_normal(nprof)[prof]
    mid = dIData.mean(axis=1)
    slope = 1.5 * rng.standard_normal(nprof)[prof]
    return X @ beta + field + pe + slope * np.exp(-mid / 0.35) + 0.7 * rng.standard_normal(len(prof))


def make_dataset(n, seed, p=8):
    xData, dIData, prof = make_profiles(n, seed)
    X = make_design(xData, dIData, prof, seed, p)
    z = make_response(xData, dIData, prof, X, seed)
    return dict(xData=xData, dIData=dIData, prof=prof, XData=X, zData=z, n=n, p=p, seed=seed)


def default_pars(kind="matern0.5", nonstat=False):
    """Natural-scale parameter list (the C ABI's wire format), SURVEY.md 8(d)."""
    P = dict(ax=10.0, nux=0.5, ad=0.35, nud=0.5, cx0=0.10, cx1=0.25, cd1=0.15, cxd0=0.4, cxd1=0.6, cme=0.05,
             sdfdPars_cd1=np.zeros(0), sdfdPars_cxd0=np.zeros(0), sdfdPars_cxd1=np.zeros(0))
    modelx, types, cmeOpt = "matern", (0, 0, 0), 1
    if kind == "matern1.5":
        P["nux"] = 1.5
    elif kind == "matern0.8":
        P["nux"] = 0.8
    elif kind == "edgeroi":          # egEdgeroi.R:83-97
        modelx, types = "spherical", (-9, 0, 0)
        P.update(nux=float("nan"), cd1=0.0, ax=25.0)
    elif kind == "wendland":
        modelx = "wendland"; P.update(nux=2.4, ax=30.0)
    elif kind == "nugget":
        modelx = "nugget"; P.update(cx1=0.0, cxd0=1.0, cxd1=0.0, ax=float("nan"), nux=float("nan"))
    elif kind != "matern0.5":
        raise ValueError(kind)
    if nonstat:
        types = tuple(-1 if t == 0 else t for t in types)
        for k, t in zip(("sdfdPars_cd1", "sdfdPars_cxd0", "sdfdPars_cxd1"), types):
            if t == -1:
                P[k] = np.array([0.6, 1.7])
    return P, modelx, types, cmeOpt


def make_blocks(xData, prof, n_blocks, seed):
    """Voronoi blocks of profile locations around random centroids + adjacency of centroids
    (compositeLikIAK3D.R:829-1133 is one-off host preprocessing; here: nearest-centroid assignment and
    Delaunay adjacency of the centroids via scipy).  Returns dict(listBlocks, subsetPairsAdj,
    subsetPairsNonadj) with 0-based indices."""
    from scipy.spatial import Delaunay
    rng = np.random.default_rng(seed + 31)
    nprof = int(prof.max()) + 1
    first = np.zeros(nprof, np.int64); first[prof[::-1]] = np.arange(len(prof))[::-1]
    pxy = xData[first]
    cen = pxy[rng.choice(nprof, n_blocks, replace=False)]
    for _ in range(8):                                      # a few Lloyd steps to balance the blocks
        d2 = ((pxy[:, None, :] - cen[None, :, :]) ** 2).sum(-1)
        lab = d2.argmin(1)
        for b in range(n_blocks):
            if np.any(lab == b):
                cen[b] = pxy[lab == b].mean(0)
    d2 = ((pxy[:, None, :] - cen[None, :, :]) ** 2).sum(-1)
    lab = d2.argmin(1)
    rows = lab[prof]
    blocks = [np.nonzero(rows == b)[0] for b in range(n_blocks)]
    if n_blocks >= 3:
        tri = Delaunay(cen)
        adj = set()
        for s in tri.simplices:
            for a in range(3):
                i, j = sorted((int(s[a]), int(s[(a + 1) % 3])))
                adj.add((i, j))
    else:
        adj = {(i, j) for i in range(n_blocks) for j in range(i + 1, n_blocks)}
    adj = np.array(sorted(adj), np.int64).reshape(-1, 2)
    aset = {tuple(r) for r in adj}
    non = np.array([(i, j) for i in range(n_blocks) for j in range(i + 1, n_blocks) if (i, j) not in aset],
                   np.int64).reshape(-1, 2)
    return dict(listBlocks=blocks, subsetPairsAdj=adj, subsetPairsNonadj=non, centroids=cen)


def make_grid(n_side, extent_km=100.0):
    """n_side x n_side cell centres over the data square (P1M: n_side = 1000 -> 100 m cells)."""
    c = (np.arange(n_side) + 0.5) * (extent_km / n_side)
    gx, gy = np.meshgrid(c, c, indexing="xy")
    return np.column_stack([gx.ravel(), gy.ravel()])


def grid_design(xMap, dI_row, seed, p=8):
    """Design rows for map cells at one depth interval (makeXvX is an input of the hot path): smooth
    pseudo-covariates of location + the same three depth columns as :func:`make_design`."""
    rng = np.random.default_rng(seed + 7)
    k = rng.uniform(0.02, 0.2, (p - 4, 2)); ph = rng.uniform(0, 2 * np.pi, p - 4)
    cov = np.cos(xMap @ k.T + ph[None, :])
    dcol = _depth_cols(np.asarray(dI_row, float).reshape(1, 2))
    return np.column_stack([np.ones(len(xMap)), cov, np.repeat(dcol, len(xMap), axis=0)])
