"""CPU restatement of the reference's fixed-effect design matrix for depth-interval supports -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module; the product (iak3d_b200/) never does.

What is restated (NumPy loops in the reference's own order of operations):
  makeXvX_lnN        makeXvX.R:246-290 ('mlr') and :437-600 ('cubist') with lnTfmdData = TRUE: X = means and vXU = var() of the point
                     designs at nDiscPts depths per support, reduced to the columns iU that vary inside a support
  makeXvX            makeXvX.R:1-708, the two branches with lnTfmdData = FALSE that predictIAK3D reaches (predictIAK3D.R:155):
                       * 'mlr'    (:212-353, the "quicker version" :294-353)  const / covariate / factor-dummy columns times
                                   the depth moments of the interval, B-spline-of-depth columns averaged over nDiscPts points
                       * 'cubist' (:360-434)  piecewise-linear-in-depth rule design averaged exactly by trapezoids between
                                   the rules' depth breaks (re-evaluated at break -/+ 1e-4), spline columns averaged over nDiscPts
  cubist2XGivenSetup cubist2XIAK3D.R:46-176   rule membership, per-rule columns, normalisation within committees
  allKnotsd2X        cubist2XIAK3D.R:187-208  bs() or the clamped natural spline basis without its constant
  getAlldBreaks      cubist2XIAK3D.R:889-900
  replaceLT/GT, minNonZero/maxNonZero  makeXvX.R:669-703

Third-party arithmetic the reference delegates to: splines::bs (R's recommended package `splines`, version unpinned) --
cubic B-spline basis on the knot vector (b1 x4, internal knots, b2 x4) without its first column; restated through
scipy.interpolate.BSpline (de Boor's recurrence).  Supports beyond the boundary knots are not restated: R extrapolates from a
pivot whose position changed between R versions; this module returns NaN there.

Reference quirks kept (each one is what the R code computes, not what its comments say):
  * the 'd2' / 'd3' columns multiply by  a^2 + ab + b^2  and  a^3 + a^2 b + a b^2 + b^3  WITHOUT the 1/3 and 1/4 of the interval
    means (makeXvX.R:303, 308);
  * when limits are applied in the mlr branch the lower / upper limits are recycled over t(X[, depth columns]) from the START of
    XLims (llim = XLims[1,] against t(X[, id123Tmp]), makeXvX.R:327-328), i.e. entry (row i, k-th depth column in the order id,
    idInt, id2, id2Int, id3, id3Int; 0-based) is clamped by the limits of column (i K + k) mod p, not by its own;
  * one -Inf lower and one +Inf upper limit anywhere in XLims switch ALL clamping off (infBnds, :190-194);
  * limits never move exact zeros (replaceZeros = FALSE).

Parity status: unpinned against R (no R here, no fixtures in the reference); pinned by construction checks in
tests/test_design.py (trapezoid average == dense quadrature of the piecewise-linear design, B-spline partition of unity,
depth moments against direct formulas).
"""
from __future__ import annotations

import numpy as np
from scipy.interpolate import BSpline

from . import iak3d_oracle as orc


# ---------------------------------------------------------------------------------------------------------------------
# helpers (makeXvX.R:669-703)
# ---------------------------------------------------------------------------------------------------------------------
def replaceLT(x, llim, replaceZeros=True):
    x = np.array(x, float)
    m = (x < llim) if replaceZeros else ((x < llim) & (x != 0))
    x[m] = np.broadcast_to(llim, x.shape)[m]
    return x


def replaceGT(x, ulim, replaceZeros=True):
    x = np.array(x, float)
    m = (x > ulim) if replaceZeros else ((x > ulim) & (x != 0))
    x[m] = np.broadcast_to(ulim, x.shape)[m]
    return x


def minNonZero(x, all0Val=np.inf):
    x = np.asarray(x, float); x = x[x != 0]
    return all0Val if x.size == 0 else float(np.min(x))


def maxNonZero(x, all0Val=-np.inf):
    x = np.asarray(x, float); x = x[x != 0]
    return all0Val if x.size == 0 else float(np.max(x))


def bs_basis(x, intKnots, bdryKnots):
    """splines::bs(x, knots = intKnots, degree = 3, intercept = F, Boundary.knots = bdryKnots): n x (len(intKnots) + 3)."""
    x = np.asarray(x, float).reshape(-1)
    b1, b2 = float(bdryKnots[0]), float(bdryKnots[1])
    ik = np.asarray(intKnots, float).reshape(-1)
    t = np.concatenate([[b1] * 4, ik, [b2] * 4])
    nb = ik.size + 4
    out = np.full((x.size, nb), np.nan)
    ok = (x >= b1) & (x <= b2)
    if np.any(ok):
        out[ok] = BSpline(t, np.eye(nb), 3, extrapolate=False)(x[ok])
        # the right boundary belongs to the last interval (R: splineDesign is right-continuous except at the last knot)
        out[ok & (x == b2)] = np.eye(nb)[nb - 1]
    return out[:, 1:]


def allKnotsd2X(dIMidPts, allKnotsd, opt_dSpline=0):
    """cubist2XIAK3D.R:187-208."""
    d = np.asarray(dIMidPts, float).reshape(-1)
    allKnotsd = np.asarray(allKnotsd, float).reshape(-1)
    if allKnotsd.size == 0:
        return np.zeros((d.size, 0))
    intKnots = allKnotsd[1:-1]
    bdry = (allKnotsd[0], allKnotsd[-1])
    if (opt_dSpline or 0) == 0:
        return bs_basis(d, intKnots, bdry)
    return orc.makeXnsClampedUG0(d, bdry, intKnots)[:, 1:]


def seq_by(a, b, nDiscPts):
    """seq(a, b, (b - a) / (nDiscPts - 1)): from + (0:n) * by with n = floor((to - from) / by + 1e-10), the last value
    never beyond `to` (R's seq.default)."""
    if nDiscPts <= 1:
        return np.array([0.5 * (a + b)])
    by = (b - a) / (nDiscPts - 1)
    n = int(np.floor((b - a) / by + 1e-10))
    v = a + np.arange(n + 1) * by
    return np.minimum(v, b)


# ---------------------------------------------------------------------------------------------------------------------
# Cubist rule design (cubist2XIAK3D.R:46-176)
# ---------------------------------------------------------------------------------------------------------------------
def cubist2XGivenSetup(cubistModel, dataFit):
    """dataFit: n x nvars numeric array in the column order cubistModel['namesDataFit'].  cubistModel fields (0-based
    indices): 'rules' = list of dict(committee, conds = [(ivariable, dir, val)] with dir in '<=', '>', 'in' (val a tuple of
    category codes), const (bool), coef_iv = [ivariable, ...]); 'comms4XCubist' (per column, 1-based committee numbers as in
    R); 'allKnotsd', 'opt_dSpline' (optional), 'dcol' = column of dIMidPts."""
    D = np.asarray(dataFit, float)
    n = D.shape[0]
    rules = cubistModel["rules"]
    nRules = len(rules)
    pc = sum((1 if r["const"] else 0) + len(r["coef_iv"]) for r in rules)
    X = np.zeros((n, pc))
    mat = np.zeros((n, nRules))
    j0 = 0
    for iR, r in enumerate(rules):
        if nRules > 1:
            m = np.ones(n, bool)
            for (iv, dr, val) in r["conds"]:
                if dr == "<=":
                    m &= D[:, iv] <= val
                elif dr == ">":
                    m &= D[:, iv] > val
                else:
                    m &= np.isin(D[:, iv], np.asarray(val, float))
        else:
            m = np.ones(n, bool)
        mat[m, iR] = 1
        nc = (1 if r["const"] else 0) + len(r["coef_iv"])
        X[m, j0] = 1
        if r["const"]:
            if len(r["coef_iv"]) > 0:
                X[np.ix_(m, range(j0 + 1, j0 + nc))] = D[np.ix_(m, r["coef_iv"])]
        else:
            X[np.ix_(m, range(j0, j0 + nc))] = D[np.ix_(m, r["coef_iv"])]
        j0 += nc
    if nRules > 1:
        comms = np.asarray(cubistModel["comms4XCubist"], int)
        rule_comm = np.asarray([r["committee"] for r in rules], int)
        iRTmp = 0
        for iC in range(1, int(cubistModel["committees"]) + 1):
            jThis = np.where(comms == iC)[0]
            nRulesThis = int(np.sum(rule_comm == iC))
            if jThis.size > 0 and nRulesThis > 0:
                cnt = mat[:, iRTmp:iRTmp + nRulesThis].sum(axis=1)
                iTmp = cnt > 0
                X[np.ix_(iTmp, jThis)] = X[np.ix_(iTmp, jThis)] / cnt[iTmp][:, None]
                iRTmp += nRulesThis
        X = X / len(np.unique(comms))
    ak = cubistModel.get("allKnotsd", ())
    if ak is not None and len(ak) > 0:
        X = np.column_stack([X, allKnotsd2X(D[:, cubistModel["dcol"]], ak, cubistModel.get("opt_dSpline", 0))])
    X[np.isnan(X).any(axis=1)] = np.nan                                    # cubist2X: rows with any NA -> all NA (:37-40)
    return X, mat


def getAlldBreaks(cubistModel):
    v = [val for r in cubistModel["rules"] for (iv, dr, val) in r["conds"] if iv == cubistModel["dcol"] and dr != "in"]
    return np.unique(np.asarray(v, float))


# ---------------------------------------------------------------------------------------------------------------------
# makeXvX (makeXvX.R)
# ---------------------------------------------------------------------------------------------------------------------
def mlr_names(modelX, covNames, covLevels):
    """namesX of a numeric modelX (makeXvX.R:38-80).  covLevels[name] = number of levels of a factor covariate, else None."""
    names = ["const"]
    for o in range(int(np.floor(modelX)) + 1):
        if o > 3:
            raise ValueError("makeXvX has only been coded for up to cubic functions of depth")
        ints = ["", "d.", "d2.", "d3."][o]
        uni = ["d", "d2", "d3", None][o]
        for nm in covNames:
            reps = (covLevels.get(nm) - 1) if covLevels.get(nm) else 1
            names += [ints + nm] * reps
        if modelX > o:
            names.append(uni)
    return names


def _rvar(XT):
    """R's var() of the rows of XT: (n - 1)-denominator sample covariance, two-pass."""
    m = XT.mean(axis=0)
    Z = XT - m[None, :]
    return (Z.T @ Z) / (XT.shape[0] - 1)


def makeXvX_lnN(covData, dIData, modelX, allKnotsd=(), opt_dSpline=0, nDiscPts=100, XLims=None, covLevels=None, iU=None):
    """lnTfmdData = TRUE (makeXvX.R:246-290 for 'mlr', :437-557 for 'cubist'): every support is discretised into nDiscPts point
    depths; X = column means of the point designs, vXU = their sample covariance (var()), reduced to the columns iU that
    vary inside a support (:566-600).  Limits act on the POINT values (set mode: min / max of the non-zero point values).
    Returns dict(X, vXU ((n pU) x pU), iU (0-based), XLims, namesX)."""
    dI = np.asarray(dIData, float).reshape(-1, 2)
    n = dI.shape[0]
    covLevels = covLevels or {}
    allKnotsd = np.asarray(allKnotsd if allKnotsd is not None else (), float).reshape(-1)
    is_cubist = isinstance(modelX, dict)
    if is_cubist:
        namesX = list(modelX["namesX"])
    elif isinstance(modelX, (int, float)):
        namesX = mlr_names(modelX, list(covData.keys()) if covData else [], covLevels)
    else:
        namesX = list(modelX)
    if allKnotsd.size > 0 and not is_cubist:
        namesX = [nm for nm in namesX if nm not in ("d", "d2", "d3")]
    intKnots, bdry = allKnotsd[1:-1], ((allKnotsd[0], allKnotsd[-1]) if allKnotsd.size else (0.0, 1.0))
    p = len(namesX)
    ipSpline = [j for j, nm in enumerate(namesX) if nm.startswith("dSpline.")] if allKnotsd.size > 0 else []
    if XLims is None:
        setXLims = True
        XLims = np.empty((2, p)); XLims[0] = np.inf; XLims[1] = -np.inf
        XLims[0, ipSpline] = -np.inf; XLims[1, ipSpline] = np.inf
        infBnds = False
    else:
        setXLims = False
        XLims = np.array(XLims, float)
        infBnds = bool(np.min(XLims[0]) == -np.inf and np.max(XLims[1]) == np.inf)
    X = np.full((n, p), np.nan); V = np.full((n, p, p), np.nan)
    if not is_cubist:
        idx = lambda pred: [j for j, nm in enumerate(namesX) if pred(nm)]
        id1 = idx(lambda s: s == "d"); id2 = idx(lambda s: s == "d2"); id3 = idx(lambda s: s == "d3")
        idInt = idx(lambda s: s.startswith("d.")); id2Int = idx(lambda s: s.startswith("d2.")); id3Int = idx(lambda s: s.startswith("d3."))
        idSpline = idx(lambda s: s.startswith("dSpline."))
        names_d = list(namesX)
        for j in id1 + id2 + id3 + idSpline: names_d[j] = "const"
        for j in idInt: names_d[j] = names_d[j][2:]
        for j in id2Int + id3Int: names_d[j] = names_d[j][3:]
        X_d = np.full((n, p), np.nan)
        lvl = 2
        for j in range(p):
            if names_d[j] == "const":
                X_d[:, j] = 1.0
            else:
                c = np.asarray(covData[names_d[j]])
                if covLevels.get(names_d[j]):
                    if j == 0 or namesX[j] != namesX[j - 1]:
                        lvl = 2
                    X_d[:, j] = (c == lvl).astype(float); lvl += 1
                else:
                    X_d[:, j] = c.astype(float)
        for i in range(n):
            dd = seq_by(dI[i, 0], dI[i, 1], nDiscPts)
            XT = np.repeat(X_d[i:i + 1], dd.size, axis=0)                              # :248
            if id1: XT[:, id1] = XT[:, id1] * dd[:, None]
            if id2: XT[:, id2] = XT[:, id2] * (dd ** 2)[:, None]
            if id3: XT[:, id3] = XT[:, id3] * (dd ** 3)[:, None]
            if idInt: XT[:, idInt] = XT[:, idInt] * dd[:, None]
            if id2Int: XT[:, id2Int] = XT[:, id2Int] * (dd ** 2)[:, None]
            if id3Int: XT[:, id3Int] = XT[:, id3Int] * (dd ** 3)[:, None]
            if idSpline: XT[:, idSpline] = XT[:, idSpline] * bs_basis(dd, intKnots, bdry)
            if setXLims:
                XLims[0] = [minNonZero(np.append(XT[:, j], XLims[0, j])) for j in range(p)]
                XLims[1] = [maxNonZero(np.append(XT[:, j], XLims[1, j])) for j in range(p)]
            elif not infBnds:
                XT = replaceLT(XT, XLims[0][None, :], replaceZeros=False)
                XT = replaceGT(XT, XLims[1][None, :], replaceZeros=False)
            X[i] = XT.mean(axis=0)
            V[i] = _rvar(XT)
        if iU is None:
            iU = [j for j, nm in enumerate(namesX) if nm in ("d", "d2") or nm.startswith("d.") or nm.startswith("d2.")
                  or nm.startswith("d3.") or nm.startswith("dSpline.")]               # :570 ('d3' itself is not listed there)
    else:
        cm = dict(modelX); cm["allKnotsd"] = allKnotsd; cm["opt_dSpline"] = opt_dSpline
        D = np.array(covData, float)
        lin = np.arange(nDiscPts) * (1.0 / (nDiscPts - 1))                             # seq(0, 1, 1 / (nDiscPts - 1))
        for i in range(n):
            dd = dI[i, 0] + lin * (dI[i, 1] - dI[i, 0])
            Di = np.repeat(D[i:i + 1], nDiscPts, axis=0); Di[:, cm["dcol"]] = dd
            XT, _ = cubist2XGivenSetup(cm, Di)
            if setXLims:
                XLims[0] = [minNonZero(np.append(XT[:, j], XLims[0, j])) for j in range(p)]
                XLims[1] = [maxNonZero(np.append(XT[:, j], XLims[1, j])) for j in range(p)]
            elif not infBnds:
                XT = replaceLT(XT, XLims[0][None, :], replaceZeros=False)
                XT = replaceGT(XT, XLims[1][None, :], replaceZeros=False)
            X[i] = XT.mean(axis=0)
            V[i] = _rvar(XT - X[i][None, :])
        if iU is None:
            iU = list(range(p))
    iU = [int(j) for j in iU]
    vXU = V[:, iU][:, :, iU].reshape(n * len(iU), len(iU))
    return dict(X=X, vXU=vXU, iU=iU, XLims=XLims, namesX=namesX)


def makeXvX(covData, dIData, modelX, allKnotsd=(), opt_dSpline=0, nDiscPts=100, XLims=None, covLevels=None):
    """lnTfmdData = FALSE.  covData: dict name -> length-n array (factor covariates: integer codes 1..L with
    covLevels[name] = L; treatment coding, first dummy = level 2, makeXvX.R:232-241), or None.  modelX: number, list of
    column names, or a cubist model dict (see cubist2XGivenSetup; covData then is an n x nvars array whose column
    cubistModel['dcol'] is overwritten by the evaluation depth).  Returns dict(X, XLims, namesX)."""
    dI = np.asarray(dIData, float).reshape(-1, 2)
    n = dI.shape[0]
    covLevels = covLevels or {}
    allKnotsd = np.asarray(allKnotsd if allKnotsd is not None else (), float).reshape(-1)
    is_cubist = isinstance(modelX, dict)
    if is_cubist:
        namesX = list(modelX["namesX"])
    elif isinstance(modelX, (int, float)):
        namesX = mlr_names(modelX, list(covData.keys()) if covData else [], covLevels)
    else:
        namesX = list(modelX)
    if allKnotsd.size > 0:
        if allKnotsd.size < 2:
            raise ValueError("must input at least two boundary knots in allKnotsd")
        intKnots, bdry = allKnotsd[1:-1], (allKnotsd[0], allKnotsd[-1])
        if not is_cubist:
            namesX = [nm for nm in namesX if nm not in ("d", "d2", "d3")]                 # :146-151
    p = len(namesX)
    X = np.full((n, p), np.nan)
    ipSpline = [j for j, nm in enumerate(namesX) if nm.startswith("dSpline.")] if allKnotsd.size > 0 else []
    if XLims is None:
        setXLims = True
        XLims = np.empty((2, p)); XLims[0] = np.inf; XLims[1] = -np.inf
        XLims[0, ipSpline] = -np.inf; XLims[1, ipSpline] = np.inf
        infBnds = False
    else:
        setXLims = False
        XLims = np.array(XLims, float)
        infBnds = bool(np.min(XLims[0]) == -np.inf and np.max(XLims[1]) == np.inf)      # :190-194

    if not is_cubist:
        idx = lambda pred: [j for j, nm in enumerate(namesX) if pred(nm)]
        id1 = idx(lambda s: s == "d"); id2 = idx(lambda s: s == "d2"); id3 = idx(lambda s: s == "d3")
        idInt = idx(lambda s: s.startswith("d.")); id2Int = idx(lambda s: s.startswith("d2.")); id3Int = idx(lambda s: s.startswith("d3."))
        idSpline = idx(lambda s: s.startswith("dSpline."))
        idAny = id1 + id2 + id3 + idInt + id2Int + id3Int + idSpline
        iNod = [j for j in range(p) if j not in idAny]
        names_d = list(namesX)
        for j in id1 + id2 + id3 + idSpline: names_d[j] = "const"
        for j in idInt: names_d[j] = names_d[j][2:]
        for j in id2Int + id3Int: names_d[j] = names_d[j][3:]
        X_d = np.full((n, p), np.nan)
        lvl = 2
        for j in range(p):
            if names_d[j] == "const":
                X_d[:, j] = 1.0
            else:
                c = np.asarray(covData[names_d[j]])
                if covLevels.get(names_d[j]):
                    if j == 0 or namesX[j] != namesX[j - 1]:
                        lvl = 2
                    X_d[:, j] = (c == lvl).astype(float)
                    lvl += 1
                else:
                    X_d[:, j] = c.astype(float)
        X[:, iNod] = X_d[:, iNod]
        a, b = dI[:, 0], dI[:, 1]
        id1T, id2T, id3T = id1 + idInt, id2 + id2Int, id3 + id3Int
        if id1T: X[:, id1T] = X_d[:, id1T] * ((a + b) / 2.0)[:, None]                     # rowMeans
        if id2T: X[:, id2T] = X_d[:, id2T] * (a ** 2 + a * b + b ** 2)[:, None]
        if id3T: X[:, id3T] = X_d[:, id3T] * (a ** 3 + (a ** 2) * b + a * (b ** 2) + b ** 3)[:, None]
        id123 = id1T + id2T + id3T
        if id123:
            if setXLims:
                XLims[0, id123] = X[:, id123].min(axis=0); XLims[1, id123] = X[:, id123].max(axis=0)
            elif not infBnds:
                # mapply over t(X[, id123]) recycles XLims[1,] / XLims[2,] from their first element (:327-328)
                K = len(id123)
                flat = X[:, id123].reshape(-1)                                           # row by row == t() column-major
                ll = np.resize(XLims[0], flat.size); ul = np.resize(XLims[1], flat.size)
                flat = replaceLT(flat, ll, replaceZeros=False)
                flat = replaceGT(flat, ul, replaceZeros=False)
                X[:, id123] = flat.reshape(n, K)
        if idSpline:
            for i in range(n):
                dd = seq_by(dI[i, 0], dI[i, 1], nDiscPts)
                XT = bs_basis(dd, intKnots, bdry)
                if setXLims:
                    XLims[0, idSpline] = [minNonZero(np.append(XT[:, k], XLims[0, idSpline[k]])) for k in range(len(idSpline))]
                    XLims[1, idSpline] = [maxNonZero(np.append(XT[:, k], XLims[1, idSpline[k]])) for k in range(len(idSpline))]
                elif not infBnds:
                    XT = replaceLT(XT, XLims[0, idSpline][None, :], replaceZeros=False)
                    XT = replaceGT(XT, XLims[1, idSpline][None, :], replaceZeros=False)
                X[i, idSpline] = XT.mean(axis=0)
        return dict(X=X, XLims=XLims, namesX=namesX)

    # ---- cubist, lnTfmdData = FALSE (:360-434) ----
    cm = dict(modelX); cm["allKnotsd"] = allKnotsd; cm["opt_dSpline"] = opt_dSpline
    D = np.array(covData, float)
    dcol = cm["dcol"]
    breaks = getAlldBreaks(cm)
    D[:, dcol] = dI[:, 0]
    XT, _ = cubist2XGivenSetup(cm, D)
    mw = np.zeros_like(XT)
    lb = dI[:, 0].copy()
    for brk in breaks:
        it = np.where((dI[:, 0] < brk) & (dI[:, 1] > brk))[0]
        if it.size:
            w = (brk - lb[it]) / (dI[it, 1] - dI[it, 0])
            Dt = D[it].copy(); Dt[:, dcol] = brk - 0.0001
            XB, _ = cubist2XGivenSetup(cm, Dt)
            mw[it] = mw[it] + 0.5 * (XB + XT[it]) * w[:, None]
            lb[it] = brk
            Dt = D[it].copy(); Dt[:, dcol] = brk + 0.0001
            XT[it], _ = cubist2XGivenSetup(cm, Dt)
    w = (dI[:, 1] - lb) / (dI[:, 1] - dI[:, 0])
    D[:, dcol] = dI[:, 1]
    XB, _ = cubist2XGivenSetup(cm, D)
    mw = mw + 0.5 * (XB + XT) * w[:, None]
    X = mw
    if setXLims:
        XLims[0] = [minNonZero(X[:, j]) for j in range(p)]; XLims[1] = [maxNonZero(X[:, j]) for j in range(p)]
    elif not infBnds:
        X = replaceLT(X, XLims[0][None, :], replaceZeros=False)
        X = replaceGT(X, XLims[1][None, :], replaceZeros=False)
    if allKnotsd.size > 0:
        cols = [j for j, nm in enumerate(namesX) if nm.startswith("dSpline.")]
        acc = np.zeros((n, len(cols)))
        for idisc in range(1, nDiscPts + 1):
            acc = acc + allKnotsd2X(dI[:, 0] + ((idisc - 1) / (nDiscPts - 1)) * (dI[:, 1] - dI[:, 0]), allKnotsd, opt_dSpline)
        X[:, cols] = acc / nDiscPts
    return dict(X=X, XLims=XLims, namesX=namesX)
