"""Host mirror of the reference's fixed-effect design matrix for depth-interval supports: makeXvX (makeXvX.R:1) with
lnTfmdData = FALSE, evaluated on the GPU (csrc/design.cu).  The model description (column names for the 'mlr' branch,
rules for the 'cubist' branch) is flattened once into the specification `iak3d_design_create` takes; the per-support
arithmetic is the kernel's.

`makeXvX(covData, dIData, modelX, ...)` keeps the reference's arguments and returns dict(X, vXU = None, namesX, XLims);
`DesignModel` is the reusable handle predictIAK3D's grid staging takes (KeptFit.stage_grid)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import dptr, f64


def mlr_names(modelX, covNames, covLevels):
    """namesX of a numeric modelX (makeXvX.R:38-80): modelX = 0 covariates only; 0.5 + d; 1 + d.cov; 1.5 + d2; ..."""
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


class DesignModel:
    """modelX + optionsModelX (allKnotsd, opt_dSpline) + nDiscPts + XLims of one makeXvX call pattern, resident on the GPU.

    mlr:    modelX a number or a list of column names ('const', covariate names, 'd', 'd2', 'd3', 'd.cov', 'd2.cov',
            'd3.cov', 'dSpline.k'); covNames = names of the covariate columns in the order the covariate matrix will have;
            covLevels[name] = number of levels of a factor covariate (integer codes 1..L, treatment coding).
    cubist: modelX a dict(namesX, namesDataFit (column order of the covariate matrix, incl. 'dIMidPts'), rules, comms4XCubist,
            committees): rules = [dict(committee, conds = [(ivariable, '<=' | '>' | 'in', value or tuple)], const, coef_iv)],
            0-based variable indices -- the content of cubistModel$listRules / listCoeffs_iv / dfCoeffs after cubist2XSetup."""

    def __init__(self, modelX, covNames=(), covLevels=None, allKnotsd=(), opt_dSpline=0, nDiscPts=10, XLims=None, ctx=None):
        self.ctx = ctx or _lib.default_context()
        covLevels = covLevels or {}
        ak = np.asarray(allKnotsd if allKnotsd is not None else (), float).reshape(-1)
        if ak.size == 1:
            raise ValueError("must input at least two boundary knots in allKnotsd")
        nint = max(ak.size - 2, 0)
        cubist = isinstance(modelX, dict)
        if cubist:
            namesX = list(modelX["namesX"])
            covNames = list(modelX["namesDataFit"])
        elif isinstance(modelX, (int, float)):
            namesX = mlr_names(modelX, list(covNames), covLevels)
        else:
            namesX = list(modelX)
        if ak.size > 0 and not cubist:
            namesX = [nm for nm in namesX if nm not in ("d", "d2", "d3")]             # makeXvX.R:146-151
        self.namesX, self.covNames = namesX, list(covNames)
        p = len(namesX)
        spl_kind = 0 if (not cubist or not opt_dSpline) else 1
        ipSpline = [j for j, nm in enumerate(namesX) if nm.startswith("dSpline.")] if ak.size > 0 else []
        nspl = len(ipSpline)
        if nspl and nspl != (nint + 3 if spl_kind == 0 else nint):
            raise ValueError("the dSpline.* columns do not match the knots")
        has_lims = XLims is not None
        if has_lims:
            XL = np.array(XLims, float).reshape(2, p)
            infB = bool(np.min(XL[0]) == -np.inf and np.max(XL[1]) == np.inf)       # makeXvX.R:190-194
        else:
            XL = np.vstack([np.full(p, -np.inf), np.full(p, np.inf)]); infB = True
        ispec = [1 if cubist else 0, p, len(covNames), nspl, spl_kind, nint, int(nDiscPts), 1 if has_lims else 0, 1 if infB else 0]
        dspec = [float(ak[0]) if ak.size else 0.0, float(ak[-1]) if ak.size else 1.0] + [float(v) for v in ak[1:-1]]
        dspec += list(XL[0]) + list(XL[1])
        if not cubist:
            cidx = {nm: k for k, nm in enumerate(covNames)}
            kinds, covs, levels = [], [], []
            lvl = 2
            for j, nm in enumerate(namesX):
                if nm == "const": k, base = 0, None
                elif nm == "d": k, base = 1, None
                elif nm == "d2": k, base = 2, None
                elif nm == "d3": k, base = 3, None
                elif nm.startswith("dSpline."): k, base = 4, None
                elif nm.startswith("d2."): k, base = 2, nm[3:]
                elif nm.startswith("d3."): k, base = 3, nm[3:]
                elif nm.startswith("d."): k, base = 1, nm[2:]
                else: k, base = 0, nm
                level = float("nan")
                if k == 4:
                    cv = ipSpline.index(j)
                elif base is None:
                    cv = -1
                else:
                    if base not in cidx:
                        raise KeyError(f"covariate {base!r} of column {nm!r} is not in covNames")
                    cv = cidx[base]
                    if covLevels.get(base):
                        if j == 0 or namesX[j] != namesX[j - 1]:
                            lvl = 2
                        level = float(lvl); lvl += 1
                kinds.append(k); covs.append(cv); levels.append(level)
            # limits of the depth columns are recycled from the first element of XLims over t(X[, depth columns]) (makeXvX.R:327-328)
            order = [j for j, nm in enumerate(namesX) if nm == "d"] + [j for j, nm in enumerate(namesX) if nm.startswith("d.")]
            order += [j for j, nm in enumerate(namesX) if nm == "d2"] + [j for j, nm in enumerate(namesX) if nm.startswith("d2.")]
            order += [j for j, nm in enumerate(namesX) if nm == "d3"] + [j for j, nm in enumerate(namesX) if nm.startswith("d3.")]
            clampcol = [-1] * p
            for kpos, j in enumerate(order):
                clampcol[j] = kpos
            for j in range(p):
                ispec += [kinds[j], covs[j], clampcol[j]]
                dspec.append(levels[j])
            self.depth_cols = order
        else:
            rules = modelX["rules"]
            dcol = covNames.index("dIMidPts")
            comms = np.asarray(modelX["comms4XCubist"], int)
            rule_comm = np.asarray([r["committee"] for r in rules], int)
            pc = sum((1 if r["const"] else 0) + len(r["coef_iv"]) for r in rules)
            if pc + nspl != p or comms.size != pc:
                raise ValueError("rule columns, comms4XCubist and namesX disagree")
            # committee blocks in the order cubist2XGivenSetup walks them (cubist2XIAK3D.R:139-152)
            blocks, col_blk, iR = [], [-1] * pc, 0
            for iC in range(1, int(modelX["committees"]) + 1):
                jThis = np.where(comms == iC)[0]
                nR = int(np.sum(rule_comm == iC))
                if jThis.size > 0 and nR > 0:
                    for j in jThis: col_blk[int(j)] = len(blocks)
                    blocks.append((iR, iR + nR)); iR += nR
            breaks = np.unique(np.asarray([v for r in rules for (iv, dr, v) in r["conds"] if iv == dcol and dr != "in"], float))
            ispec += [pc, len(rules), dcol, int(len(np.unique(comms))), int(breaks.size), len(blocks)]
            for r in rules:
                ispec += [1 if r["const"] else 0, len(r["conds"]), len(r["coef_iv"])]
                for (iv, dr, v) in r["conds"]:
                    if dr == "in":
                        vals = [float(x) for x in np.atleast_1d(v)]
                        ispec += [int(iv), 2, len(vals)]; dspec.append(0.0); dspec += vals
                    else:
                        ispec += [int(iv), 0 if dr == "<=" else 1, 0]; dspec.append(float(v))
                ispec += [int(v) for v in r["coef_iv"]]
            ispec += col_blk
            for (r0, r1) in blocks: ispec += [r0, r1]
            dspec += [float(b) for b in breaks]
        self.p, self.ncov, self.nspl, self.cubist, self.ipSpline = p, len(covNames), nspl, cubist, ipSpline
        ia = np.asarray(ispec, np.int32); da = np.asarray(dspec, np.float64)
        h = C.c_void_p()
        st = self.ctx.lib.iak3d_design_create(self.ctx.h, _lib.iptr(ia), ia.size, dptr(da), da.size, C.byref(h))
        self.ctx.check(st)
        self.h = h
        self._close_order = 0
        self.ctx.adopt(self)

    def eval(self, covs, dIData, want_spline_limits=False):
        """X (m x p) of m supports: covs m x ncov (column order covNames), dIData m x 2."""
        dI = f64(np.asarray(dIData, float).reshape(-1, 2))
        m = dI.shape[0]
        cv = f64(np.asarray(covs, float).reshape(m, self.ncov)) if self.ncov > 0 else None
        X = np.empty((m, self.p), order="F")
        mn = mx = None
        if want_spline_limits and self.nspl:
            mn = np.empty((m, self.nspl), order="F"); mx = np.empty((m, self.nspl), order="F")
        self.ctx.check(self.ctx.lib.iak3d_design_eval(self.h, m, dptr(cv), dptr(dI), dptr(X), dptr(mn), dptr(mx)))
        return (X, mn, mx) if want_spline_limits else X

    def eval_lnN(self, covs, dIData, iU, want_limits=False):
        """lnTfmdData = TRUE: X (m x p) and vXU ((m pU) x pU) of m supports; iU = 0-based columns that vary inside a support."""
        dI = f64(np.asarray(dIData, float).reshape(-1, 2))
        m = dI.shape[0]
        cv = f64(np.asarray(covs, float).reshape(m, self.ncov)) if self.ncov > 0 else None
        iU = np.ascontiguousarray(np.asarray(iU, np.int32).reshape(-1)); pU = iU.size
        X = np.empty((m, self.p), order="F"); V = np.empty((m * pU, pU), order="F")
        mn = np.empty((m, self.p), order="F") if want_limits else None
        mx = np.empty((m, self.p), order="F") if want_limits else None
        self.ctx.check(self.ctx.lib.iak3d_design_eval_lnN(self.h, m, dptr(cv), dptr(dI), pU, _lib.iptr(iU), dptr(X), dptr(V), dptr(mn), dptr(mx)))
        return (X, V, mn, mx) if want_limits else (X, V)

    def close(self):
        if getattr(self, "h", None) is not None and self.h.value:
            self.ctx.lib.iak3d_design_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _cov_matrix(covData, covNames, n):
    if covData is None or len(covNames) == 0:
        return np.zeros((n, 0))
    if isinstance(covData, dict):
        return np.column_stack([np.asarray(covData[nm], float) for nm in covNames])
    return np.asarray(covData, float).reshape(n, len(covNames))


def makeXvX(covData=None, dIData=None, modelX=0, allKnotsd=(), opt_dSpline=0, iU=None, nDiscPts=100, lnTfmdData=False, XLims=None,
            covLevels=None, ctx=None):
    """makeXvX (makeXvX.R:1): design matrix of interval supports.  covData: dict name -> column (mlr; factor covariates as
    integer codes with covLevels[name] = number of levels) or an n x nvars array in modelX['namesDataFit'] order (cubist).
    lnTfmdData = TRUE (the within-support covariance vXU) is not built on the device."""
    dI = np.asarray(dIData, float).reshape(-1, 2)
    n = dI.shape[0]
    cubist = isinstance(modelX, dict)
    covNames = list(modelX["namesDataFit"]) if cubist else (list(covData.keys()) if isinstance(covData, dict) else [])
    dm = DesignModel(modelX, covNames, covLevels, allKnotsd, opt_dSpline, nDiscPts, XLims, ctx=ctx)
    try:
        cv = _cov_matrix(covData, covNames, n)
        if lnTfmdData:
            # makeXvX.R:246-290 ('mlr'), :437-557 ('cubist'): point designs at nDiscPts depths -> mean and var(); iU (:566-600)
            nm = dm.namesX
            if iU is None:
                iU = (list(range(dm.p)) if cubist else
                      [j for j, s_ in enumerate(nm) if s_ in ("d", "d2") or s_.startswith(("d.", "d2.", "d3.", "dSpline."))])
            iU = [int(j) for j in iU]
            if XLims is not None:
                X, V = dm.eval_lnN(cv, dI, iU)
                XL = np.array(XLims, float)
            else:
                X, V, mn, mx = dm.eval_lnN(cv, dI, iU, want_limits=True)
                with np.errstate(all="ignore"):
                    lo = np.where(np.all(np.isnan(mn), axis=0), np.inf, np.nanmin(np.where(np.isnan(mn), np.inf, mn), axis=0))
                    hi = np.where(np.all(np.isnan(mx), axis=0), -np.inf, np.nanmax(np.where(np.isnan(mx), -np.inf, mx), axis=0))
                XL = np.vstack([lo, hi])
                for j in dm.ipSpline:
                    XL[0, j], XL[1, j] = -np.inf, np.inf                # folded into the (-Inf, Inf) the spline columns start with (:176-179)
            return dict(X=np.asarray(X), vXU=np.asarray(V), iU=iU, namesX=nm, XLims=XL)
        if XLims is not None:
            X = dm.eval(cv, dI)
            XL = np.array(XLims, float)
        else:
            # setXLims (makeXvX.R:172-183): limits are an OUTPUT, nothing is clamped
            X, mn, mx = dm.eval(cv, dI, want_spline_limits=True)
            p = dm.p
            XL = np.vstack([np.full(p, np.inf), np.full(p, -np.inf)])
            if cubist:
                for j in range(p):
                    # the spline columns' limits come from their trapezoid values, taken before makeXvX.R:421-430 replaces them
                    col = mn[:, dm.ipSpline.index(j)] if j in dm.ipSpline else X[:, j]
                    nz = col[col != 0]
                    XL[0, j] = nz.min() if nz.size else np.inf
                    XL[1, j] = nz.max() if nz.size else -np.inf
            else:
                for j in dm.ipSpline: XL[0, j], XL[1, j] = -np.inf, np.inf
                for j in dm.depth_cols:
                    XL[0, j], XL[1, j] = X[:, j].min(), X[:, j].max()
                for k, j in enumerate(dm.ipSpline):              # min / max of the non-zero POINT values, folded into (-Inf, Inf)
                    XL[0, j] = min(-np.inf, mn[:, k].min()); XL[1, j] = max(np.inf, mx[:, k].max())
        return dict(X=np.asarray(X), vXU=None, namesX=dm.namesX, XLims=XL)
    finally:
        dm.close()
