"""Host-side mirror of the reference's interface for the hot path (same names, argument meaning and error
behaviour as the R functions), driving libiak3d.so through the C ABI.  Everything O(n^2) or larger runs on the
GPU; what stays here is what stays in R in the reference-side integration: the parameter
transforms and the O(p^3) REML / ML tail.

    setupIAK3D / setupIAK3D2      iaCovMatern.R:413-548, 581-678
    readParsIAK3D, logitab, tauTfm fitIAK3D.R:1332-1417, 1587-1635
    setCIAK3D / setCIAK3D2        fitIAK3D.R:955-1112, 1117-1252
    lndetANDinvCb                 optim_functions.R:268-312
    nllIAK3D, gradnllIAK3D        fitIAK3D.R:670-949, 1865-1886
    nllIAK3D_CL                   compositeLikIAK3D.R:6-307
    predictIAK3D                  predictIAK3D.R:5-287
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _lib
from ._lib import BAD_PARS, NOT_SPD, OK, Context, Iak3dError, Pars, default_context, dptr, f64, iptr, load, make_model, make_pars, spl_ptrs

BADNLL = 9e99   # fitIAK3D.R:685


# ---------------------------------------------------------------------------------------------
# parameter transforms (host; fitIAK3D.R:1332-1417, 1474-1504, 1587-1635)
# ---------------------------------------------------------------------------------------------
def logitab(z, a=0.0, b=1.0, invt=False):
    z = np.asarray(z, dtype=float)
    if invt:
        with np.errstate(over="ignore"):
            return (1.0 / (1.0 + np.exp(-z))) * (b - a) + a
    u = (z - a) / (b - a)
    return np.log(u / (1.0 - u))


def tauTfm(parsIn, parBnds, invt=False):
    v = np.asarray(parsIn, dtype=float).reshape(-1)
    if v.size == 0:
        return np.zeros(0)
    if v.size != 2:
        raise ValueError("tauTfm: not ready for this number of parameters")
    lo, hi = parBnds["tau2Bnds"]
    maxd = parBnds["maxd"]
    if invt:
        t2 = float(logitab(v[1], lo, hi, invt=True))
        with np.errstate(over="ignore", invalid="ignore"):
            return np.array([float((1.0 - np.exp(v[0])) / (1.0 - np.exp(-t2 * maxd))), t2])
    return np.array([math.log(1.0 - v[0] * (1.0 - math.exp(-v[1] * maxd))), float(logitab(v[1], lo, hi))])


def readParsIAK3D(pars, modelx, nud, sdfdType_cd1, sdfdType_cxd0, sdfdType_cxd1, cmeOpt, prodSum, parBnds, lnTfmdData=False):
    """Unconstrained optimiser vector -> parsBTfmd (natural scale), in the reference's consumption order."""
    v = list(np.asarray(pars, dtype=float).reshape(-1))
    take = lambda: v.pop(0)

    def exp_(x):      # R's exp(): Inf on overflow (-> parsOK = FALSE -> badnll, fitIAK3D.R:718-741, 885), never an exception
        try:
            return math.exp(x)
        except OverflowError:
            return math.inf
    ev = lambda: 1e-8 + exp_(take())
    if modelx not in ("matern", "wendland", "spherical", "nugget"):
        raise ValueError("unknown modelx")
    nugget = modelx == "nugget"
    ax = nux = float("nan")
    if not nugget:
        ax = float(logitab(take(), *parBnds["axBnds"], invt=True))
        if modelx != "spherical":
            nux = float(logitab(take(), *parBnds["nuxBnds"], invt=True))
    ad = float(logitab(take(), *parBnds["adBnds"], invt=True))
    cx0 = cx1 = 0.0
    if prodSum:
        cx0 = ev()
        if not nugget:
            cx1 = ev()
    cd1 = ev() if sdfdType_cd1 != -9 else 0.0
    if not lnTfmdData:
        if nugget:
            cxd0, cxd1 = 1.0, 0.0
        else:
            cxd1 = float(logitab(take(), *parBnds["sxd1Bnds"], invt=True))
            cxd0 = 1.0 - cxd1
    else:
        cxd0 = ev()
        cxd1 = 0.0 if nugget else ev()
    sd = []
    for t in (sdfdType_cd1, sdfdType_cxd0, sdfdType_cxd1):
        if t == -1:
            sd.append(tauTfm([take(), take()], parBnds, invt=True))
        elif t > 0:
            sd.append(np.array([take() for _ in range(t)]))
        elif t in (0, -9):
            sd.append(np.zeros(0))
        else:
            raise ValueError("sdfd option not yet programmed")
    cme = exp_(take()) if cmeOpt == 1 else 0.0
    out = dict(ax=ax, nux=nux, ad=ad, nud=nud, cx0=cx0, cx1=cx1, cd1=cd1, cxd0=cxd0, cxd1=cxd1,
               sdfdPars_cd1=sd[0], sdfdPars_cxd0=sd[1], sdfdPars_cxd1=sd[2], cme=cme)
    if v:
        out["cssre"] = np.exp(np.array(v))
    return out


# ---------------------------------------------------------------------------------------------
# spline basis of the sd functions (host, once per dataset: splineBasisFns.R:158-325)
# ---------------------------------------------------------------------------------------------
def makeXnsClampedUG0(x, bdryKnots, intKnots, paramVs=0):
    """Natural cubic spline basis with the gradient clamped to 0 beyond the upper boundary knot.  `x` is a vector of
    points or an n x 2 matrix of intervals (then the columns are interval averages).  paramVs = 1 is the
    parameterisation (a, g1..gK-1, w) used for the sd functions (iaCovMatern.R:520-539)."""
    x = np.asarray(x, float)
    ia = x.ndim == 2 and x.shape[1] == 2
    if not ia:
        x = x.reshape(-1)
    n = x.shape[0]
    xL, xU = float(bdryKnots[0]), float(bdryKnots[1])
    knots = np.asarray(intKnots if intKnots is not None else [], float).reshape(-1)
    K = knots.size
    X = np.ones((n, 1 + K))
    if K == 0:
        return X
    if paramVs not in (0, 1):
        raise ValueError("Unknonwn parameterisation option for spline function!")
    pos = lambda v: np.where(v < 0, 0.0, v)
    if ia:
        w = x[:, 1] - x[:, 0]
        E3 = lambda k: 0.25 * (pos(x[:, 1] - k) ** 4 - pos(x[:, 0] - k) ** 4) / w        # E[(x - k)_+^3] over the interval
        lin = 0.5 * (x[:, 0] + x[:, 1])
    else:
        E3 = lambda k: pos(x - k) ** 3
        lin = x
    EL, EU = E3(xL), E3(xU)
    h = lambda k: E3(k) - ((xU - k) / (xU - xL)) * EL + ((xL - k) / (xU - xL)) * EU + 3 * (xU - k) * (k - xL) * lin
    if paramVs == 0:
        for ik in range(K):
            X[:, 1 + ik] = h(knots[ik])
        return X
    kK = knots[K - 1]
    hK = h(kK)
    den = (xU - kK) * (kK - xL) * (kK + xL + xU)
    X[:, 0] = 1 - hK / den
    for ik in range(K - 1):
        k = knots[ik]
        q = ((xU - k) * (k - xL) * (k + xL + xU)) / den
        X[:, 1 + ik] = h(k) - q * hK
    X[:, K] = hK / den
    return X


_SDFD_COMPS = ("cd1", "cxd0", "cxd1")


def _sdfd_spline_bases(dI, sdfdTypes, sdfdKnots):
    """XsdfdSplineU_{cd1,cxd0,cxd1} of iaCovMatern.R:520-539 on the intervals `dI` (None where the type is not > 0)."""
    out = []
    for c, t in zip(_SDFD_COMPS, sdfdTypes):
        if t > 0:
            if sdfdKnots is None:
                raise ValueError("sdfdKnots are required when an sdfdType is > 0")
            out.append(f64(makeXnsClampedUG0(dI, sdfdKnots["bdryKnots"], sdfdKnots["intKnots_" + c], paramVs=1)))
        else:
            out.append(None)
    return out


# ---------------------------------------------------------------------------------------------
# setupMats
# ---------------------------------------------------------------------------------------------
class SetupMats:
    """`setupMats` of the reference as a device-resident dataset handle (iak3d_ds)."""

    def __init__(self, xData, dIData, W=None, ctx=None, sdfdType_cd1=0, sdfdType_cxd0=0, sdfdType_cxd1=0, sdfdKnots=None,
                 siteIDData=None, XData=None, cols4ssre=None):
        self.ctx = ctx or default_context()
        self.xData = f64(np.asarray(xData, float).reshape(len(xData), -1))
        if self.xData.shape[1] != 2:
            raise ValueError("xData must have two coordinate columns")
        self.dIData = f64(np.asarray(dIData, float).reshape(-1, 2))
        self.n = self.xData.shape[0]
        if self.dIData.shape[0] != self.n:
            raise ValueError("xData and dIData must have the same number of rows")
        self._W = None
        h = C.c_void_p()
        Wp, q = None, 0
        if W is not None:
            self._W = f64(W)
            Wp, q = dptr(self._W), self._W.shape[1]
        self.ctx.check(self.ctx.lib.iak3d_dataset_create(self.ctx.h, self.n, dptr(self.xData), dptr(self.dIData), Wp, q, C.byref(h)))
        self.h = h
        self.ctx.adopt(self)
        n_, nx, nd, q_ = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int32()
        self.ctx.check(self.ctx.lib.iak3d_dataset_info(self.h, C.byref(n_), C.byref(nx), C.byref(nd), C.byref(q_)))
        self.nxU, self.ndIU, self.q = nx.value, nd.value, q_.value
        self.maxd = max(float(self.dIData.max()), 2.0)     # iaCovMatern.R:435
        self.sdfdKnots = sdfdKnots
        self.sdfdTypes = (sdfdType_cd1, sdfdType_cxd0, sdfdType_cxd1)
        if max(self.sdfdTypes) > 0:                        # iaCovMatern.R:520-539
            _, did = self.ids()
            first = np.full(self.ndIU, -1, np.int64)
            first[did[::-1]] = np.arange(self.n)[::-1]     # first occurrence of each unique interval
            self.dIU = self.dIData[first]
            self.XsdfdSplineU = _sdfd_spline_bases(self.dIU, self.sdfdTypes, sdfdKnots)
            for c, B in enumerate(self.XsdfdSplineU):
                if B is not None:
                    self.ctx.check(self.ctx.lib.iak3d_dataset_set_sdfd_spline(self.h, c, B.shape[1], dptr(B)))
        self.nssre, self.site_codes, self.cols4ssre = 0, None, None
        if siteIDData is not None:                         # site-specific random effects (iaCovMatern.R:440-460)
            X = np.asarray(XData, float).reshape(self.n, -1)
            if not np.all(X[:, 0] == 1):
                raise ValueError("Error - the lmer-type mixed-effects option (if siteIDData given) only works if first column of X is 1s!")
            self.cols4ssre = [int(c) for c in cols4ssre]
            self.site_codes = {}
            sid = site_ids(siteIDData, self.site_codes)
            Xs = f64(X[:, self.cols4ssre])
            self.ctx.check(self.ctx.lib.iak3d_dataset_set_ssre(self.h, Xs.shape[1], iptr(sid), dptr(Xs)))
            self.nssre = Xs.shape[1]

    def set_W(self, W):
        W = f64(W)
        if W.shape[0] != self.n:
            raise ValueError("W must have n rows")
        if self._W is not None and W.shape == self._W.shape and np.array_equal(W, self._W):
            return
        self.ctx.check(self.ctx.lib.iak3d_dataset_set_W(self.h, dptr(W), W.shape[1]))
        self._W, self.q = W, W.shape[1]

    def upload_W(self, W):
        """Unconditional host->device copy of W (bench.py's end-to-end leg)."""
        W = f64(W)
        self.ctx.check(self.ctx.lib.iak3d_dataset_set_W(self.h, dptr(W), W.shape[1]))
        self._W, self.q = W, W.shape[1]

    def ids(self):
        xid = np.empty(self.n, np.int32); did = np.empty(self.n, np.int32)
        self.ctx.check(self.ctx.lib.iak3d_dataset_ids(self.h, iptr(xid), iptr(did)))
        return xid, did

    def close(self):
        if getattr(self, "h", None) is not None and self.h.value:
            self.ctx.lib.iak3d_dataset_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def site_ids(labels, codes):
    """Site labels (any hashable) -> int32 ids; `codes` (label -> id) is extended in place, so that the data and the map
    supports of one fit share their numbering (equal ids = same site; predictIAK3D.R:15-18 compares the labels as characters)."""
    out = np.empty(len(labels), np.int32)
    for i, lab in enumerate(np.asarray(labels).tolist()):
        key = str(lab)
        if key not in codes:
            codes[key] = len(codes)
        out[i] = codes[key]
    return out


def setupIAK3D(xData, dIData, ctx=None, sdfdType_cd1=0, sdfdType_cxd0=0, sdfdType_cxd1=0, sdfdKnots=None, siteIDData=None, XData=None,
               colnames4ssre=None, **_ignored):
    """iaCovMatern.R:413-548.  colnames4ssre: 0-based columns of XData (the reference's names) of the site-specific random effects."""
    return SetupMats(xData, dIData, ctx=ctx, sdfdType_cd1=sdfdType_cd1, sdfdType_cxd0=sdfdType_cxd0, sdfdType_cxd1=sdfdType_cxd1,
                     sdfdKnots=sdfdKnots, siteIDData=siteIDData, XData=XData, cols4ssre=colnames4ssre)


class SetupMats2:
    """setupIAK3D2: set 1 (rows of the result) stays on the host, set 2 is a resident dataset."""

    def __init__(self, xData, dIData, xData2, dIData2, ctx=None, setup2=None, sdfdType_cd1=0, sdfdType_cxd0=0, sdfdType_cxd1=0,
                 sdfdKnots=None, siteIDData=None, XData=None, siteIDData2=None, XData2=None, cols4ssre=None):
        self.x1 = f64(np.asarray(xData, float).reshape(len(xData), -1))
        self.d1 = f64(np.asarray(dIData, float).reshape(-1, 2))
        types = (sdfdType_cd1, sdfdType_cxd0, sdfdType_cxd1)
        self.set2 = setup2 if setup2 is not None else SetupMats(xData2, dIData2, ctx=ctx, sdfdType_cd1=sdfdType_cd1,
                                                                sdfdType_cxd0=sdfdType_cxd0, sdfdType_cxd1=sdfdType_cxd1,
                                                                sdfdKnots=sdfdKnots, siteIDData=siteIDData2, XData=XData2,
                                                                cols4ssre=cols4ssre)
        self.sid1 = self.Xs1 = None
        if siteIDData is not None:           # iaCovMatern.R:645-663: both sets carry site ids, or neither
            if self.set2.nssre == 0:
                raise ValueError("Error - for setupIAK3D2 define both or neither of siteIDData and siteIDData2!")
            self.sid1 = site_ids(siteIDData, self.set2.site_codes)
            self.Xs1 = f64(np.asarray(XData, float).reshape(self.x1.shape[0], -1)[:, self.set2.cols4ssre])
        if setup2 is not None and sdfdKnots is None:
            sdfdKnots, types = setup2.sdfdKnots, setup2.sdfdTypes
        # spline basis on the rows of set 1 (XsdfdSplineU_* expanded by Kd, iaCovMatern.R:627-645)
        self.Xspl1 = _sdfd_spline_bases(self.d1, types, sdfdKnots) if max(types) > 0 else [None, None, None]


def setupIAK3D2(xData, dIData, xData2, dIData2, ctx=None, sdfdType_cd1=0, sdfdType_cxd0=0, sdfdType_cxd1=0, sdfdKnots=None,
                setup2=None, siteIDData=None, XData=None, siteIDData2=None, XData2=None, colnames4ssre=None, **_ignored):
    """`setup2`: an existing SetupMats of set 2 (then xData2 / dIData2 are not re-uploaded)."""
    return SetupMats2(xData, dIData, xData2, dIData2, ctx=ctx, setup2=setup2, sdfdType_cd1=sdfdType_cd1, sdfdType_cxd0=sdfdType_cxd0,
                      sdfdType_cxd1=sdfdType_cxd1, sdfdKnots=sdfdKnots, siteIDData=siteIDData, XData=XData, siteIDData2=siteIDData2,
                      XData2=XData2, cols4ssre=colnames4ssre)


# ---------------------------------------------------------------------------------------------
# covariance assembly
# ---------------------------------------------------------------------------------------------
def setCIAK3D(parsBTfmd, modelx, sdfdType_cd1, sdfdType_cxd0, sdfdType_cxd1, cmeOpt, setupMats, rtnC=True):
    """-> dict(C, sigma2Vec); C is None (R: NA) for invalid parameters."""
    sm = setupMats
    P = make_pars(parsBTfmd, modelx, sdfdType_cd1, sdfdType_cxd0, sdfdType_cxd1, cmeOpt)
    Cm = np.empty((sm.n, sm.n), order="F") if rtnC else None
    s2 = np.empty(sm.n)
    st = sm.ctx.check(sm.ctx.lib.iak3d_cov_build(sm.h, C.byref(P), dptr(Cm), dptr(s2)), allow=(BAD_PARS,))
    if st == BAD_PARS:
        return dict(C=None, sigma2Vec=None)
    return dict(C=Cm, sigma2Vec=s2)


def factor_build_fetch(parsBTfmd, modelx, sdfdType_cd1, sdfdType_cxd0, sdfdType_cxd1, cmeOpt, setupMats, batched=False, rtnW=False):
    """setCIAK3D as the likelihood's factor build forms it (`cov_factor_kernel` into the blocked factor buffer -- the
    kernel the covariance roofline is quoted on), copied back as a full symmetric matrix.  For the parity tests: the
    reference-facing :func:`setCIAK3D` goes through the term-by-term rectangular kernel instead.  None for invalid
    parameters; with `rtnW` also the bordered W' rows (q x n) as laid out by the kernel."""
    sm = setupMats
    P = make_pars(parsBTfmd, modelx, sdfdType_cd1, sdfdType_cxd0, sdfdType_cxd1, cmeOpt)
    A = np.empty((sm.n, sm.n), order="F")
    Wr = np.empty((sm.q, sm.n), order="F") if (rtnW and sm.q > 0) else None
    st = sm.ctx.check(sm.ctx.lib.iak3d_factor_build_fetch(sm.h, C.byref(P), 1 if batched else 0, dptr(A), dptr(Wr)), allow=(BAD_PARS,))
    if st == BAD_PARS:
        return None
    return (A, Wr) if rtnW else A


def setCIAK3D2(parsBTfmd, modelx, sdfdType_cd1, sdfdType_cxd0, sdfdType_cxd1, cmeOpt, setupMats):
    sm = setupMats
    P = make_pars(parsBTfmd, modelx, sdfdType_cd1, sdfdType_cxd0, sdfdType_cxd1, cmeOpt)
    n1, n2 = sm.x1.shape[0], sm.set2.n
    out = np.empty((n1, n2), order="F")
    ctx = sm.set2.ctx
    if sm.set2.nssre > 0:                                      # site-specific random effects, fitIAK3D.R:1233-1240
        if getattr(sm, "sid1", None) is None:
            raise ValueError("Error - for setupIAK3D2 define both or neither of siteIDData and siteIDData2!")
        st = ctx.check(ctx.lib.iak3d_cov_cross_ssre(sm.set2.h, C.byref(P), n1, dptr(sm.x1), dptr(sm.d1), iptr(sm.sid1), dptr(sm.Xs1),
                                                    dptr(out)), allow=(BAD_PARS,))
    elif max(sdfdType_cd1, sdfdType_cxd0, sdfdType_cxd1) > 0:     # setCApproxIAK3D2, fitIAK3D.R:1121-1124
        st = ctx.check(ctx.lib.iak3d_cov_cross_spl(sm.set2.h, C.byref(P), n1, dptr(sm.x1), dptr(sm.d1), spl_ptrs(sm.Xspl1),
                                                   dptr(out)), allow=(BAD_PARS,))
    else:
        st = ctx.check(ctx.lib.iak3d_cov_cross(sm.set2.h, C.byref(P), n1, dptr(sm.x1), dptr(sm.d1), dptr(out)), allow=(BAD_PARS,))
    return None if st == BAD_PARS else out


def iaCovMatern2(dIData, dIData2, ad, nud, sdfdPars, sdfdType, ctx=None):
    ctx = ctx or default_context()
    d1 = f64(np.asarray(dIData, float).reshape(-1, 2)); d2 = f64(np.asarray(dIData2, float).reshape(-1, 2))
    out = np.empty((d1.shape[0], d2.shape[0]), order="F"); av = np.empty(d1.shape[0])
    sp = f64(np.asarray(sdfdPars, float).reshape(-1)[:2]) if sdfdType == -1 else None
    st = ctx.check(ctx.lib.iak3d_iacov(ctx.h, d1.shape[0], dptr(d1), d2.shape[0], dptr(d2), float(ad), float(nud), int(sdfdType),
                                        dptr(sp), dptr(out), dptr(av)), allow=(BAD_PARS,))
    if st == BAD_PARS:
        raise ValueError("This function can only be applied with nud = 0.5, 1.5 or 2.5 and sdfdType = 0 or -1!")
    return out, av


def spatialCovIAK3D(D, pars, covModel="matern", ctx=None):
    ctx = ctx or default_context()
    Dm = f64(D)
    out = np.empty(Dm.shape, order="F")
    nu = float(pars[2]) if len(pars) > 2 and pars[2] is not None else float("nan")
    st = ctx.check(ctx.lib.iak3d_spatial_cov(ctx.h, Dm.size, dptr(Dm), _lib.MODELX[covModel], float(pars[0]), float(pars[1]), nu,
                                              dptr(out)), allow=(BAD_PARS,))
    return None if st == BAD_PARS else out


# ---------------------------------------------------------------------------------------------
# factorisation
# ---------------------------------------------------------------------------------------------
def lndetANDinvCb(Cm, b=None, ctx=None, methodSolveTEST=1):
    """optim_functions.R:268-312 -> dict(lndetC, invCb); (nan, None) where the reference returns (NA, NA).

    methodSolveTEST mirrors the reference's global of that name (:270; never set by the reference itself, default 1 = Cholesky).
    Branch 0 (:272-289, solve() + determinant()) gives the same lndetC and invCb as branch 1 for every symmetric positive
    definite C -- the only kind the likelihood path forms -- and is served by the same factorisation; where it would differ (an
    indefinite but invertible C, which LU / Bunch-Kaufman can still solve) this implementation returns (NA, NA) like branch 1."""
    if methodSolveTEST not in (0, 1):
        raise ValueError("Error - enter valid methodSolveTEST!")
    ctx = ctx or default_context()
    Cm = f64(Cm)
    n = Cm.shape[0]
    lnd = C.c_double()
    if b is None:
        out = np.empty((n, n), order="F")
        st = ctx.check(ctx.lib.iak3d_lndet_invCb(ctx.h, n, dptr(Cm), 0, None, C.byref(lnd), dptr(out)), allow=(NOT_SPD,))
    else:
        bm = f64(np.asarray(b, float).reshape(n, -1))
        out = np.empty(bm.shape, order="F")
        st = ctx.check(ctx.lib.iak3d_lndet_invCb(ctx.h, n, dptr(Cm), bm.shape[1], dptr(bm), C.byref(lnd), dptr(out)), allow=(NOT_SPD,))
    if st == NOT_SPD:
        return dict(lndetC=float("nan"), invCb=None)
    return dict(lndetC=lnd.value, invCb=out)


# ---------------------------------------------------------------------------------------------
# lognormal data (lnTfmdData = TRUE): setmuIAK3D, calcbetavXbeta, nrUpdatesIAK3DlnN in the Gram domain
# ---------------------------------------------------------------------------------------------
def calcbetavXbeta(vX, beta):
    """fitIAK3D.R:1314-1328: beta' vX_i beta for every observation; vX is (n pU) x pU (blocks stacked)."""
    beta = np.asarray(beta, float).reshape(-1)
    pU = beta.size
    if pU == 0:
        return np.zeros((0, 1))
    vX = np.asarray(vX, float).reshape(-1, pU)
    n = vX.shape[0] // pU
    return ((vX @ beta).reshape(n, pU) @ beta).reshape(n, 1)


def setmuIAK3D(XData, vXU, iU, beta, diagC, sigma2Vec):
    """fitIAK3D.R:1261-1267."""
    beta = np.asarray(beta, float).reshape(-1, 1)
    n = XData.shape[0]
    bvb = calcbetavXbeta(vXU, beta[list(iU), 0]) if len(iU) else np.zeros((n, 1))
    return XData @ beta + 0.5 * (np.asarray(sigma2Vec, float).reshape(n, 1) - np.asarray(diagC, float).reshape(n, 1) + bvb)


def lnN_design(XData, zData, vXU, iU):
    """F = [X, z, s, V_ab (a <= b over iU)]: every vector nrUpdatesIAK3DlnN ever multiplies by iC or TK is a linear
    combination of these columns -- resU = z - XU betaU - (s + sum w_ab V_ab) / 2 with s = sigma2Vec - diag(C), and
    XU + dbetavXbeta (nrUpdatesIAK3DlnN.R:188-207) -- so ONE bordered factorisation with W = F returns everything the
    Newton iteration needs as the q x q Gram matrix F' iC F (no explicit iC or TK, which the reference forms, :22,63).
    The s column depends on the covariance parameters: it is filled on the device (iak3d_dataset_set_lnN_column)."""
    n, p = XData.shape
    pU = len(iU)
    cols = [XData, np.asarray(zData, float).reshape(n, 1), np.zeros((n, 1))]
    pairs = [(a, b) for a in range(pU) for b in range(a, pU)]
    if pU:
        v = np.asarray(vXU, float).reshape(n, pU, pU)
        if not np.allclose(v, np.transpose(v, (0, 2, 1)), rtol=1e-12, atol=0):
            raise ValueError("vXU: the within-support covariance blocks must be symmetric")
        cols += [v[:, a, b].reshape(n, 1) for (a, b) in pairs]
    F = np.column_stack(cols)
    if F.shape[1] > 64:
        raise ValueError("lognormal data: p + 2 + pU (pU + 1) / 2 must not exceed 64")
    return F, pairs


def nrUpdatesIAK3DlnN_gram(G, lndetC, n, p, iU, pairs):
    """nrUpdatesIAK3DlnN + gradHessIAK3DlnN (nrUpdatesIAK3DlnN.R:1-226) on G = F' iC F (see :func:`lnN_design`).
    u' iC v = cu' G cv and u' TK v = cu' GT cv with GT = G - G[:,K] G[K,K]^-1 G[K,:] for vectors u = F cu, v = F cv.
    The Hessian branch of the reference (`checkHess$cholC` on the result of solve(), :109) cannot run as written; the
    evident intent is implemented: Newton step from the third update on, Fisher scoring if the Hessian solve fails."""
    G = 0.5 * (G + G.T)                                                # :31
    q = G.shape[0]
    iU = list(iU); pU = len(iU)
    iK = [j for j in range(p) if j not in iU]
    zc, sc = p, p + 1
    vcol = {pr: p + 2 + k for k, pr in enumerate(pairs)}
    vcol.update({(b, a): c for (a, b), c in list(vcol.items())})
    e = lambda j: np.eye(q)[:, j]
    XiCX = G[:p, :p]
    if pU == 0:                                                        # :37-50
        c = e(zc) - 0.5 * e(sc)
        XiCz = G[:p, :] @ c
        beta = np.linalg.solve(XiCX, XiCz).reshape(-1, 1)
        res = float(c @ G @ c - 2.0 * (beta[:, 0] @ XiCz) + beta[:, 0] @ XiCX @ beta[:, 0])
        return dict(betahat=beta, nll=0.5 * (n * math.log(2.0 * math.pi) + lndetC + res), grad=None, hess=None, fim=-XiCX,
                    noits=0, errFlag=0)
    betaU = np.linalg.solve(XiCX, G[:p, zc])[iU]                       # :60
    if iK:
        GKK = G[np.ix_(iK, iK)]
        GT = G - G[:, iK] @ np.linalg.solve(GKK, G[iK, :])             # TK, :62-64
    else:
        GKK, GT = None, G

    def gh(bU, GTm=GT, K=iK, U=iU, lnd=lndetC):
        cres = e(zc) - 0.5 * e(sc)                                     # resU, :194
        ca = []
        for a, ja in enumerate(U):
            cres = cres - bU[a] * e(ja)
            c = e(ja)
            for b in range(len(U)):
                c = c + bU[b] * e(vcol[(a, b)])                        # XU + dbetavXbeta, :204-207
                cres = cres - 0.5 * bU[a] * bU[b] * e(vcol[(a, b)])    # betavXbeta / 2
            ca.append(c)
        Ca = np.column_stack(ca)
        Tres = GTm @ cres
        nll = 0.5 * (n * math.log(2.0 * math.pi) + lnd + float(cres @ Tres))      # :213
        grad = -(Ca.T @ Tres)                                          # :215
        fim = Ca.T @ GTm @ Ca                                          # :222
        vT = np.array([[Tres[vcol[(a, b)]] for b in range(len(U))] for a in range(len(U))])
        hess = -vT + fim                                               # :217-220
        betahat = np.full((p, 1), np.nan)
        betahat[U, 0] = bU
        if K:
            betahat[K, 0] = np.linalg.solve(GKK, G[K, :] @ cres)      # :199
        return nll, grad, 0.5 * (hess + hess.T), 0.5 * (fim + fim.T), betahat

    nllNew, grad, hess, fim, betahat = gh(betaU)
    tol, noits, nllPrev = 1e-6, 0, math.inf
    while noits < 20 and (nllNew < nllPrev - tol or noits < 4):        # :94
        nllPrev = nllNew
        step = None
        if noits >= 2:
            try:
                step = np.linalg.solve(hess, grad)
            except np.linalg.LinAlgError:
                step = None
        if step is None:
            try:
                step = np.linalg.solve(fim, grad)
            except np.linalg.LinAlgError:
                return dict(betahat=betahat, nll=nllNew, grad=grad, hess=hess, fim=fim, noits=noits, errFlag=1)
        betaU = betaU - step
        nllNew, grad, hess, fim, betahat = gh(betaU)
        noits += 1
    rnd = tol * 10                                                     # :160-161
    return dict(betahat=betahat, nll=float(np.round(nllNew / rnd) * rnd), nll_unrounded=nllNew, grad=grad, hess=hess, fim=fim, noits=noits,
                errFlag=2 if noits == 20 else 0, gh=gh)


def _nll_lnN(pars, zData, XData, vXU, iU, sm, P, parsBTfmd, modelx, types, cmeOpt, rtnAll, attachBigMats):
    """lnTfmdData branch of nllIAK3D (fitIAK3D.R:762-781, 919-945).  iU is 0-based."""
    n, p = XData.shape
    iU = [] if iU is None else [int(j) for j in iU]
    F, pairs = lnN_design(XData, zData, vXU, iU)
    ctx = sm.ctx
    sm.set_W(F)
    ctx.check(ctx.lib.iak3d_dataset_set_lnN_column(sm.h, p + 1))
    q = F.shape[1]
    lnd = C.c_double(); G = np.empty((q, q), order="F")
    try:
        st = ctx.check(ctx.lib.iak3d_nll_terms(sm.h, C.byref(P), C.byref(lnd), dptr(G), None), allow=(NOT_SPD, BAD_PARS))
    finally:
        ctx.check(ctx.lib.iak3d_dataset_set_lnN_column(sm.h, -1))
    bad = dict(nll=BADNLL, pars=pars, parsBTfmd=parsBTfmd, betahat=None, vbetahat=None, cxdhat=float("nan")) if rtnAll else BADNLL
    if st != OK:
        return bad
    nr = nrUpdatesIAK3DlnN_gram(np.array(G), lnd.value, n, p, iU, pairs)
    nll = nr["nll"] if np.isfinite(nr["nll"]) else BADNLL
    if not rtnAll:
        return nll
    betahat = nr["betahat"]
    # vbetahat = solve(fim) with every column treated as varying and TK = iC (fitIAK3D.R:768-773)
    Gs = 0.5 * (np.array(G) + np.array(G).T)
    ca = []
    vcol = {pr: p + 2 + k for k, pr in enumerate(pairs)}
    vcol.update({(b, a): c for (a, b), c in list(vcol.items())})
    for j in range(p):
        c = np.eye(q)[:, j]
        if j in iU:
            a = iU.index(j)
            for b, jb in enumerate(iU):
                c = c + betahat[jb, 0] * np.eye(q)[:, vcol[(a, b)]]
        ca.append(c)
    Ca = np.column_stack(ca)
    vbetahat = np.linalg.inv(Ca.T @ Gs @ Ca)
    s2 = np.empty(n); dC = np.empty(n)
    ctx.check(ctx.lib.iak3d_cov_build(sm.h, C.byref(P), None, dptr(s2)))
    ctx.check(ctx.lib.iak3d_cov_diag(sm.h, C.byref(P), dptr(dC)))
    muhat = setmuIAK3D(XData, vXU, iU, betahat, dC, s2)                # :921
    out = dict(nll=nll, pars=pars, parsBTfmd=parsBTfmd, betahat=betahat, vbetahat=vbetahat, cxdhat=1.0, sigma2Vec=s2, XData=XData,
               zData=zData, vXU=vXU, iU=iU, lndetA=lnd.value, modelx=modelx, nud=parsBTfmd["nud"], sdfdType_cd1=types[0],
               sdfdType_cxd0=types[1], sdfdType_cxd1=types[2], cmeOpt=cmeOpt, setupMats=sm, xData=sm.xData, dIData=sm.dIData,
               lnTfmdData=True, diagC=dC, noits=nr["noits"], errFlag=nr["errFlag"], nll_unrounded=nr.get("nll_unrounded", nll))
    sm.set_W(np.column_stack([XData, zData]))
    if attachBigMats:
        fit = KeptFit(sm, parsBTfmd, modelx, types, cmeOpt, betahat, vbetahat, 1.0, z_muhat=zData.reshape(-1, 1) - muhat)
        iCX, iCz = fit.get_small()
        out.update(fit=fit, iCX=iCX, iCz_muhat=iCz)
    return out


# ---------------------------------------------------------------------------------------------
# likelihood
# ---------------------------------------------------------------------------------------------
def _small_chol_solve(A, b):
    """lndetANDinvCb on the p x p system (host: O(p^3), as in the reference-side R glue)."""
    try:
        L = np.linalg.cholesky(A)
    except np.linalg.LinAlgError:
        return float("nan"), None
    d = np.diag(L)
    if not np.all(np.isfinite(d)) or np.any(d <= 0):
        return float("nan"), None
    y = np.linalg.solve(L, b)
    return 2.0 * float(np.sum(np.log(d))), np.linalg.solve(L.T, y)


def reml_tail(WiAW, lndetA, n, p, useReml=True, n_for_cxd=None, n_for_lndet=None):
    """fitIAK3D.R:821-880 / compositeLikIAK3D.R:191-265.  -> dict(nll, betahat, cxdhat, lndetXiAX, resiAres).  One element
    of :func:`reml_tail_batch`, so that a single evaluation and the same evaluation inside a batch agree bit for bit."""
    t = reml_tail_batch(np.asarray(WiAW, float)[None], np.array([lndetA], float), n, p, useReml, n_for_cxd, n_for_lndet, full=True)
    if not t["ok"][0]:
        return dict(nll=BADNLL, betahat=None, cxdhat=float("nan"))
    return dict(nll=float(t["nll"][0]), betahat=t["betahat"][0], cxdhat=float(t["cxdhat"][0]), lndetXiAX=float(t["lndetXiAX"][0]),
                resiAres=float(t["resiAres"][0]))


def reml_tail_batch(WiAW, lndetA, n, p, useReml=True, n_for_cxd=None, n_for_lndet=None, full=False):
    """The closed-form REML / ML tail for a stack of evaluations (count x q x q, count): one call of the library's host routine
    per evaluation (`iak3d_reml_tail`, csrc/tail.cu -- O(p^3) C++, a few microseconds; the NumPy version of round 1 cost
    0.18 ms per evaluation, a fifth of an Edgeroi-size evaluation).  A single evaluation and the same evaluation inside a
    batch go through the same arithmetic, so they agree bit for bit.  -> nll (count,), BADNLL where the p x p system is not
    positive definite; with `full` a dict of arrays (nll, ok, betahat, cxdhat, lndetXiAX, resiAres)."""
    lib = load()
    WiAW = np.ascontiguousarray(WiAW, dtype=np.float64); lndetA = np.asarray(lndetA, float).reshape(-1)
    cnt = WiAW.shape[0]
    q = p + 1
    nll = np.empty(cnt)
    beta = np.full((cnt, p, 1), np.nan); cxd = np.empty(cnt); ldx = np.empty(cnt); res = np.empty(cnt); ok = np.zeros(cnt, np.int32)
    nc = -1.0 if n_for_cxd is None else float(n_for_cxd)
    nl = -1.0 if n_for_lndet is None else float(n_for_lndet)
    # the routine reads the upper triangle of a column-major matrix: hand it the transposed copy of the C-ordered stack
    Wt = np.ascontiguousarray(np.transpose(WiAW, (0, 2, 1)))
    for k in range(cnt):
        st = lib.iak3d_reml_tail(dptr(Wt[k]), float(lndetA[k]), int(n), int(p), 1 if useReml else 0, nc, nl, dptr(nll[k:]), dptr(beta[k]),
                                 dptr(cxd[k:]), dptr(ldx[k:]), dptr(res[k:]), iptr(ok[k:]))
        if st != 0:
            raise Iak3dError("reml_tail: bad arguments")
    if not full:
        return nll
    return dict(nll=nll, ok=ok.astype(bool), betahat=beta, cxdhat=cxd, lndetXiAX=ldx, resiAres=res)


def nll_terms_batch(ctx, setups, parsList):
    """One launch for many (setupMats, struct pars) pairs -> lndetA[count], WiAW[count,q,q], status[count]."""
    count = len(setups)
    q = setups[0].q
    dsarr = (C.c_void_p * count)(*[s.h for s in setups])
    parr = (Pars * count)(*parsList)
    lnd = np.empty(count); W = np.empty((count, q * q)); st = np.empty(count, np.int32)
    ctx.check(ctx.lib.iak3d_nll_terms_batch(ctx.h, count, dsarr, parr, dptr(lnd), dptr(W), iptr(st)))
    # G is stored column-major q x q and is symmetric
    return lnd, W.reshape(count, q, q), st


def _symm_upper(M):
    return np.triu(M) + np.triu(M, 1).T     # forceSymmetric keeps the upper triangle (fitIAK3D.R:811)


def nllIAK3D(pars, zData, XData, vXU=None, iU=None, modelx="matern", nud=0.5, sdfdType_cd1=0, sdfdType_cxd0=0, sdfdType_cxd1=0,
             cmeOpt=0, prodSum=True, setupMats=None, parBnds=None, useReml=True, lnTfmdData=False, rtnAll=False,
             forCompLik=False, attachBigMats=False, nCores=8, parsBTfmd=None):
    """fitIAK3D.R:670-949.  `parsBTfmd` may be passed to skip the transform.  With lnTfmdData = TRUE (no profiling of
    the variance, no REML: :147-155) `iU` holds the 0-based columns of XData that vary inside the sample supports and
    `vXU` their (n pU) x pU within-support covariances."""
    if setupMats is None:
        raise ValueError("setupMats (from setupIAK3D) is required")
    sm = setupMats
    zData = np.asarray(zData, float).reshape(-1)
    XData = np.asarray(XData, float).reshape(zData.size, -1)
    n, p = XData.shape
    if parsBTfmd is None and not (lnTfmdData or rtnAll or forCompLik):
        # what an optimiser calls: the vector goes to the library as it is (iak3d_nll_vector: readParsIAK3D, device terms and
        # the closed-form tail natively -- one call, no interpreter work per evaluation beyond this function)
        sm.set_W(np.column_stack([XData, zData]))
        M = make_model(modelx, nud, sdfdType_cd1, sdfdType_cxd0, sdfdType_cxd1, cmeOpt, prodSum, parBnds)
        pv = np.ascontiguousarray(pars, dtype=np.float64).reshape(-1)
        v = C.c_double()
        sm.ctx.check(sm.ctx.lib.iak3d_nll_vector(sm.h, C.byref(M), dptr(pv), pv.size, 1 if useReml else 0, C.byref(v)))
        return v.value
    if parsBTfmd is None:
        parsBTfmd = readParsIAK3D(pars, modelx, nud, sdfdType_cd1, sdfdType_cxd0, sdfdType_cxd1, cmeOpt, prodSum, parBnds,
                                  lnTfmdData=lnTfmdData)
    parsBTfmd = dict(parsBTfmd)
    if lnTfmdData:
        if forCompLik:
            raise ValueError("composite likelihood is not defined for lognormal data (fitIAK3D.R:135-140)")
        P = make_pars(parsBTfmd, modelx, sdfdType_cd1, sdfdType_cxd0, sdfdType_cxd1, cmeOpt)
        return _nll_lnN(pars, zData, XData, vXU, iU, sm, P, parsBTfmd, modelx, (sdfdType_cd1, sdfdType_cxd0, sdfdType_cxd1), cmeOpt,
                        rtnAll, attachBigMats)
    sm.set_W(np.column_stack([XData, zData]))                      # W <- cbind(XData, zData), :757
    P = make_pars(parsBTfmd, modelx, sdfdType_cd1, sdfdType_cxd0, sdfdType_cxd1, cmeOpt)
    ctx = sm.ctx
    if not (rtnAll or forCompLik):                                  # what an optimiser calls: one library call, device + host tail
        v, stt = C.c_double(), C.c_int32()
        ctx.check(ctx.lib.iak3d_nll_reml(sm.h, C.byref(P), 1 if useReml else 0, C.byref(v), None, None, None, None, C.byref(stt)))
        return BADNLL if stt.value != OK else v.value
    lnd = C.c_double(); G = np.empty((p + 1, p + 1), order="F")
    want_iAW = forCompLik or (rtnAll and not attachBigMats)
    iAW = np.empty((n, p + 1), order="F") if want_iAW else None
    st = ctx.check(ctx.lib.iak3d_nll_terms(sm.h, C.byref(P), C.byref(lnd), dptr(G), dptr(iAW)), allow=(NOT_SPD, BAD_PARS))
    bad = dict(nll=BADNLL, pars=pars, parsBTfmd=parsBTfmd, betahat=None, vbetahat=None, cxdhat=float("nan")) if rtnAll else BADNLL
    if st != OK:
        if forCompLik:
            return dict(pars=pars, parsBTfmd=parsBTfmd, WiAW=np.full((p + 1, p + 1), np.nan), lndetA=float("nan"))
        return bad
    WiAW = _symm_upper(np.array(G))
    lndetA = lnd.value
    if forCompLik:
        out = dict(pars=pars, parsBTfmd=parsBTfmd, WiAW=WiAW, lndetA=lndetA, iAW=iAW)
        out["sigma2Vec"] = setCIAK3D(parsBTfmd, modelx, sdfdType_cd1, sdfdType_cxd0, sdfdType_cxd1, cmeOpt, sm, rtnC=attachBigMats)
        if attachBigMats:
            out["A"] = out["sigma2Vec"]["C"]
        out["sigma2Vec"] = out["sigma2Vec"]["sigma2Vec"]
        return out
    t = reml_tail(WiAW, lndetA, n, p, useReml)
    if t["nll"] == BADNLL and t["betahat"] is None:
        return bad
    if not rtnAll:
        return t["nll"]
    cxdhat, betahat = t["cxdhat"], t["betahat"]
    unscaled = dict(parsBTfmd)
    for k in ("cx0", "cx1", "cd1", "cxd0", "cxd1"):                # :857-861
        parsBTfmd[k] = cxdhat * parsBTfmd[k]
    if cmeOpt == 1:
        parsBTfmd["cme"] = cxdhat * parsBTfmd["cme"]
    if parsBTfmd.get("cssre") is not None:                         # :863
        parsBTfmd["cssre"] = cxdhat * np.asarray(parsBTfmd["cssre"], float)
    vbetahat = cxdhat * np.linalg.inv(WiAW[:p, :p])                # :870
    s2 = setCIAK3D(unscaled, modelx, sdfdType_cd1, sdfdType_cxd0, sdfdType_cxd1, cmeOpt, sm, rtnC=False)["sigma2Vec"]
    out = dict(nll=t["nll"], pars=pars, parsBTfmd=parsBTfmd, betahat=betahat, vbetahat=vbetahat, cxdhat=cxdhat,
               sigma2Vec=cxdhat * s2, XData=XData, zData=zData, vXU=vXU, iU=iU, lndetA=lndetA, WiAW=WiAW,
               modelx=modelx, nud=nud, sdfdType_cd1=sdfdType_cd1, sdfdType_cxd0=sdfdType_cxd0, sdfdType_cxd1=sdfdType_cxd1,
               cmeOpt=cmeOpt, setupMats=sm, xData=sm.xData, dIData=sm.dIData)
    if attachBigMats:
        # the big matrices stay on the device as a kept factorisation (iak3d_fit); iCX / iCz_muhat are copied out
        fit = KeptFit(sm, parsBTfmd, modelx, (sdfdType_cd1, sdfdType_cxd0, sdfdType_cxd1), cmeOpt, betahat, vbetahat, cxdhat)
        iCX, iCz = fit.get_small()
        out.update(fit=fit, iCX=iCX, iCz_muhat=iCz)
    else:
        out.update(iAX=iAW[:, :p], iAz=iAW[:, p:p + 1])
    return out


def gradnllIAK3D(pars, zData, XData, vXU=None, iU=None, modelx="matern", nud=0.5, sdfdType_cd1=0, sdfdType_cxd0=0,
                 sdfdType_cxd1=0, cmeOpt=0, prodSum=True, setupMats=None, parBnds=None, useReml=True, lnTfmdData=False,
                 delta=1e-6, rtn_nll=False, analytic=False):
    """Forward-difference gradient (fitIAK3D.R:1865-1886): the npar + 1 evaluations go to the GPU as ONE batch
    (the reference forks them with mclapply, :1804).  NA differences become 0 (:1813).

    `analytic=True` (not in the reference): the same candidates, but ONE factorisation, the explicit
    inverse and one pass over it -- d nll / d theta_k = 1/2 sum_ij w_ij (A_k - A_0)_ij / delta (`iak3d_nll_grad`); the
    differences are taken on the covariance ENTRIES, not on the likelihood, so no factorisation noise enters.
    `analytic="auto"` picks it from n = 4000 on, where it is faster than the batch."""
    if lnTfmdData:
        # nothing is profiled in lognormal mode (fitIAK3D.R:147-155): every candidate is a full nllIAK3D with its own
        # Newton-Raphson for beta -- one bordered factorisation each.  The values are the reference's (rounded to 1e-5,
        # :160-161), and so is the difference quotient.
        pars = np.asarray(pars, float).reshape(-1)
        vals = np.empty(pars.size + 1)
        for k in range(pars.size + 1):
            v = pars.copy()
            if k > 0:
                v[k - 1] += delta
            vals[k] = nllIAK3D(v, zData, XData, vXU=vXU, iU=iU, modelx=modelx, nud=nud, sdfdType_cd1=sdfdType_cd1,
                               sdfdType_cxd0=sdfdType_cxd0, sdfdType_cxd1=sdfdType_cxd1, cmeOpt=cmeOpt, prodSum=prodSum,
                               setupMats=setupMats, parBnds=parBnds, useReml=useReml, lnTfmdData=True)
        g = (vals[1:] - vals[0]) / delta
        g[~np.isfinite(g)] = 0.0
        return (g, vals[0]) if rtn_nll else g
    sm = setupMats
    pars = np.asarray(pars, float).reshape(-1)
    zData = np.asarray(zData, float).reshape(-1)
    XData = np.asarray(XData, float).reshape(zData.size, -1)
    n, p = XData.shape
    sm.set_W(np.column_stack([XData, zData]))
    npar = pars.size
    if analytic == "auto":            # measured on B200: the pass pays from n ~ 4000 on (1.7x at 5000, 2.8x at 20000)
        analytic = n >= 4000 and getattr(sm, "nssre", 0) == 0      # site-specific random effects: forward differences
    if not analytic:
        # forward differences: npar + 1 evaluations in one batched launch, transform and tails in the library (iak3d_gradnll_fd)
        M = make_model(modelx, nud, sdfdType_cd1, sdfdType_cxd0, sdfdType_cxd1, cmeOpt, prodSum, parBnds)
        g = np.empty(npar); v0 = C.c_double()
        sm.ctx.check(sm.ctx.lib.iak3d_gradnll_fd(sm.h, C.byref(M), dptr(pars), npar, float(delta), 1 if useReml else 0, dptr(g), C.byref(v0)))
        return (g, v0.value) if rtn_nll else g
    cand = [pars.copy()]
    for i in range(npar):
        v = pars.copy(); v[i] += delta
        cand.append(v)
    structs = []
    for v in cand:
        pb = readParsIAK3D(v, modelx, nud, sdfdType_cd1, sdfdType_cxd0, sdfdType_cxd1, cmeOpt, prodSum, parBnds)
        structs.append(make_pars(pb, modelx, sdfdType_cd1, sdfdType_cxd0, sdfdType_cxd1, cmeOpt))
    if analytic:
        parr = (Pars * len(cand))(*structs)
        dl = f64(np.full(npar, float(delta)))
        lnd1 = C.c_double(); G1 = np.empty((p + 1, p + 1), order="F"); gs = np.empty(npar)
        st1 = sm.ctx.check(sm.ctx.lib.iak3d_nll_grad(sm.h, npar, parr, dptr(dl), 1 if useReml else 0, C.byref(lnd1), dptr(G1), dptr(gs)),
                           allow=(NOT_SPD, BAD_PARS))
        if st1 == OK:
            g = 0.5 * gs
            g[~np.isfinite(g)] = 0.0
            if not rtn_nll:
                return g
            return g, reml_tail(_symm_upper(np.array(G1)), lnd1.value, n, p, useReml)["nll"]
        # the base point failed (not positive definite / invalid): fall through to the batch, whose failure behaviour is the
        # reference's -- (badnll - nll0) / delta per failed candidate, 0 where the difference is NA (fitIAK3D.R:1812-1813)
    lnd, G, st = nll_terms_batch(sm.ctx, [sm] * len(cand), structs)
    Gt = np.transpose(G, (0, 2, 1))
    Gs = np.triu(Gt) + np.transpose(np.triu(Gt, 1), (0, 2, 1))           # forceSymmetric keeps the upper triangle (fitIAK3D.R:811)
    Gs[st != OK] = np.nan
    vals = reml_tail_batch(Gs, np.where(st == OK, lnd, np.nan), n, p, useReml)
    g = (vals[1:] - vals[0]) / delta
    g[~np.isfinite(g)] = 0.0
    return (g, vals[0]) if rtn_nll else g


def cl_terms_sets(units, structs, q, comm=None, timings=None):
    """The device side of composite-likelihood evaluations for `len(structs)` parameter sets at once
    (compositeLikIAK3D.R:49-93, 105-155 for each set): every (unit of this rank, parameter set) pair is one problem of a
    batched launch, the per-unit terms are summed on the device per parameter set (`iak3d_nll_terms_batch_wsum`), and ONE
    all-reduce of nsets x (q^2 + 2) doubles combines the ranks (compositeLikIAK3D.R:63-64, Reduce('+')).  The sets are sent
    in as many launches as device memory allows (a factor buffer per problem in flight).
    -> array (nsets, q*q + 2): [sum w forceSymmetric(WiAW) (column-major), sum w lndetA, number of failed units].

    A rank whose device call raises still takes part in the all-reduce (with NaN), so that the others do not hang in
    NCCL; every rank raises afterwards."""
    nsets, qq = len(structs), q * q
    width = qq + 2
    dev = None
    if comm is not None and comm.world > 1:
        comm.attach_peer(units[0]["setup"].ctx if units else None)    # first call only: map the peers' mailboxes (collective)
        dev = comm.device_buffer(nsets * width)                       # torch CUDA tensor, or None (gloo without peers: host sums)
    host = np.zeros((nsets, width))
    err = None
    t_dev0 = None
    ctxs = []
    for u in units:
        if not any(u["setup"].ctx is c for c in ctxs):
            ctxs.append(u["setup"].ctx)
    if len(ctxs) > 1:
        # several GPUs driven from THIS process (what an R session can do: no fork, no NCCL): one host thread per device inside
        # the library (iak3d_nll_terms_batch_wsum_multi), per-device sums added on the host in device order
        if comm is not None and comm.world > 1:
            raise ValueError("units on several contexts cannot be combined with a multi-process communicator")
        order = [u for c in ctxs for u in units if u["setup"].ctx is c]
        counts = np.array([sum(1 for u in units if u["setup"].ctx is c) for c in ctxs], np.int32)
        carr = (C.c_void_p * len(ctxs))(*[c.h for c in ctxs])
        dsarr = (C.c_void_p * len(order))(*[u["setup"].h for u in order])
        w = f64(np.array([u["weight"] for u in order], float))
        grp = np.zeros(len(order), np.int32)
        for k in range(nsets):                     # one parameter set per call: a factor buffer per unit in flight
            parr = (Pars * len(order))(*[structs[k]] * len(order))
            ctxs[0].check(ctxs[0].lib.iak3d_nll_terms_batch_wsum_multi(len(ctxs), carr, iptr(counts), dsarr, parr, dptr(w), iptr(grp), 1,
                                                                      q, dptr(host[k:k + 1])))
        return host
    if units:
        ctx = units[0]["setup"].ctx
        try:
            g = 1
            if nsets > 1:                        # parameter sets per launch: a factor buffer per problem in flight must fit
                need = 0
                for u in units:
                    if "slot_bytes" not in u:
                        sb = C.c_int64()
                        ctx.check(ctx.lib.iak3d_dataset_slot_bytes(u["setup"].h, C.byref(sb), None))
                        u["slot_bytes"] = sb.value
                    need += u["slot_bytes"]
                g = int(max(1, min(nsets, (0.7 * ctx.device_info()["total_mem"]) // max(need, 1))))
            wts = f64(np.array([u["weight"] for u in units], float))
            for s0 in range(0, nsets, g):
                s1 = min(nsets, s0 + g)
                cnt = len(units) * (s1 - s0)
                dsarr = (C.c_void_p * cnt)(*[u["setup"].h for _ in range(s0, s1) for u in units])
                parr = (Pars * cnt)(*[structs[k] for k in range(s0, s1) for _ in units])
                w = f64(np.tile(wts, s1 - s0))
                grp = np.repeat(np.arange(s1 - s0, dtype=np.int32), len(units))
                dptr_out = None if dev is None else C.c_void_p(dev.data_ptr() + s0 * width * 8)
                hptr = dptr(host[s0:s1]) if dev is None else None
                ctx.check(ctx.lib.iak3d_nll_terms_batch_wsum(ctx.h, cnt, dsarr, parr, dptr(w), iptr(grp), s1 - s0, dptr_out, hptr))
        except Exception as e:                      # noqa: BLE001 -- re-raised below, after the collective
            err = e
    if comm is None or comm.world == 1:
        if err is not None:
            raise err
        return host
    if dev is None:                                 # gloo (CPU tests of the host logic): host buffer through the same exchange
        if err is not None:
            host[:] = np.nan
        out = comm.allreduce_sum(host.reshape(-1)).reshape(nsets, width)
    else:
        if not units:
            dev[:nsets * width].zero_()
        if err is not None:
            dev[:nsets * width].fill_(float("nan"))
        out = comm.allreduce_device(nsets * width, timings).reshape(nsets, width)
    if err is not None:
        raise err
    if np.any(np.isnan(out)):
        raise Iak3dError("composite likelihood: another rank failed during the evaluation (see its log)")
    return out


def _cl_n_sum(compLikMats):
    """n_SUM of compositeLikIAK3D.R:64,83: the observations of all adjacent pairs -- known to every rank from the block
    lists, so it needs no exchange."""
    if compLikMats["compLikOptn"] >= 4:
        return 0.0
    blocks = compLikMats["listBlocks"]
    return float(sum(len(blocks[k]) + len(blocks[l]) for (k, l) in np.asarray(compLikMats["subsetPairsAdj"]).reshape(-1, 2)))


def nllIAK3D_CL(pars, zData, XData, modelx, nud, sdfdType_cd1, sdfdType_cxd0, sdfdType_cxd1, cmeOpt, prodSum, parBnds, useReml,
                compLikMats, parsBTfmd=None, rtnTerms=False, comm=None, rtnAll=False, timings=None):
    """compositeLikIAK3D.R:6-307.  `compLikMats` carries compLikOptn, listBlocks, subsetPairsAdj, subsetPairsNonadj and
    the per-unit device datasets made by :func:`setupIAK3D_CL` (first the adjacent pairs, then the single blocks --
    the order of iaCovMatern.R:551-575).  With `comm` (a parallel.BlockComm) the units are sharded across ranks and
    the sums are combined by one all-reduce of (p+1)^2 + 2 doubles."""
    zData = np.asarray(zData, float).reshape(-1)
    XData = np.asarray(XData, float).reshape(zData.size, -1)
    n, p = XData.shape
    optn = compLikMats["compLikOptn"]
    nB = len(compLikMats["listBlocks"])
    if parsBTfmd is None:
        parsBTfmd = readParsIAK3D(pars, modelx, nud, sdfdType_cd1, sdfdType_cxd0, sdfdType_cxd1, cmeOpt, prodSum, parBnds)
    P = make_pars(parsBTfmd, modelx, sdfdType_cd1, sdfdType_cxd0, sdfdType_cxd1, cmeOpt)
    t = cl_terms_sets(compLikMats["units"], [P], p + 1, comm, timings)[0]
    S = t[:(p + 1) ** 2].reshape(p + 1, p + 1); lnd_sum, failed = t[-2], t[-1]
    n_SUM = _cl_n_sum(compLikMats)
    if not rtnAll:
        return cl_reduce_and_tail(S, lnd_sum, n_SUM, failed, n, p, optn, nB, useReml, None, rtnTerms)
    # rtnAll (compositeLikIAK3D.R:231-256, 295-303): the fitted list predictIAK3D_CL works from
    t = cl_reduce_and_tail(S, lnd_sum, n_SUM, failed, n, p, optn, nB, useReml, None, True)
    if not isinstance(t, dict):
        return dict(nll=BADNLL, pars=pars, parsBTfmd=parsBTfmd)
    cxdhat = t["cxdhat"]
    fitted = dict(parsBTfmd)
    for k in ("cx0", "cx1", "cd1", "cxd0", "cxd1"):
        fitted[k] = cxdhat * fitted[k]
    if cmeOpt == 1:
        fitted["cme"] = cxdhat * fitted["cme"]
    A = t["XziAXz_SUM"] / (nB - 1) if optn == 1 else t["XziAXz_SUM"]
    if optn in (2, 3):
        A = n * A / t["n_SUM"]                                          # XziAXz_APPROX, :238
    vbetahat = cxdhat * np.linalg.inv(A[:p, :p])
    return dict(nll=t["nll"], pars=pars, parsBTfmd=fitted, betahat=t["betahat"], vbetahat=vbetahat, cxdhat=cxdhat, XData=XData,
                zData=zData, XziAXz_APPROX=A, XziAXz_SUM=t["XziAXz_SUM"], n_SUM=t["n_SUM"], compLikMats=compLikMats, modelx=modelx,
                nud=nud, sdfdType_cd1=sdfdType_cd1, sdfdType_cxd0=sdfdType_cxd0, sdfdType_cxd1=sdfdType_cxd1, cmeOpt=cmeOpt,
                lnTfmdData=False, sdfdKnots=compLikMats.get("sdfdKnots"))


def gradnllIAK3D_CL(pars, zData, XData, modelx, nud, sdfdType_cd1, sdfdType_cxd0, sdfdType_cxd1, cmeOpt, prodSum, parBnds, useReml,
                    compLikMats, comm=None, delta=1e-6, rtn_nll=False, timings=None):
    """Forward-difference gradient of the composite likelihood (gradnllIAK3D_CL[.mc], fitIAK3D.R:1819-1911): the npar + 1
    composite-likelihood evaluations are (units x parameter sets) problems of batched launches and ONE all-reduce of
    (npar + 1) x ((p+1)^2 + 2) doubles (the reference forks the parameter sets with mclapply, each a serial or forked loop
    over the pairs); NA differences become 0 (:1813)."""
    pars = np.asarray(pars, float).reshape(-1)
    zData = np.asarray(zData, float).reshape(-1)
    XData = np.asarray(XData, float).reshape(zData.size, -1)
    n, p = XData.shape
    optn = compLikMats["compLikOptn"]
    nB = len(compLikMats["listBlocks"])
    structs = []
    for k in range(pars.size + 1):
        v = pars.copy()
        if k > 0:
            v[k - 1] += delta
        pb = readParsIAK3D(v, modelx, nud, sdfdType_cd1, sdfdType_cxd0, sdfdType_cxd1, cmeOpt, prodSum, parBnds)
        structs.append(make_pars(pb, modelx, sdfdType_cd1, sdfdType_cxd0, sdfdType_cxd1, cmeOpt))
    T = cl_terms_sets(compLikMats["units"], structs, p + 1, comm, timings)
    n_SUM = _cl_n_sum(compLikMats)
    vals = np.array([cl_reduce_and_tail(t[:(p + 1) ** 2].reshape(p + 1, p + 1), t[-2], n_SUM, t[-1], n, p, optn, nB, useReml)
                     for t in T])
    g = (vals[1:] - vals[0]) / delta
    g[~np.isfinite(g)] = 0.0
    return (g, vals[0]) if rtn_nll else g


def cl_reduce_and_tail(S, lnd_sum, n_SUM, failed, n, p, optn, nB, useReml, comm=None, rtnTerms=False):
    """The exchange step of the composite likelihood -- Reduce('+') over the workers, compositeLikIAK3D.R:63-64 -- as
    one all-reduce of (p+1)^2 + 3 doubles, then the closed-form tail (:181-265) on every rank."""
    buf = np.concatenate([np.asarray(S, float).reshape(-1), [lnd_sum, n_SUM, failed]])
    if comm is not None:
        buf = comm.allreduce_sum(buf)
    S = buf[:(p + 1) ** 2].reshape(p + 1, p + 1); lnd_sum, n_SUM, failed = buf[-3], buf[-2], buf[-1]
    if failed > 0:
        return BADNLL
    if optn == 3:
        n_SUM = n * (nB - 1)
    if optn == 1:
        XziAXz = S / (nB - 1); lndetA = lnd_sum / (nB - 1)
    else:
        XziAXz = S; lndetA = lnd_sum
    if optn in (2, 3):
        if useReml:
            raise ValueError("Eidsvik was only defined for ML, not REML estimation!")
        t = reml_tail(XziAXz, lndetA, n, p, useReml, n_for_cxd=n_SUM, n_for_lndet=n_SUM)
    else:
        t = reml_tail(XziAXz, lndetA, n, p, useReml)
    if rtnTerms:
        return dict(nll=t["nll"], XziAXz_SUM=S, lndetA_SUM=lnd_sum, n_SUM=n_SUM, cxdhat=t["cxdhat"], betahat=t["betahat"])
    return t["nll"]


def cl_plan_units(compLikMats, world=1):
    """Units of one composite-likelihood evaluation and their owner ranks.

    Units: every adjacent pair (optn < 4) and every single block that some non-adjacent pair needs (optn 1, 3: a
    block's terms are added once per non-adjacent pair it belongs to, compositeLikIAK3D.R:157-164) or all single
    blocks (optn 4).  Owners: longest-processing-time on n^3 (the factorisation cost)."""
    optn = compLikMats["compLikOptn"]
    blocks = compLikMats["listBlocks"]
    units = []
    if optn < 4:
        for (k, l) in np.asarray(compLikMats["subsetPairsAdj"]).reshape(-1, 2):
            units.append(dict(kind="pair", weight=1.0, rows=np.concatenate([blocks[k], blocks[l]]), pair=(int(k), int(l))))
    if optn in (1, 3):
        cnt = np.zeros(len(blocks))
        for (k, l) in np.asarray(compLikMats.get("subsetPairsNonadj", np.zeros((0, 2), int))).reshape(-1, 2):
            cnt[k] += 1; cnt[l] += 1
        for b, c in enumerate(cnt):
            if c > 0:
                units.append(dict(kind="block", weight=float(c), rows=blocks[b]))
    elif optn == 4:
        for b in range(len(blocks)):
            units.append(dict(kind="block", weight=1.0, rows=blocks[b]))
    order = np.argsort([-len(u["rows"]) for u in units], kind="stable")
    load = np.zeros(world)
    for i in order:
        r = int(np.argmin(load)); load[r] += float(len(units[i]["rows"])) ** 3
        units[i]["owner"] = r
    return units


def setupIAK3D_CL(xData, dIData, zData, XData, compLikMats, ctx=None, rank=0, world=1, sdfdType_cd1=0, sdfdType_cxd0=0,
                  sdfdType_cxd1=0, sdfdKnots=None):
    """iaCovMatern.R:551-575 + the unit weights of compositeLikIAK3D.R:42-179, as a list of device datasets owned by
    this rank (see :func:`cl_plan_units`)."""
    multi = isinstance(ctx, (list, tuple))          # several GPUs from one process: units dealt to the contexts, all kept here
    ctx = ctx or default_context()
    xData = np.asarray(xData, float); dIData = np.asarray(dIData, float).reshape(-1, 2)
    zData = np.asarray(zData, float).reshape(-1); XData = np.asarray(XData, float).reshape(zData.size, -1)
    units = cl_plan_units(compLikMats, len(ctx) if multi else world)
    out = []
    for u in units:
        if not multi and u["owner"] != rank:
            continue
        rows = u["rows"]
        W = np.column_stack([XData[rows], zData[rows]])
        u["setup"] = SetupMats(xData[rows], dIData[rows], W=W, ctx=ctx[u["owner"]] if multi else ctx, sdfdType_cd1=sdfdType_cd1,
                               sdfdType_cxd0=sdfdType_cxd0, sdfdType_cxd1=sdfdType_cxd1, sdfdKnots=sdfdKnots)
        u["n"] = len(rows)
        out.append(u)
    cl = dict(compLikMats)
    cl["units"] = out
    cl["n_units_total"] = len(units)
    cl["xData"], cl["dIData"] = xData, dIData       # lmmFit$xData, lmmFit$dIData of the block predictors
    cl["sdfdKnots"] = sdfdKnots
    return cl


def gradHessIAK3D(setupMats, parsBTfmd, modelx, sdfdTypes, lnvPars, comm=None):
    """nrUpdatesIAK3D.R:118-216 for C = sum_i exp(lnvPars[i]) dA_i with dA_i the unit-variance terms of setCIAK3D in the
    order cx0, cx1, cd1, cxd0, cxd1[, cme] (the `dA` list the reference is handed, :130-138).  The dataset must carry
    W = [X z].  -> dict(nll, grad, hess, fim, betahat, vbetahat); nll is nan where chol fails.

    comm (parallel.BlockComm, world > 1): every rank holds the same dataset; the six T dC_i products and the 21 pair traces
    are cut over the ranks by tile ownership and one all-reduce of ncomp + ncomp^2 doubles joins them."""
    sm = setupMats
    th = f64(np.asarray(lnvPars, float).reshape(-1))
    m = th.size
    p = sm.q - 1
    P = make_pars(parsBTfmd, modelx, *sdfdTypes, 1 if m == 6 else 0, quirk_cd1_unscaled=False)
    nll = C.c_double()
    beta = np.empty(p); vb = np.empty((p, p), order="F")
    rank, world = (comm.rank, comm.world) if comm is not None else (0, 1)
    zq = np.empty(m); qij = np.empty((m, m)); tr = np.zeros(m + m * m + 1)
    st = sm.ctx.check(sm.ctx.lib.iak3d_gradhess_shard(sm.h, C.byref(P), m, dptr(th), rank, world, C.byref(nll), dptr(beta), dptr(vb),
                                                      dptr(zq), dptr(qij), dptr(tr)), allow=(NOT_SPD, BAD_PARS))
    tr[m + m * m] = 0.0 if st == OK else 1.0         # a failing rank must not leave the others in the all-reduce
    if world > 1:
        tr = comm.allreduce_sum(tr)
    if tr[m + m * m] != 0.0:
        return dict(nll=float("nan"))
    trM, trP = tr[:m], tr[m:m + m * m].reshape(m, m)
    grad = 0.5 * (trM - zq)                          # :197
    fim = 0.5 * trP                                  # :210
    hess = 0.5 * (-trP + 2.0 * qij)                  # :198-207
    hess[np.diag_indices(m)] += 0.5 * (trM - zq)
    return dict(nll=nll.value, grad=grad, hess=hess, fim=fim, betahat=beta.reshape(-1, 1), vbetahat=vb)


def nr_iterate(gh, lnvParInits, smallV):
    """The iteration of nrUpdatesIAK3D (nrUpdatesIAK3D.R:1-116) around a gradient/Hessian callback `gh(lnvPars)`:
    at least 4, at most 20 updates, the first two by Fisher scoring, log-variances below `smallV` that still move down
    are snapped to -20 (:57-65), the final nll rounded to 1e-5 (:107-108).  The reference's Hessian branch tests
    `checkHess$cholC` on the value of solve() (:44), which cannot run; the evident intent is implemented (Newton step from
    the third update on, Fisher scoring if that solve fails)."""
    lnv = np.asarray(lnvParInits, float).reshape(-1).copy()
    t = gh(lnv)
    if not np.isfinite(t["nll"]):
        return dict(lnvPars=lnv, nll=float("nan"), noits=0, errFlag=1)
    sym = lambda M: 0.5 * (M + M.T)
    tol, noits, nllPrev = 1e-6, 0, math.inf
    while noits < 20 and (t["nll"] < nllPrev - tol or noits < 4):
        nllPrev = t["nll"]
        step = None
        if noits >= 2:
            try:
                step = np.linalg.solve(sym(t["hess"]), t["grad"])
            except np.linalg.LinAlgError:
                step = None
        if step is None:
            try:
                step = np.linalg.solve(sym(t["fim"]), t["grad"])
            except np.linalg.LinAlgError:
                return dict(lnvPars=lnv, noits=noits, errFlag=1, **t)
        small = (lnv < smallV) & (step > 0)
        lnv = lnv - step
        lnv[small] = -20.0
        tn = gh(lnv)
        if not np.isfinite(tn["nll"]):
            return dict(lnvPars=lnv, noits=noits, errFlag=1, **t)
        t = tn
        noits += 1
    out = dict(t)
    out.update(lnvPars=lnv, nll_unrounded=t["nll"], nll=float(np.round(t["nll"] / 1e-5) * 1e-5), noits=noits,
               errFlag=2 if noits == 20 else 0)
    return out


def nrUpdatesIAK3D(setupMats, parsBTfmd, modelx, sdfdTypes, lnvParInits, smallV=None):
    """nrUpdatesIAK3D.R:1-116 with every gradHessIAK3D evaluation on the GPU (:func:`gradHessIAK3D`).  The dataset must
    carry W = [X z]; smallV defaults to log(var(z) / 10000) (:5-7)."""
    sm = setupMats
    if smallV is None or (isinstance(smallV, float) and math.isnan(smallV)):
        smallV = math.log(float(np.var(sm._W[:, -1], ddof=1)) / 10000.0)
    return nr_iterate(lambda v: gradHessIAK3D(sm, parsBTfmd, modelx, sdfdTypes, v), lnvParInits, smallV)


# ---------------------------------------------------------------------------------------------
# kept fit + kriging
# ---------------------------------------------------------------------------------------------
class KeptFit:
    """The big matrices of lmmFit (C, iC, iCX, iCz_muhat) as a factorisation kept on the device (iak3d_fit)."""

    def __init__(self, setupMats, parsBTfmd, modelx, sdfdTypes, cmeOpt, betahat, vbetahat, cxdhat=1.0, z_muhat=None):
        sm = setupMats
        self.sm, self.ctx = sm, sm.ctx
        self.p = sm.q - 1
        # C = cxdhat * A (fitIAK3D.R:871): the term the reference adds without its cd1 multiplier (:1052) is scaled too
        P = make_pars(parsBTfmd, modelx, *sdfdTypes, cmeOpt, quirk_scale=cxdhat)
        b = f64(np.asarray(betahat, float).reshape(-1)); V = f64(np.asarray(vbetahat, float).reshape(self.p, self.p))
        h = C.c_void_p()
        if z_muhat is not None:        # lognormal data: the residual is z - setmuIAK3D(...), not z - X betahat
            r = f64(np.asarray(z_muhat, float).reshape(-1))
            st = self.ctx.check(self.ctx.lib.iak3d_fit_create_res(sm.h, C.byref(P), dptr(b), dptr(V), dptr(r), C.byref(h)),
                                allow=(NOT_SPD, BAD_PARS))
        else:
            st = self.ctx.check(self.ctx.lib.iak3d_fit_create(sm.h, C.byref(P), dptr(b), dptr(V), C.byref(h)), allow=(NOT_SPD, BAD_PARS))
        if st != OK:
            raise Iak3dError("fit_create: the fitted covariance matrix is not positive definite" if st == NOT_SPD
                             else "fit_create: invalid parameters")
        self.h = h
        self._close_order = 0          # kept fits go before their datasets
        self.ctx.adopt(self)

    def get_small(self):
        n, p = self.sm.n, self.p
        iCX = np.empty((n, p), order="F"); iCz = np.empty((n, 1), order="F")
        self.ctx.check(self.ctx.lib.iak3d_fit_get(self.h, dptr(iCX), dptr(iCz), None, None))
        return iCX, iCz

    def get_big(self, want_C=True, want_iC=True):
        n = self.sm.n
        Cm = np.empty((n, n), order="F") if want_C else None
        iC = np.empty((n, n), order="F") if want_iC else None
        self.ctx.check(self.ctx.lib.iak3d_fit_get(self.h, None, None, dptr(Cm), dptr(iC)))
        return Cm, iC

    def predict(self, xMap, dIMap, XMap, siteIDMap=None):
        self.stage(xMap, dIMap, XMap, siteIDMap)
        self.enqueue()
        return self.fetch()

    def stage(self, xMap, dIMap, XMap, siteIDMap=None):
        xMap = f64(np.asarray(xMap, float).reshape(-1, 2)); dIMap = f64(np.asarray(dIMap, float).reshape(-1, 2))
        XMap = f64(np.asarray(XMap, float).reshape(xMap.shape[0], self.p))
        if self.sm.nssre > 0:                # site-specific random effects (predictIAK3D.R:9-18, 166-195)
            if siteIDMap is None:
                raise ValueError("Error - lmmFit was fitted with site-specific random effects (ie siteIDData given), but siteIDMap was null in the predict function!")
            sid = site_ids(siteIDMap, self.sm.site_codes)
            Xs = f64(XMap[:, self.sm.cols4ssre])
            self.ctx.check(self.ctx.lib.iak3d_predict_stage_ssre(self.h, xMap.shape[0], dptr(xMap), dptr(dIMap), dptr(XMap), iptr(sid), dptr(Xs)))
            self._m = xMap.shape[0]
            return
        if max(self.sm.sdfdTypes) > 0:       # spline sd functions: basis on the map supports (setCApproxIAK3D2)
            spl = _sdfd_spline_bases(dIMap, self.sm.sdfdTypes, self.sm.sdfdKnots)
            self.ctx.check(self.ctx.lib.iak3d_predict_stage_spl(self.h, xMap.shape[0], dptr(xMap), dptr(dIMap), dptr(XMap), spl_ptrs(spl)))
        else:
            self.ctx.check(self.ctx.lib.iak3d_predict_stage(self.h, xMap.shape[0], dptr(xMap), dptr(dIMap), dptr(XMap)))
        self._m = xMap.shape[0]

    def stage_grid(self, design, xCell, covCell, dIInt, nxPerBatch=2000):
        """The whole map in one staging (predictIAK3D.R:144-155): ncell cells x nint intervals, design rows made on the
        device by `design` (design.DesignModel == makeXvX).  Results come back interval-major: row i * ncell + c."""
        xCell = f64(np.asarray(xCell, float).reshape(-1, 2)); dIInt = f64(np.asarray(dIInt, float).reshape(-1, 2))
        ncell, nint = xCell.shape[0], dIInt.shape[0]
        cv = f64(np.asarray(covCell, float).reshape(ncell, design.ncov)) if design.ncov > 0 else None
        self.ctx.check(self.ctx.lib.iak3d_predict_stage_grid(self.h, design.h, ncell, dptr(xCell), dptr(cv), nint, dptr(dIInt), int(nxPerBatch)))
        self._m = ncell * nint

    def fetch_X(self):
        X = np.empty((self._m, self.p), order="F")
        self.ctx.check(self.ctx.lib.iak3d_predict_fetch_X(self.h, dptr(X)))
        return X

    def enqueue(self):
        self.ctx.check(self.ctx.lib.iak3d_predict_enqueue(self.h))

    def fetch(self):
        z = np.empty(self._m); v = np.empty(self._m)
        self.ctx.check(self.ctx.lib.iak3d_predict_fetch(self.h, dptr(z), dptr(v)))
        return z, v

    def panel_fetch(self, chunk=0):
        """Ckh (setCIAK3D2 of the staged map supports of one chunk against the data) as the predictor's panel kernels
        build it, before the solve overwrites it -- parity tests."""
        rows = C.c_int64()
        cap = min(self._m, 4 * 148 * 128)                    # a chunk never holds more supports than this
        flat = np.empty(cap * self.sm.n)
        self.ctx.check(self.ctx.lib.iak3d_predict_panel_fetch(self.h, int(chunk), dptr(flat), cap, C.byref(rows)))
        return flat[:rows.value * self.sm.n].reshape(rows.value, self.sm.n, order="F")

    def set_option(self, option, value):
        self.ctx.check(self.ctx.lib.iak3d_fit_set_option(self.h, int(option), int(value)))

    def fetch_CkhiCX(self):
        out = np.empty((self._m, self.p), order="F")
        self.ctx.check(self.ctx.lib.iak3d_predict_fetch_CkhiCX(self.h, dptr(out)))
        return out

    def fetch_terms(self):
        """Ckk, sigma2Veck, CkhiCz_muhat, CkhiCChk of the last enqueued batch (predictIAK3D.R:183-208)."""
        out = [np.empty(self._m) for _ in range(4)]
        self.ctx.check(self.ctx.lib.iak3d_predict_fetch_terms(self.h, *[dptr(a) for a in out]))
        return dict(Ckk=out[0], sigma2Veck=out[1], CkhiCz_muhat=out[2], CkhiCChk=out[3])

    def close(self):
        if getattr(self, "h", None) is not None and self.h.value:
            self.ctx.lib.iak3d_fit_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def predictIAK3D(xMap, dIMap, XMap_fn, lmmFit, nxPerBatch=2000, cells=None, vXUMap_fn=None, rqrBTfmdPreds=True, siteIDMap=None):
    """predictIAK3D.R:5-287 for compLikOptn = 0.  `lmmFit` is the rtnAll/attachBigMats dict of :func:`nllIAK3D`;
    `XMap_fn(iDepth, cellIdx)` supplies the design rows (makeXvX is an input of the hot path) and, for lognormal data,
    `vXUMap_fn(iDepth, cellIdx)` their within-support covariances ((m pU) x pU).  Rows of XMap containing NaN are skipped
    and stay NaN (predictIAK3D.R:160-163).  `cells` restricts the work to a slice of the map (prediction-grid tiles per
    GPU).  -> dict(zMap, vMap, pi90LMap, pi90UMap) [ndIMap x nxMap]."""
    xMap = np.asarray(xMap, float).reshape(len(xMap), -1)
    dIMap = np.round(np.asarray(dIMap, float).reshape(-1, 2), 2)       # :29
    nd, nx = dIMap.shape[0], xMap.shape[0]
    fit = lmmFit["fit"]
    lnN = bool(lmmFit.get("lnTfmdData", False))
    idx_all = np.arange(nx) if cells is None else np.asarray(cells)
    zMap = np.full((nd, nx), np.nan); vMap = np.full((nd, nx), np.nan)
    for i in range(nd):
        XM = np.asarray(XMap_fn(i, idx_all), float)
        ok = ~np.isnan(XM.mean(axis=1))
        if not ok.any():
            continue
        sel = idx_all[ok]
        dI = np.repeat(dIMap[i:i + 1], sel.size, axis=0)                # :150
        z, v = fit.predict(xMap[sel], dI, XM[ok], None if siteIDMap is None else np.asarray(siteIDMap)[sel])
        if lnN:                                                         # :229-255
            t = fit.fetch_terms()
            iU = lmmFit["iU"]
            vX = None
            if len(iU):
                pU = len(iU)
                vX = np.asarray(vXUMap_fn(i, idx_all), float).reshape(idx_all.size, pU, pU)[ok].reshape(-1, pU)
            mu = setmuIAK3D(XM[ok], vX, iU, lmmFit["betahat"], t["Ckk"], t["sigma2Veck"]).reshape(-1)
            z = mu + t["CkhiCz_muhat"]
            v = t["Ckk"] - t["CkhiCChk"]                                # no beta uncertainty in the lognormal case (:251)
        zMap[i, sel] = z; vMap[i, sel] = v
    with np.errstate(invalid="ignore"):
        s_ = np.sqrt(vMap)
    lo, hi = zMap - 1.645 * s_, zMap + 1.645 * s_                       # :272-273
    if lnN and rqrBTfmdPreds:                                           # :276-284
        zMap = np.exp(zMap + 0.5 * vMap)
        lo, hi = np.exp(lo), np.exp(hi)
        vMap = (np.exp(vMap) - 1.0) * zMap ** 2
    return dict(zMap=zMap, vMap=vMap, pi90LMap=lo, pi90UMap=hi)


def profilePredictIAK3D(xMap, dIMap, XMap, lmmFit, vXUMap=None, rqrBTfmdPreds=True, rtnBigMats=False):
    """predictIAK3D.R:292-504 for compLikOptn = 0: predictions for a full profile -- many depth intervals `dIMap` at ONE
    location `xMap` -- with the design rows `XMap` given (makeXvX is an input of the hot path).  Returns what the reference
    returns: zMap, vMap, pi90LMap, pi90UMap, XMap, muhatMap, Ckk and, with `rtnBigMats`, Ckh (setCIAK3D2) and iC."""
    xMap = np.asarray(xMap, float).reshape(-1, 2)
    if xMap.shape[0] != 1:
        raise ValueError("Error - profilePredictIAK3D is for a single location at a time (ie xMap should have one row)!")
    dIMap = np.round(np.asarray(dIMap, float).reshape(-1, 2), 2)        # :349
    nd = dIMap.shape[0]
    XMap = np.asarray(XMap, float).reshape(nd, -1)
    xRep = np.repeat(xMap, nd, axis=0)                                  # :395-397
    fit = lmmFit["fit"]
    z, v = fit.predict(xRep, dIMap, XMap)
    t = fit.fetch_terms()
    lnN = bool(lmmFit.get("lnTfmdData", False))
    if lnN:                                                             # :468-472, 488-491
        muhat = setmuIAK3D(XMap, vXUMap, lmmFit["iU"], lmmFit["betahat"], t["Ckk"], t["sigma2Veck"]).reshape(-1)
        z = muhat + t["CkhiCz_muhat"]; v = t["Ckk"] - t["CkhiCChk"]
    else:
        muhat = (XMap @ np.asarray(lmmFit["betahat"], float).reshape(-1, 1)).reshape(-1)
    with np.errstate(invalid="ignore"):
        s_ = np.sqrt(v)
    lo, hi = z - 1.645 * s_, z + 1.645 * s_
    if lnN and rqrBTfmdPreds:
        z = np.exp(z + 0.5 * v); lo, hi = np.exp(lo), np.exp(hi); v = (np.exp(v) - 1.0) * z ** 2
    out = dict(zMap=z, vMap=v, pi90LMap=lo, pi90UMap=hi, XMap=XMap, muhatMap=muhat, Ckk=t["Ckk"], lmmFit=lmmFit)
    if rtnBigMats:
        sm = lmmFit["setupMats"]
        types = (lmmFit["sdfdType_cd1"], lmmFit["sdfdType_cxd0"], lmmFit["sdfdType_cxd1"])
        out["Ckh"] = setCIAK3D2(lmmFit["parsBTfmd"], lmmFit["modelx"], *types, lmmFit["cmeOpt"], SetupMats2(xRep, dIMap, None, None, setup2=sm))
        out["iC"] = fit.get_big(want_C=False, want_iC=True)[1]
    return out

# ---------------------------------------------------------------------------------------------
# composite-likelihood block prediction (compositeLikIAK3D.R:471-607; dispatched at predictIAK3D.R:211-226)
# ---------------------------------------------------------------------------------------------
def getBlocksFromLocations(xMap, compLikMats, ctx=None):
    """Nearest block centroid, first index on ties (compositeLikIAK3D.R:803-812): one device pass (blocks.py)."""
    from .blocks import getBlocksFromLocations as _g
    return _g(xMap, compLikMats, ctx=ctx)


def predMatsIAK3D_CL_ByBlock(z_muhat, XData, xMap, dIMap, lmmFit):
    """compositeLikIAK3D.R:471-607 (compLikOptn 1 and 4): for the map supports of block k, the kriging terms against the
    data of every adjacent pair (k, l), averaged over the neighbours l.  Each pair is one kept factorisation on the
    GPU -- the pair dataset the likelihood was fitted with -- shared by the map supports of both of its blocks.
    -> dict(diagCkk, diagCkhiCChk, CkhiCX, CkhiCz_muhat)."""
    cl = lmmFit["compLikMats"]
    blocks = cl["listBlocks"]
    pairs = np.asarray(cl["subsetPairsAdj"]).reshape(-1, 2)
    xMap = np.asarray(xMap, float).reshape(-1, 2); dIMap = np.asarray(dIMap, float).reshape(-1, 2)
    m, p = xMap.shape[0], XData.shape[1]
    blockMap = getBlocksFromLocations(xMap, cl)
    nadj = np.zeros(len(blocks))
    for k, l in pairs:
        nadj[k] += 1; nadj[l] += 1
    used = np.unique(blockMap)
    if np.any(nadj[used] == 0):
        raise ValueError("Error - every block must have a neighbour - check this!")
    unit_of = {}
    for u in cl["units"]:
        if u["kind"] == "pair":
            unit_of[tuple(u["pair"])] = u
    types = (lmmFit["sdfdType_cd1"], lmmFit["sdfdType_cxd0"], lmmFit["sdfdType_cxd1"])
    Ckk = np.full(m, np.nan); q = np.zeros(m); cz = np.zeros(m); cx = np.zeros((m, p))
    z_muhat = np.asarray(z_muhat, float).reshape(-1)
    for k, l in pairs:
        pts = np.nonzero((blockMap == k) | (blockMap == l))[0]
        if pts.size == 0:
            continue
        u = unit_of.get((int(k), int(l)))
        if u is None:
            raise ValueError("composite-likelihood prediction: the pair dataset is not on this rank (setupIAK3D_CL)")
        rows = u["rows"]
        # parsBTfmd are the rescaled parameters; C is rebuilt from them (setCIAK3D of the joined supports, :540-552)
        fit = KeptFit(u["setup"], lmmFit["parsBTfmd"], lmmFit["modelx"], types, lmmFit["cmeOpt"], np.zeros(p), np.eye(p), 1.0,
                      z_muhat=z_muhat[rows])
        try:
            fit.set_option(_lib.FIT_KEEP_CKHICX, 1); fit.set_option(_lib.FIT_CROSS_SQUARE, 1)
            fit.stage(xMap[pts], dIMap[pts], np.zeros((pts.size, p)))
            fit.enqueue(); fit.fetch()
            t = fit.fetch_terms()
            Ckk[pts] = t["Ckk"]; q[pts] += t["CkhiCChk"]; cz[pts] += t["CkhiCz_muhat"]; cx[pts] += fit.fetch_CkhiCX()
        finally:
            fit.close()
    w = nadj[blockMap]
    return dict(diagCkk=Ckk, diagCkhiCChk=q / w, CkhiCX=cx / w[:, None], CkhiCz_muhat=cz / w)


def predMatsIAK3D_CLEV_ByBlock(z_muhat, XData, xMap, dIMap, lmmFit):
    """compositeLikIAK3D.R:609-798 (compLikOptn 2 and 3, Eidsvik et al.): the map supports of block k are predicted
    jointly from the pairwise conditionals with every adjacent block.  All matrices -- the joined-support covariance
    C0klAll (:671), its sub-blocks, the pair inverses, A0, Bk0, J0 -- stay on the GPU as :class:`DeviceMat`; products run
    on the tile engine (DMMA), `solve` / `chol2inv(chol(.))` are the factorisation + explicit inverse.  The
    `loadFromFit` branch (:688-707) reads the same matrices from the fit object and is folded into the rebuild branch.
    -> dict(zk, vk); X betahat is added by the caller (predictIAK3D.R:231-234)."""
    from .devmat import DeviceMat, cov_mat
    cl = lmmFit["compLikMats"]
    blocks = cl["listBlocks"]
    pairs = np.asarray(cl["subsetPairsAdj"]).reshape(-1, 2)
    optn = cl["compLikOptn"]
    xData = np.asarray(lmmFit.get("xData", cl.get("xData")), float); dIData = np.asarray(lmmFit.get("dIData", cl.get("dIData")), float)
    xMap = np.asarray(xMap, float).reshape(-1, 2); dIMap = np.asarray(dIMap, float).reshape(-1, 2)
    nx = xMap.shape[0]
    types = (lmmFit["sdfdType_cd1"], lmmFit["sdfdType_cxd0"], lmmFit["sdfdType_cxd1"])
    P = make_pars(lmmFit["parsBTfmd"], lmmFit["modelx"], *types, lmmFit["cmeOpt"])
    ctx = cl["units"][0]["setup"].ctx if cl.get("units") else default_context()
    z_muhat = np.asarray(z_muhat, float).reshape(-1)
    blockMap = getBlocksFromLocations(xMap, cl)
    zk = np.full(nx, np.nan); vk = np.full(nx, np.nan)
    mm = DeviceMat.matmul_nt
    for k in range(int(blockMap.max()) + 1):
        pts = np.nonzero(blockMap == k)[0]
        m = pts.size
        if m == 0:
            continue
        adj = [int(b) for a, b in pairs if a == k] + [int(a) for a, b in pairs if b == k]   # getAdjacentBlocks, :817-827
        if not adj:
            raise ValueError("Error - every block must have a neighbour - check this!")
        iDatak = np.asarray(blocks[k]); nk = iDatak.size
        datL = [np.asarray(blocks[l]) for l in adj]
        i2L, cnt = [], m + nk
        for d in datL:
            i2L.append((cnt, d.size)); cnt += d.size
        rows = np.concatenate([iDatak] + datL)
        sm = SetupMats(np.vstack([xMap[pts], xData[rows]]), np.vstack([dIMap[pts], dIData[rows]]), ctx=ctx, sdfdType_cd1=types[0],
                       sdfdType_cxd0=types[1], sdfdType_cxd1=types[2], sdfdKnots=lmmFit.get("sdfdKnots"))      # :668
        Call = cov_mat(sm, P)                                                                                # :671
        sm.close()
        if Call is None:
            raise ValueError("predMatsIAK3D_CLEV_ByBlock: invalid covariance parameters")
        C0k = Call.sub(0, 0, m + nk, m + nk)                           # :680
        C00 = Call.sub(0, 0, m, m)
        A0 = DeviceMat(m, m, ctx); b0 = DeviceMat(m, 1, ctx); Bk0 = DeviceMat(m, m + nk, ctx)
        QL = []
        for d, (s2, n2) in zip(datL, i2L):                             # :692-745
            kl = [(m, nk), (s2, n2)]
            iS = Call.gather_blocks(kl, kl).spd_inverse()              # solve(Ckl_klThis, .), :717
            C0kl = Call.gather_blocks([(0, m)], kl)                    # C0klAll[i0, c(i1,i2)]
            Tt = mm(C0kl, iS)                                          # t(iCkl_kl_Ckl_0This)
            iS.close()
            IF = DeviceMat(m, m, ctx).assign(C00).gemm_nt(C0kl, Tt, alpha=-1.0, accumulate=True)   # :720
            C0kl.close()
            iIF = IF.symmetrised().spd_inverse()                       # :722-725
            IF.close()
            T = Tt.t(); Tt.close()
            Q12 = mm(iIF, T, alpha=-1.0)                               # :726
            T.close()
            QL.append(Q12.sub(0, nk, m, n2))                           # :730
            zrow = DeviceMat.from_host(z_muhat[np.concatenate([iDatak, d])].reshape(1, -1), ctx)
            A0.assign(iIF, beta=1.0)                                   # :737-739
            b0.gemm_nt(Q12, zrow, alpha=-1.0, accumulate=True)
            Bk0.assign(iIF, beta=1.0)
            Bk0.assign(Q12, dc0=m, nr=m, nc=nk, beta=1.0)
            for M in (iIF, Q12, zrow):
                M.close()
        if optn == 3:                                                  # :748-770: non-adjacent pairs taken as independent
            nnon = float(len(blocks) - len(adj) - 1)
            iCkk = Call.sub(m, m, nk, nk).spd_inverse()
            C0_k = Call.sub(0, m, m, nk)
            C0kiC = mm(C0_k, iCkk)
            IF = DeviceMat(m, m, ctx).assign(C00).gemm_nt(C0kiC, C0_k, alpha=-1.0, accumulate=True)
            iIF = IF.symmetrised().spd_inverse()
            C0kiCt = C0kiC.t()
            iIFC = mm(iIF, C0kiCt)
            zrow = DeviceMat.from_host(z_muhat[iDatak].reshape(1, -1), ctx)
            A0.assign(iIF, alpha=nnon, beta=1.0)
            b0.gemm_nt(iIFC, zrow, alpha=nnon, accumulate=True)
            Bk0.assign(iIF, alpha=nnon, beta=1.0)
            Bk0.assign(iIFC, dc0=m, alpha=-nnon, beta=1.0)
            for M in (iCkk, C0_k, C0kiC, IF, iIF, C0kiCt, iIFC, zrow):
                M.close()
        BC = mm(Bk0, C0k)                                              # C0k_0k is symmetric: Bk0 %*% C0k_0k
        J0 = mm(BC, Bk0)                                               # :772
        BC.close()
        for a, (s2a, n2a) in enumerate(i2L):                           # :773-781
            Cl = Call.sub(s2a, 0, n2a, m + nk)                         # t(C0k_lList[[a]]) by symmetry of C0klAll
            BCl = mm(Bk0, Cl); Cl.close()
            J0.gemm_nt(BCl, QL[a], alpha=2.0, accumulate=True); BCl.close()
            for b_, (s2b, n2b) in enumerate(i2L):
                Cba = Call.sub(s2b, s2a, n2b, n2a)                     # t(C0klAll[i2a, i2b])
                QC = mm(QL[a], Cba); Cba.close()
                J0.gemm_nt(QC, QL[b_], accumulate=True); QC.close()
        iA0 = A0.symmetrised().spd_inverse()                           # :784-788
        J0s = J0.symmetrised()
        b0t = b0.t()
        zkd = mm(iA0, b0t)                                             # :790
        zk[pts] = zkd.to_host().reshape(-1)
        JA = mm(J0s, iA0)
        vk[pts] = iA0.coldot(JA)                                       # :795
        for M in [Call, C0k, C00, A0, b0, Bk0, J0, iA0, J0s, b0t, zkd, JA] + QL:
            M.close()
    ctx.check(ctx.lib.iak3d_mat_trim(ctx.h))
    return dict(zk=zk, vk=vk)


def predictIAK3D_CL(xMap, dIMap, XMap_fn, lmmFit, cells=None):
    """predictIAK3D (predictIAK3D.R:5-287) when the model was fitted by composite likelihood with compLikOptn 1 or 4
    (:211-226 -> predMatsIAK3D_CL_ByBlock) or with the Eidsvik options 2 or 3 (-> predMatsIAK3D_CLEV_ByBlock, no
    fixed-effect term in the variance, :231-234).  `lmmFit` is the rtnAll dict of :func:`nllIAK3D_CL`."""
    optn = lmmFit["compLikMats"]["compLikOptn"]
    xMap = np.asarray(xMap, float).reshape(len(xMap), -1)
    dIMap = np.round(np.asarray(dIMap, float).reshape(-1, 2), 2)
    nd, nx = dIMap.shape[0], xMap.shape[0]
    XData, betahat, vbetahat = lmmFit["XData"], lmmFit["betahat"], lmmFit["vbetahat"]
    z_muhat = lmmFit["zData"].reshape(-1) - (XData @ betahat).reshape(-1)
    idx_all = np.arange(nx) if cells is None else np.asarray(cells)
    zMap = np.full((nd, nx), np.nan); vMap = np.full((nd, nx), np.nan)
    for i in range(nd):
        XM = np.asarray(XMap_fn(i, idx_all), float)
        ok = ~np.isnan(XM.mean(axis=1))
        if not ok.any():
            continue
        sel = idx_all[ok]
        if optn in (2, 3):
            t = predMatsIAK3D_CLEV_ByBlock(z_muhat, XData, xMap[sel], np.repeat(dIMap[i:i + 1], sel.size, axis=0), lmmFit)
            zMap[i, sel] = (XM[ok] @ betahat).reshape(-1) + t["zk"]
            vMap[i, sel] = t["vk"]
            continue
        t = predMatsIAK3D_CL_ByBlock(z_muhat, XData, xMap[sel], np.repeat(dIMap[i:i + 1], sel.size, axis=0), lmmFit)
        zMap[i, sel] = (XM[ok] @ betahat).reshape(-1) + t["CkhiCz_muhat"]
        d = XM[ok] - t["CkhiCX"]
        vMap[i, sel] = t["diagCkk"] - t["diagCkhiCChk"] + np.sum(d * (vbetahat @ d.T).T, axis=1)
    with np.errstate(invalid="ignore"):
        s_ = np.sqrt(vMap)
    return dict(zMap=zMap, vMap=vMap, pi90LMap=zMap - 1.645 * s_, pi90UMap=zMap + 1.645 * s_)
