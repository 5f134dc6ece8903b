"""Shared builders for the parity tests: seeded inputs + the oracle side of each comparison."""
import numpy as np

from iak3d_b200 import synth
from oracle import iak3d_oracle as orc

PARBNDS = dict(axBnds=(0.05, 60.0), nuxBnds=(0.05, 20.0), adBnds=(0.02, 1.2), tau2Bnds=(0.01, 20.0),
               sxd1Bnds=(1e-8, 1.0 - 1e-8), maxd=2.0)

# (label, default_pars kind, nonstat, nud)
MODEL_CASES = [
    ("matern0.5-stat", "matern0.5", False, 0.5),
    ("matern1.5-stat-nud1.5", "matern1.5", False, 1.5),
    ("matern0.8-bessel-nud2.5", "matern0.8", False, 2.5),
    ("matern0.5-nonstat", "matern0.5", True, 0.5),
    ("matern1.5-nonstat-nud2.5", "matern1.5", True, 2.5),
    ("edgeroi-spherical", "edgeroi", False, 0.5),
    ("edgeroi-nonstat", "edgeroi", True, 1.5),
    ("wendland", "wendland", False, 0.5),
    ("nugget", "nugget", False, 0.5),
]


def case_pars(kind, nonstat, nud):
    P, modelx, types, cmeOpt = synth.default_pars(kind, nonstat)
    P = dict(P); P["nud"] = nud
    return P, modelx, types, cmeOpt


def relerr(a, b):
    a = np.asarray(a, float); b = np.asarray(b, float)
    den = np.maximum(np.abs(b), 1e-300)
    return float(np.max(np.abs(a - b) / den)) if a.size else 0.0


def scaled_err(a, b):
    """max |a-b| / max|b|: for matrices whose entries may be exact zeros (compact-support models)."""
    a = np.asarray(a, float); b = np.asarray(b, float)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)) if a.size else 0.0


def oracle_setup(ds):
    return orc.setupIAK3D(ds["xData"], ds["dIData"])


# spline sd functions (sdfdType > 0 -> setCApproxIAK3D): knots and parameter sets used by the CPU and GPU tests
SPLINE_KNOTS = dict(bdryKnots=(0.0, 1.6), intKnots_cd1=np.array([0.2, 0.7]), intKnots_cxd0=np.array([0.15, 0.45, 1.0]),
                    intKnots_cxd1=np.array([0.5]))
# (label, kind, types, spline coefficient lists per component)
SPLINE_CASES = [
    ("all-spline", "matern0.5", (2, 3, 1), dict(sdfdPars_cd1=[0.8, 0.55], sdfdPars_cxd0=[1.1, 0.7, 0.4], sdfdPars_cxd1=[0.6])),
    ("mixed-stat-exp-spline", "matern1.5", (0, -1, 1), dict(sdfdPars_cd1=[], sdfdPars_cxd0=[0.6, 1.7], sdfdPars_cxd1=[0.35])),
    ("edgeroi-no-cd1", "edgeroi", (-9, 3, 0), dict(sdfdPars_cd1=[], sdfdPars_cxd0=[0.9, 0.8, 0.5], sdfdPars_cxd1=[])),
    ("nugget-spline", "nugget", (2, 3, 0), dict(sdfdPars_cd1=[1.2, 0.9], sdfdPars_cxd0=[0.9, 0.8, 0.5], sdfdPars_cxd1=[])),
]


def spline_case_pars(kind, types, coefs, nud=0.5):
    P, modelx, _, cmeOpt = synth.default_pars(kind, False)
    P = dict(P); P["nud"] = nud
    for k, v in coefs.items():
        P[k] = np.asarray(v, float)
    if types[0] == -9:
        P["cd1"] = 0.0
    elif P.get("cd1", 0.0) == 0.0:
        P["cd1"] = 0.15
    return P, modelx, types, cmeOpt


def lognormal_case(n, seed, pU=2):
    """Synthetic lognormal-data inputs: the last pU columns of X vary inside the sample supports (iU, 0-based) with
    within-support covariance blocks vXU ((n pU) x pU, symmetric positive semi-definite)."""
    ds = synth.make_dataset(n, seed)
    rng = np.random.default_rng(seed + 99)
    p = ds["XData"].shape[1]
    iU = list(range(p - pU, p))
    B = rng.standard_normal((n, pU, pU)) * 0.15
    v = np.einsum("nij,nkj->nik", B, B) * (ds["dIData"][:, 1] - ds["dIData"][:, 0])[:, None, None]
    z = 0.4 * ds["zData"] / max(np.std(ds["zData"]), 1e-9) + 1.0          # log-scale response
    return ds, z, iU, v.reshape(n * pU, pU)
