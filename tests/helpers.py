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
