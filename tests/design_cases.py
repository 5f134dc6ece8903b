"""Seeded inputs of the makeXvX parity tests (shared by the CPU oracle tests and the GPU tests)."""
import numpy as np

ALLKNOTS = np.array([0.0, 0.3, 1.0, 2.0, 7.13])          # egEdgeroi.R:106


def intervals(n, seed):
    rng = np.random.default_rng(seed)
    a = np.round(rng.uniform(0.0, 1.6, n), 2)
    t = rng.choice([0.05, 0.10, 0.15, 0.30, 0.40, 1.0], n)
    return np.column_stack([a, np.round(a + t, 2)])


def mlr_case(n, seed, factor=True):
    rng = np.random.default_rng(seed)
    cov = {"elev": rng.normal(300.0, 40.0, n), "twi": rng.normal(0.0, 1.0, n)}
    levels = {}
    if factor:
        cov["geol"] = rng.integers(1, 5, n).astype(float)       # factor with 4 levels -> 3 dummy columns
        levels["geol"] = 4
    return cov, levels, intervals(n, seed + 1)


def cubist_model(with_spline=True, opt_dSpline=1):
    """Two committees; committee 1 = three rules split on twi and on depth (0.3, 0.9), committee 2 = two rules split on depth
    (0.5) with a category condition; rule columns as cubist2XSetup lays them out (const first, then the rule's regressors)."""
    names = ["elev", "twi", "geol", "dIMidPts"]
    E, T, G, D = 0, 1, 2, 3
    rules = [
        dict(committee=1, conds=[(T, "<=", 0.2), (D, "<=", 0.3)], const=True, coef_iv=[E, D]),
        dict(committee=1, conds=[(T, "<=", 0.2), (D, ">", 0.3)], const=True, coef_iv=[E, T, D]),
        dict(committee=1, conds=[(T, ">", 0.2), (D, ">", 0.9)], const=True, coef_iv=[D]),
        dict(committee=1, conds=[(T, ">", -0.5)], const=False, coef_iv=[E]),
        dict(committee=2, conds=[(D, "<=", 0.5), (G, "in", (1.0, 3.0))], const=True, coef_iv=[T, D]),
        dict(committee=2, conds=[(D, ">", 0.5)], const=True, coef_iv=[E]),
    ]
    comms, namesX = [], []
    for iR, r in enumerate(rules):
        nc = (1 if r["const"] else 0) + len(r["coef_iv"])
        comms += [r["committee"]] * nc
        namesX += ([f"const_R{iR + 1}"] if r["const"] else []) + [f"{names[v]}_R{iR + 1}" for v in r["coef_iv"]]
    m = dict(namesDataFit=names, rules=rules, comms4XCubist=comms, committees=2, dcol=D)
    if with_spline:
        nspl = (len(ALLKNOTS) - 2) if opt_dSpline == 1 else (len(ALLKNOTS) - 2 + 3)
        namesX = namesX + [f"dSpline.{k + 1}" for k in range(nspl)]
    m["namesX"] = namesX
    return m


def cubist_data(n, seed):
    rng = np.random.default_rng(seed)
    D = np.column_stack([rng.normal(300.0, 40.0, n), rng.normal(0.0, 1.0, n), rng.integers(1, 5, n).astype(float), np.zeros(n)])
    return D, intervals(n, seed + 1)
