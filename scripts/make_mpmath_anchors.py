"""Writes tests/golden/mpmath_anchors.json: anchors for the stages no reference-held vector pins -- the setCIAK3D
assembly, the REML / ML tail and the universal-kriging algebra -- computed from the DEFINITIONS of the model in 40-digit
arithmetic (mpmath), by code that shares nothing with the repository:

  * this script imports neither `oracle/` nor `iak3d_b200/` (only mpmath, numpy for the seeded inputs, json);
  * the depth terms are Gauss-Legendre quadratures of the defining double integral
        Phi(I, J) = 1 / (|I| |J|)  int_I int_J  s(x) s(y) rho(|x - y|) dy dx          (iaCovMatern.R:4-10 states it)
    split at the kink x = y -- NOT the reference's closed form (intRy / intTy, iaCovMatern.R:219-322);
  * the horizontal terms are the textbook Matern correlation 2^(1-nu) / Gamma(nu) t^nu K_nu(t), t = sqrt(2 nu) h / a, with
    mpmath's besselk and gamma, and the spherical model;
  * the likelihood comes from a dense mpmath Cholesky of A and the REML / ML formula of the paper (fitIAK3D.R:849-878 is the
    reference's statement of it);
  * predictions come from the textbook universal-kriging system  [C X; X' 0] [lambda; mu] = [c_k; x_k],
        z = lambda' z_data,   v = C_kk - 2 lambda' c_k + lambda' C lambda,
    not from the reference's beta-hat / v-beta-hat form (predictIAK3D.R:202-249).

What IS taken from the reference are its modelling decisions, each cited where it is applied: which terms make up A
(fitIAK3D.R:973-1083), the profiled variance, that C_kk carries the measurement-error variance while c_k does not
(predictIAK3D.R:189, fitIAK3D.R:1243), and -- case C only -- that setCIAK3D adds a non-stationary cd1 term without its
multiplier (fitIAK3D.R:1052).

    python scripts/make_mpmath_anchors.py          # ~5 minutes; rewrites the fixture
"""
import json
import os
import sys

import mpmath as mp
import numpy as np

mp.mp.dps = 40
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")


# ---------------------------------------------------------------------------------------------------------------------
# Gauss-Legendre nodes in 40-digit arithmetic (Newton on the three-term recurrence)
# ---------------------------------------------------------------------------------------------------------------------
def gl_nodes(m):
    x0 = np.polynomial.legendre.leggauss(m)[0]
    xs, ws = [], []
    for g in x0:
        x = mp.mpf(float(g))
        for _ in range(8):
            p0, p1 = mp.mpf(1), x
            for k in range(2, m + 1):
                p0, p1 = p1, ((2 * k - 1) * x * p1 - (k - 1) * p0) / k
            dp = m * (x * p1 - p0) / (x * x - 1)
            x = x - p1 / dp
        p0, p1 = mp.mpf(1), x
        for k in range(2, m + 1):
            p0, p1 = p1, ((2 * k - 1) * x * p1 - (k - 1) * p0) / k
        dp = m * (x * p1 - p0) / (x * x - 1)
        xs.append(x); ws.append(2 / ((1 - x * x) * dp * dp))
    return xs, ws


GL = {m: gl_nodes(m) for m in (20, 28)}


def gl_int(f, lo, hi, m):
    if hi <= lo:
        return mp.mpf(0)
    xs, ws = GL[m]
    h, c = (hi - lo) / 2, (hi + lo) / 2
    return h * mp.fsum(w * f(c + h * x) for x, w in zip(xs, ws))


# ---------------------------------------------------------------------------------------------------------------------
# the model, from its definition
# ---------------------------------------------------------------------------------------------------------------------
def rho_d(h, ad, nud):
    """Matern correlation of the depth process at lag h >= 0 for nu = 1/2, 3/2, 5/2."""
    if nud == 0.5:
        return mp.exp(-h / ad)
    if nud == 1.5:
        t = mp.sqrt(3) * h / ad
        return (1 + t) * mp.exp(-t)
    if nud == 2.5:
        t = mp.sqrt(5) * h / ad
        return (1 + t + t * t / 3) * mp.exp(-t)
    raise ValueError(nud)


def sd_fn(d, tau):
    """standard-deviation function of depth: 1 (stationary) or (1 - tau1) + tau1 exp(-tau2 d)  (iaCovMatern.R:36-46)"""
    if tau is None:
        return mp.mpf(1)
    return (1 - tau[0]) + tau[0] * mp.exp(-tau[1] * d)


def phi_d(I, J, ad, nud, tau, m=20):
    """Increment-averaged depth covariance of intervals I = (a, b), J = (c, d) by nested Gauss-Legendre quadrature; the
    inner integral is split at y = x and the outer one at the end points of J, so every piece is analytic."""
    a, b = I; c, d = J
    def inner(x):
        sx = sd_fn(x, tau)
        f = lambda y: sd_fn(y, tau) * rho_d(abs(x - y), ad, nud)
        if c < x < d:
            return sx * (gl_int(f, c, x, m) + gl_int(f, x, d, m))
        return sx * gl_int(f, c, d, m)
    cuts = sorted({a, b} | {t for t in (c, d) if a < t < b})
    tot = mp.fsum(gl_int(inner, lo, hi, m) for lo, hi in zip(cuts[:-1], cuts[1:]))
    return tot / ((b - a) * (d - c))


def av_var(I, tau, m=20):
    """interval average of s(d)^2 (the variance of a point inside the interval, fitIAK3D.R:1083)"""
    a, b = I
    return gl_int(lambda x: sd_fn(x, tau) ** 2, a, b, m) / (b - a)


def phi_x(h, modelx, ax, nux):
    if h == 0:
        return mp.mpf(1)
    if modelx == "spherical":
        r = h / ax
        return mp.mpf(0) if r > 1 else 1 - mp.mpf(3) / 2 * r + r ** 3 / 2
    if modelx == "matern":
        t = mp.sqrt(2 * nux) * h / ax
        return 2 ** (1 - nux) / mp.gamma(nux) * t ** nux * mp.besselk(nux, t)
    raise ValueError(modelx)


M = lambda v: mp.mpf(float(v))     # inputs are float64 numbers, taken exactly


class Model:
    def __init__(self, case):
        self.c = case
        P = case["pars"]
        self.P = {k: (M(v) if isinstance(v, float) else v) for k, v in P.items()}
        self.modelx = case["modelx"]
        self.nud = case["nud"]
        self.tau = {k: (None if case["tau"][k] is None else (M(case["tau"][k][0]), M(case["tau"][k][1]))) for k in ("cd1", "cxd0", "cxd1")}
        self.has = {"cd1": case["types"][0] != -9, "cxd0": True, "cxd1": True}
        self._phi = {}
        self._av = {}

    def Phi(self, comp, I, J):
        tau = self.tau[comp]
        key = (None if tau is None else (float(tau[0]), float(tau[1])), tuple(map(float, I)), tuple(map(float, J)))
        if key not in self._phi:
            v = phi_d(I, J, self.P["ad"], self.nud, tau)
            self._phi[key] = v
            self._phi[(key[0], key[2], key[1])] = v
        return self._phi[key]

    def avvar(self, comp, I):
        tau = self.tau[comp]
        key = (None if tau is None else (float(tau[0]), float(tau[1])), tuple(map(float, I)))
        if key not in self._av:
            self._av[key] = av_var(I, tau)
        return self._av[key]

    def cov(self, x1, I1, x2, I2, same_obs, square):
        """Covariance of the interval averages at (x1, I1) and (x2, I2): the six terms of fitIAK3D.R:973-1071.
        same_obs adds the measurement-error variance; `square` selects setCIAK3D's treatment of a non-stationary cd1
        term (added WITHOUT cd1, fitIAK3D.R:1052) against setCIAK3D2's (with it, :1211)."""
        P = self.P
        h = mp.sqrt((x1[0] - x2[0]) ** 2 + (x1[1] - x2[1]) ** 2)
        same_x = (x1[0] == x2[0]) and (x1[1] == x2[1])
        px = phi_x(h, self.modelx, P["ax"], P.get("nux"))
        v = mp.mpf(0)
        if same_obs:
            v += P["cme"]
        if same_x:
            v += P["cx0"] + P["cxd0"] * self.Phi("cxd0", I1, I2)
        v += P["cx1"] * px
        if self.has["cd1"]:
            k = P["cd1"]
            if square and self.tau["cd1"] is not None and self.c.get("cd1_quirk", False):
                k = mp.mpf(1) * self.c.get("quirk_scale_mp", mp.mpf(1))
            v += k * self.Phi("cd1", I1, I2)
        v += P["cxd1"] * px * self.Phi("cxd1", I1, I2)
        return v


def build_case(case):
    n, p = case["n"], case["p"]
    x = [(M(a), M(b)) for a, b in case["xData"]]
    dI = [(M(a), M(b)) for a, b in case["dIData"]]
    X = mp.matrix([[M(v) for v in row] for row in case["XData"]])
    z = mp.matrix([M(v) for v in case["zData"]])
    mod = Model(case)
    A = mp.zeros(n, n)
    for i in range(n):
        for j in range(i + 1):
            A[i, j] = mod.cov(x[i], dI[i], x[j], dI[j], i == j, True)
            A[j, i] = A[i, j]
        print(f"  {case['name']}: row {i + 1}/{n}", end="\r", file=sys.stderr)
    L = mp.cholesky(A)
    lndetA = 2 * mp.fsum(mp.log(L[i, i]) for i in range(n))
    W = mp.matrix(n, p + 1)
    for i in range(n):
        for k in range(p):
            W[i, k] = X[i, k]
        W[i, p] = z[i]
    iAW = mp.matrix(n, p + 1)
    for k in range(p + 1):          # L y = w, L' x = y, column by column
        y = [mp.mpf(0)] * n
        for i in range(n):
            y[i] = (W[i, k] - mp.fsum(L[i, j] * y[j] for j in range(i))) / L[i, i]
        for i in range(n - 1, -1, -1):
            iAW[i, k] = (y[i] - mp.fsum(L[j, i] * iAW[j, k] for j in range(i + 1, n))) / L[i, i]
    WiAW = W.T * iAW
    XiAX = WiAW[:p, :p]; XiAz = WiAW[:p, p]; ziAz = WiAW[p, p]
    beta = mp.lu_solve(XiAX, XiAz)
    res = ziAz - 2 * (beta.T * XiAz)[0] + (beta.T * XiAX * beta)[0]
    useReml = case["useReml"]
    cxd = res / (n - p) if useReml else res / n
    lndetXiAX = mp.log(mp.det(XiAX))
    two_pi = 2 * mp.pi
    if useReml:    # fitIAK3D.R:865-876
        nll = ((n - p) * mp.log(two_pi) + lndetA + n * mp.log(cxd) + lndetXiAX - p * mp.log(cxd) + res / cxd) / 2
    else:          # :878
        nll = (n * mp.log(two_pi) + lndetA + n * mp.log(cxd) + res / cxd) / 2
    s2 = []
    for i in range(n):   # sigma2Vec, fitIAK3D.R:1083-1086: the variance of a POINT inside interval i
        P = mod.P
        v = P["cxd0"] * mod.avvar("cxd0", dI[i]) + P["cxd1"] * mod.avvar("cxd1", dI[i]) + P["cme"] + P["cx0"] + P["cx1"]
        if mod.has["cd1"]:
            v += P["cd1"] * mod.avvar("cd1", dI[i])
        s2.append(v)
    out = dict(A_lower=[[float(A[i, j]) for j in range(i + 1)] for i in range(n)], sigma2Vec=[float(v) for v in s2],
               lndetA=float(lndetA), WiAW=[[float(WiAW[i, j]) for j in range(p + 1)] for i in range(p + 1)], nll=float(nll),
               cxdhat=float(cxd), betahat=[float(b) for b in beta])
    if case.get("predict"):
        # universal kriging from C = cxdhat * A (fitIAK3D.R:857-871: every variance parameter is rescaled by cxdhat)
        xm = [(M(a), M(b)) for a, b in case["xMap"]]
        dm = [(M(a), M(b)) for a, b in case["dIMap"]]
        Xm = [[M(v) for v in row] for row in case["XMap"]]
        K = mp.zeros(n + p, n + p)
        for i in range(n):
            for j in range(n):
                K[i, j] = cxd * A[i, j]
            for k in range(p):
                K[i, n + k] = X[i, k]; K[n + k, i] = X[i, k]
        zs, vs = [], []
        for s in range(len(xm)):
            ck = [cxd * mod.cov(xm[s], dm[s], x[j], dI[j], False, False) for j in range(n)]     # no cme (fitIAK3D.R:1243)
            ckk = cxd * mod.cov(xm[s], dm[s], xm[s], dm[s], True, True)                        # with cme (predictIAK3D.R:189)
            rhs = mp.matrix(ck + Xm[s])
            sol = mp.lu_solve(K, rhs)
            lam = sol[:n]
            zhat = mp.fsum(lam[j] * z[j] for j in range(n))
            lc = mp.fsum(lam[j] * ck[j] for j in range(n))
            lCl = mp.fsum(lam[i] * mp.fsum(K[i, j] * lam[j] for j in range(n)) for i in range(n))
            zs.append(float(zhat)); vs.append(float(ckk - 2 * lc + lCl))
        out.update(zMap=zs, vMap=vs)
    return out


# ---------------------------------------------------------------------------------------------------------------------
# seeded inputs (stored in the fixture, so the tests read them from there)
# ---------------------------------------------------------------------------------------------------------------------
def make_inputs(seed, nprof, p):
    rng = np.random.default_rng(seed)
    xs, ds = [], []
    for k in range(nprof):
        xy = np.round(rng.uniform(0.0, 30.0, 2), 3)
        top = 0.0
        for h in range(int(rng.integers(2, 5))):
            t = float(rng.choice([0.1, 0.2, 0.3]))
            bot = round(top + t, 2)
            xs.append(xy); ds.append((round(top, 2), bot)); top = bot
    xs = np.array(xs); ds = np.array(ds)
    n = len(ds)
    mid = ds.mean(1)
    X = np.column_stack([np.ones(n), np.round(rng.standard_normal(n), 6), np.round(mid - 0.4, 6)])[:, :p]
    z = np.round(20 + X[:, 1:] @ np.array([1.5, -4.0])[:p - 1] + 2.0 * np.cos(xs[:, 0] / 7.0) + rng.standard_normal(n), 6)
    return xs, ds, X, z


def make_map(xs, ds, seed, p):
    rng = np.random.default_rng(seed + 5)
    xm = [xs[0].tolist(), np.round(rng.uniform(0, 30, 2), 3).tolist(), np.round(rng.uniform(0, 30, 2), 3).tolist(),
          xs[len(xs) // 2].tolist(), [200.0, 200.0], np.round(xs[3] + 0.05, 3).tolist()]
    dm = [ds[0].tolist(), [0.05, 0.15], [0.3, 0.6], [0.0, 0.05], [0.15, 0.3], [0.6, 1.0]]
    Xm = [[1.0, float(np.round(rng.standard_normal(), 6)), float(np.round(0.5 * (d[0] + d[1]) - 0.4, 6))][:p] for d in dm]
    return xm, dm, Xm


CASES = [
    dict(name="A-matern-bessel-stationary-reml", seed=7101, nprof=20, p=3, modelx="matern", nud=0.5, types=[0, 0, 0], useReml=True, predict=True,
         pars=dict(ax=9.0, nux=0.8, ad=0.35, cx0=0.10, cx1=0.25, cd1=0.15, cxd0=0.4, cxd1=0.6, cme=0.05),
         tau=dict(cd1=None, cxd0=None, cxd1=None)),
    dict(name="B-spherical-nocd1-nonstat-ml", seed=7102, nprof=19, p=3, modelx="spherical", nud=1.5, types=[-9, -1, -1], useReml=False, predict=True,
         pars=dict(ax=14.0, nux=float("nan"), ad=0.5, cx0=0.2, cx1=0.3, cd1=0.0, cxd0=0.35, cxd1=0.65, cme=0.08),
         tau=dict(cd1=None, cxd0=[0.6, 1.7], cxd1=[0.3, 0.9])),
    dict(name="C-matern32-cd1quirk-nud25-reml", seed=7103, nprof=18, p=3, modelx="matern", nud=2.5, types=[-1, 0, -1], useReml=True, predict=False,
         cd1_quirk=True,
         pars=dict(ax=6.0, nux=1.5, ad=0.3, cx0=0.05, cx1=0.2, cd1=0.15, cxd0=0.45, cxd1=0.55, cme=0.03),
         tau=dict(cd1=[0.5, 2.2], cxd0=None, cxd1=[0.6, 1.7])),
]


def main():
    # convergence of the quadrature: 20 against 28 nodes per piece on the hardest shapes
    worst = mp.mpf(0)
    for I, J in (((M(0.0), M(0.3)), (M(0.1), M(0.2))), ((M(0.0), M(0.1)), (M(0.0), M(0.1))), ((M(0.2), M(0.5)), (M(0.4), M(1.0)))):
        for nud in (0.5, 1.5, 2.5):
            a = phi_d(I, J, M(0.3), nud, (M(0.6), M(1.7)), 20); b = phi_d(I, J, M(0.3), nud, (M(0.6), M(1.7)), 28)
            worst = max(worst, abs(a - b) / abs(b))
    print("quadrature 20 vs 28 nodes, max relative difference:", mp.nstr(worst, 3), file=sys.stderr)
    assert worst < mp.mpf(10) ** -25
    out = dict(source="scripts/make_mpmath_anchors.py: 40-digit mpmath evaluation of the model's DEFINITIONS (quadrature of the defining "
                      "double integral, besselk, dense Cholesky, textbook universal-kriging system); no repository code involved",
               quadrature_check=float(worst), cases=[])
    for c in CASES:
        xs, ds, X, z = make_inputs(c["seed"], c["nprof"], c["p"])
        case = dict(c)
        case.update(n=len(z), xData=xs.tolist(), dIData=ds.tolist(), XData=X.tolist(), zData=z.tolist())
        if c["predict"]:
            xm, dm, Xm = make_map(xs, ds, c["seed"], c["p"])
            case.update(xMap=xm, dIMap=dm, XMap=Xm)
        res = build_case(case)
        case.update(res)
        case["pars"] = {k: (None if isinstance(v, float) and np.isnan(v) else v) for k, v in c["pars"].items()}
        out["cases"].append(case)
        print(f"\n{c['name']}: n = {case['n']}, nll = {res['nll']:.12f}, cxdhat = {res['cxdhat']:.12f}", file=sys.stderr)
    with open(os.path.join(ROOT, "tests", "golden", "mpmath_anchors.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", len(out["cases"]), "cases", file=sys.stderr)


if __name__ == "__main__":
    main()
