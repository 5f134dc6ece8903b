// iak3d_math.h -- scalar FP64 kernels of the iak3d hot path, written once for host and device.
//
// Everything here is a pure function of doubles so that the SAME source is (a) inlined into the
// sm_100a table kernels (tables.cu) and (b) compiled for the host by gcc into the test-only
// library tests/_hostmath.so, which lets the CPU test-suite check the device arithmetic against
// the oracle without a GPU.  The product never calls the host build.
//
// Reference semantics restated (file:line in the reference):
//   depth kernel   iaCovMatern.R:4-102,107-191,196-322
//   horizontal     iaCovMatern.R:327-408 (spherical, Matern via besselK/lgamma, Wendland :1357-1473)
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define IAK_HD __host__ __device__ __forceinline__
#else
#define IAK_HD static inline
#endif

// ---------------------------------------------------------------------------------------------
// Depth kernel: increment-averaged Matern (nu = 1/2, 3/2, 5/2) with sd(d) = tau0 + tau1 exp(-tau2 d)
// ---------------------------------------------------------------------------------------------
struct IaPrm {
  // point model rho(h) = (r0 + r1 h + r2 h^2) exp(-psi h)            iaCovMatern.R:12-29
  double psi, pt;                 // pt = clamped psi - tau2            iaCovMatern.R:231-237
  double tau0, tau1, tau2;
  double g11[3], g12[3];          // gammafnInt chains from gamma1 = gamma(r, psi)        :229-239
  double h11[3], h12[3];          // the same from gamma2 = gamma(r, psi + tau2)          :251-262
  double g1_0, g2_0;              // gamma1[1], gamma2[1] (R 1-based) used by intTy       :305,314
  double cA, cB, cC, cD;          // -(t0^2/psi^2), -(t0 t1/(psi pt)), -(t0 t1/((psi+t2) psi)), -(t1^2/((psi+t2) pt))
  double cE, cF, cG, cH;          // intTy2/intTy4 coefficients                            :305-317
  int stationary;                 // tau1 == 0: every tau1 term is an exact zero and is skipped
};

IAK_HD void iak_gammafn(const double* al, double a, double* g) {   // iaCovMatern.R:196-208
  g[0] = -(al[0] + al[1] / a + 2.0 * al[2] / (a * a));
  g[1] = -(al[1] + 2.0 * al[2] / a);
  g[2] = -(al[2]);
}

// returns 0 on success, 2 on invalid (nud not in {0.5,1.5,2.5} -> the reference stop()s, :27-29)
IAK_HD int iak_prm_init(IaPrm* P, double ad, double nud, int sdfdType, double s1, double s2) {
  double r[3];
  if (nud == 0.5) { r[0] = 1.0; r[1] = 0.0; r[2] = 0.0; P->psi = 1.0 / ad; }
  else if (nud == 1.5) { r[0] = 1.0; r[1] = sqrt(3.0) / ad; r[2] = 0.0; P->psi = sqrt(3.0) / ad; }
  else if (nud == 2.5) { r[0] = 1.0; r[1] = sqrt(5.0) / ad; r[2] = 5.0 / (3.0 * (ad * ad)); P->psi = sqrt(5.0) / ad; }
  else return 2;
  if (sdfdType == 0) { P->tau0 = 1.0; P->tau1 = 0.0; P->tau2 = 1.0; }              // :36-39
  else if (sdfdType == -1) { P->tau0 = 1.0 - s1; P->tau1 = s1; P->tau2 = s2; }     // :40-43
  else return 2;
  const double psi = P->psi, t0 = P->tau0, t1 = P->tau1, t2 = P->tau2;
  double pt = psi - t2;
  if (pt >= 0.0 && pt < 1e-3) pt = 1e-3; else if (pt < 0.0 && pt > -1e-3) pt = -1e-3;
  P->pt = pt;
  double g1[3], g2[3];
  iak_gammafn(r, psi, g1); iak_gammafn(g1, psi, P->g11); iak_gammafn(g1, pt, P->g12);
  iak_gammafn(r, psi + t2, g2); iak_gammafn(g2, psi, P->h11); iak_gammafn(g2, pt, P->h12);
  P->g1_0 = g1[0]; P->g2_0 = g2[0];
  P->cA = -((t0 * t0) / (psi * psi));
  P->cB = -((t0 * t1) / (psi * pt));
  P->cC = -((t0 * t1) / ((psi + t2) * psi));
  P->cD = -((t1 * t1) / ((psi + t2) * pt));
  P->cE = ((t0 * t0) / psi) * g1[0];                                 // intTy2 first term  (* (bb-aa))
  P->cF = ((t0 * t1) / (psi * t2)) * g1[0];                          // intTy2 second term
  P->cG = -((t0 * t1) / (t2 * (psi + t2))) * g2[0];                  // intTy4 first term
  P->cH = ((t1 * t1) / (2.0 * (psi + t2) * t2)) * g2[0];             // intTy4 second term
  P->stationary = (t1 == 0.0);
  return 0;
}

IAK_HD double iak_p2e(double w, const double* al, double psi) {     // poly2Exp, iaCovMatern.R:210-213
  return (al[0] + al[1] * w + al[2] * (w * w)) * exp(-psi * w);
}

IAK_HD double iak_intRy12(double ee, double aa, double bb, const IaPrm* P) {   // :227-247
  double v = P->cA * (iak_p2e(ee - bb, P->g11, P->psi) - iak_p2e(ee - aa, P->g11, P->psi));
  if (!P->stationary)
    v = v + P->cB * exp(-P->tau2 * ee) * (iak_p2e(ee - bb, P->g12, P->pt) - iak_p2e(ee - aa, P->g12, P->pt));
  return v;
}
IAK_HD double iak_intRy34(double ee, double aa, double bb, const IaPrm* P) {   // :249-270
  return P->cC * exp(-P->tau2 * ee) * (iak_p2e(ee - bb, P->h11, P->psi) - iak_p2e(ee - aa, P->h11, P->psi))
       + P->cD * exp(-2.0 * P->tau2 * ee) * (iak_p2e(ee - bb, P->h12, P->pt) - iak_p2e(ee - aa, P->h12, P->pt));
}
IAK_HD double iak_intRy(double aa, double bb, double cc, double dd, const IaPrm* P) {   // :219-225
  double v = iak_intRy12(dd, aa, bb, P) - iak_intRy12(cc, aa, bb, P);
  if (!P->stationary) v = v + iak_intRy34(dd, aa, bb, P) - iak_intRy34(cc, aa, bb, P);
  return v;
}
IAK_HD double iak_intTy(double aa, double bb, const IaPrm* P) {     // :276-322
  const double w = bb - aa;
  double i1 = P->cA * (P->g11[0] - iak_p2e(w, P->g11, P->psi));
  double i2 = P->cE * w;
  if (P->stationary) return i1 - i2;
  const double eb = exp(-P->tau2 * bb), ea = exp(-P->tau2 * aa);
  const double e2b = exp(-P->tau2 * (2.0 * bb)), e2a = exp(-P->tau2 * (2.0 * aa));
  i1 = i1 + P->cB * eb * (P->g12[0] - iak_p2e(w, P->g12, P->pt));
  i2 = i2 - P->cF * (eb - ea);
  double i3 = P->cC * eb * (P->h11[0] - iak_p2e(w, P->h11, P->psi))
            + P->cD * e2b * (P->h12[0] - iak_p2e(w, P->h12, P->pt));
  double i4 = P->cG * (eb - ea) - P->cH * (e2b - e2a);
  return i1 - i2 + i3 - i4;
}

// avCov of intervals (a,b) and (c,d): swap so that a <= c (first index wins ties, :68-70), then the
// three region cases (:76-89).
IAK_HD double iak_avcov(double a, double b, double c, double d, const IaPrm* P) {
  if (c < a) { double t = a; a = c; c = t; t = b; b = d; d = t; }
  double v;
  if (b <= c) v = iak_intRy(a, b, c, d, P);
  else if (b <= d) v = iak_intRy(a, c, c, d, P) + iak_intRy(c, b, b, d, P) + 2.0 * iak_intTy(c, b, P);
  else v = iak_intRy(a, c, c, d, P) + iak_intRy(c, d, d, b, P) + 2.0 * iak_intTy(c, d, P);
  return v / ((d - c) * (b - a));
}

IAK_HD double iak_avvar(double a, double b, const IaPrm* P) {       // iaCovMatern.R:53-54
  const double t0 = P->tau0, t1 = P->tau1, t2 = P->tau2;
  double v = (t0 * t0) * (b - a) - 2.0 * (t0 * t1 / t2) * (exp(-t2 * b) - exp(-t2 * a))
           - 0.5 * ((t1 * t1) / t2) * (exp(-2.0 * t2 * b) - exp(-2.0 * t2 * a));
  return v / (b - a);
}

// ---------------------------------------------------------------------------------------------
// Modified Bessel function of the second kind K_nu(x), nu >= 0, x > 0.
// Temme's series (x <= 2) / Steed's continued fraction CF2 (x > 2) for |mu| <= 1/2, then the stable
// upward recurrence in the order.  R's besselK (Cody's rkbesl) uses the same two regimes; both are
// accurate to a few ulp.  Unscaled: underflows to 0 beyond x ~ 705 like R's non-expon.scaled call.
// ---------------------------------------------------------------------------------------------
IAK_HD double iak_chebev(const double* c, int m, double x) {        // Clenshaw on [-1,1]
  double d = 0.0, dd = 0.0, sv, y2 = 2.0 * x;
  for (int j = m - 1; j >= 1; j--) { sv = d; d = y2 * d - dd + c[j]; dd = sv; }
  return x * d - dd + 0.5 * c[0];
}

// Optional per-order reciprocal table `rt` (device: built once per thread block in shared memory by
// iak_besselk_table): the divisors of both iterations depend only on the order, not on x, so the table
// removes every division from the Temme loop and two of the three from the Steed loop.
//   rt[4*i + 0] = 1/(i*i - mu^2)   rt[4*i + 1] = 1/(i - mu)   rt[4*i + 2] = 1/(i + mu)        (Temme, i >= 1)
//   rt[4*i + 3] = 1/a_i with a_i = -(1/4 - mu^2) - i(i-1)                                      (Steed, i >= 2)
#define IAK_BK_NRT 48
IAK_HD void iak_besselk_table(double nu, double* rt, int i) {       // entry i of the table (1 <= i < IAK_BK_NRT)
  const int nl = (int)(nu + 0.5);
  const double xmu = nu - nl, xmu2 = xmu * xmu;
  rt[4 * i + 0] = 1.0 / (i * (double)i - xmu2);
  rt[4 * i + 1] = 1.0 / (i - xmu);
  rt[4 * i + 2] = 1.0 / (i + xmu);
  rt[4 * i + 3] = 1.0 / (-(0.25 - xmu2) - (double)i * (i - 1));
}

// Steed's continued fraction CF2 for x >= 2, WITHOUT the sqrt(pi/(2x)) e^-x factor:
//   K_mu(x) = sqrt(pi/(2x)) e^-x * hm ,   K_{mu+1}(x) = sqrt(pi/(2x)) e^-x * h1        (|mu| <= 1/2)
IAK_HD void iak_bk_cf2_scaled(double xmu, double x, const double* rt, double* hm, double* h1) {
  const double EPS = 1e-16;
  const double xmu2 = xmu * xmu, xi = 1.0 / x;
  double b = 2.0 * (1.0 + x), d = 1.0 / b, h = d, delh = d, q1 = 0.0, q2 = 1.0;
  const double a1 = 0.25 - xmu2;
  double q = a1, c = a1, a = -a1, s = 1.0 + q * delh;
  for (int i = 2; i <= 10000; i++) {
    a -= 2 * (i - 1);
    double qnew;
    if (rt != nullptr && i < IAK_BK_NRT) {
      c = -a * c * (1.0 / i);
      qnew = (q1 - b * q2) * rt[4 * i + 3];
    } else {
      c = -a * c / i;
      qnew = (q1 - b * q2) / a;
    }
    q1 = q2; q2 = qnew;
    q += c * qnew;
    b += 2.0;
    d = 1.0 / (b + a * d);
    delh = (b * d - 1.0) * delh;
    h += delh;
    const double dels = q * delh;
    s += dels;
    if (fabs(dels / s) < EPS) break;
  }
  h = a1 * h;
  *hm = 1.0 / s;
  *h1 = *hm * (xmu + x + 0.5 - h) * xi;
}

// A table of K_nu on the unique location pairs is one order nu and ~10^6 arguments: the order-dependent work is done ONCE.
// hm(u), h1(u) above are smooth functions of u = 1/x on [0, 1/2] (they tend to 1 as u -> 0); on each of the four octaves
// [0,1/16], [1/16,1/8], [1/8,1/4], [1/4,1/2] a Chebyshev interpolant of 12 terms reproduces them to 2e-15 relative (checked
// against 30-digit values for mu in [-1/2, 3/2]; tests/test_device_math_on_host.py compares the whole K_nu with the
// continued fraction over nu in [0.05, 20]).  The 96 coefficients are made on the host by evaluating the continued fraction at
// the Chebyshev nodes (hx_init) and travel in HxPrm; per argument the device then runs two 12-term Clenshaw sums instead of
// 15-40 iterations of the continued fraction with a division each (the phi_x tables were 3.8 % of the S5 step).
#define IAK_BK_NSEG 4
#define IAK_BK_NCH 12
struct BkFit { double c[2][IAK_BK_NSEG][IAK_BK_NCH]; int on; };

IAK_HD void iak_bk_fit_init(BkFit* F, double nu) {
  const double PI = 3.141592653589793238462643;
  const int nl = (int)(nu + 0.5);
  const double xmu = nu - nl;
  const double edge[IAK_BK_NSEG + 1] = {0.0, 0.0625, 0.125, 0.25, 0.5};
  for (int sgm = 0; sgm < IAK_BK_NSEG; sgm++) {
    double f[2][IAK_BK_NCH];
    const double mid = 0.5 * (edge[sgm] + edge[sgm + 1]), half = 0.5 * (edge[sgm + 1] - edge[sgm]);
    for (int k = 0; k < IAK_BK_NCH; k++) {
      const double u = mid + half * cos(PI * (k + 0.5) / IAK_BK_NCH);
      iak_bk_cf2_scaled(xmu, 1.0 / u, nullptr, &f[0][k], &f[1][k]);
    }
    for (int w = 0; w < 2; w++)
      for (int j = 0; j < IAK_BK_NCH; j++) {
        double acc = 0.0;
        for (int k = 0; k < IAK_BK_NCH; k++) acc += f[w][k] * cos(PI * j * (k + 0.5) / IAK_BK_NCH);
        F->c[w][sgm][j] = 2.0 * acc / IAK_BK_NCH;
      }
  }
  F->on = 1;
}

IAK_HD double iak_besselk(double nu, double x, const double* rt, const BkFit* fit = nullptr) {
  const double EPS = 1e-16, PI = 3.141592653589793238462643;
  if (!(x > 0.0)) return INFINITY;
  if (x > 705.342) return 0.0;                                       // R nmath xmax_BESS_K
  const int nl = (int)(nu + 0.5);
  const double xmu = nu - nl, xmu2 = xmu * xmu, xi = 1.0 / x, xi2 = 2.0 * xi;
  double rkmu, rk1;
  if (x < 2.0) {
    const double c1[7] = {-1.142022680371168e0, 6.5165112670737e-3, 3.087090173086e-4, -3.4706269649e-6,
                          6.9437664e-9, 3.67795e-11, -1.356e-13};
    const double c2[8] = {1.843740587300905e0, -7.68528408447867e-2, 1.2719271366546e-3, -4.9717367042e-6,
                          -3.31261198e-8, 2.423096e-10, -1.702e-13, -1.49e-15};
    const double xx = 8.0 * xmu * xmu - 1.0;
    const double gam1 = iak_chebev(c1, 7, xx), gam2 = iak_chebev(c2, 8, xx);
    const double gampl = gam2 - xmu * gam1, gammi = gam2 + xmu * gam1;   // 1/Gamma(1+mu), 1/Gamma(1-mu)
    double d = -log(0.5 * x), e = xmu * d;
    const double fact2 = (fabs(e) < EPS ? 1.0 : sinh(e) / e);
    const double pimu = PI * xmu;
    const double fact = (fabs(pimu) < EPS ? 1.0 : pimu / sin(pimu));
    double ff = fact * (gam1 * cosh(e) + gam2 * fact2 * d);
    double sum = ff;
    e = exp(e);
    double p = 0.5 * e / gampl, q = 0.5 / (e * gammi), c = 1.0;
    d = 0.25 * x * x;
    double sum1 = p;
    for (int i = 1; i <= 10000; i++) {
      if (rt != nullptr && i < IAK_BK_NRT) {
        ff = (i * ff + p + q) * rt[4 * i + 0];
        c *= d * (1.0 / i);
        p *= rt[4 * i + 1];
        q *= rt[4 * i + 2];
      } else {
        ff = (i * ff + p + q) / (i * (double)i - xmu2);
        c *= (d / i);
        p /= (i - xmu);
        q /= (i + xmu);
      }
      const double del = c * ff;
      sum += del;
      sum1 += c * (p - i * ff);
      if (fabs(del) < fabs(sum) * EPS) break;
    }
    rkmu = sum; rk1 = sum1 * xi2;
  } else {
    double hm, h1;
    if (fit != nullptr && fit->on) {
      const int sgm = xi > 0.25 ? 3 : (xi > 0.125 ? 2 : (xi > 0.0625 ? 1 : 0));
      const double sc = sgm == 3 ? 8.0 : (sgm == 2 ? 16.0 : 32.0), of = sgm == 0 ? -1.0 : -3.0;
      const double t = xi * sc + of;                                 // the octave mapped to [-1, 1] (exact constants)
      hm = iak_chebev(fit->c[0][sgm], IAK_BK_NCH, t);
      h1 = iak_chebev(fit->c[1][sgm], IAK_BK_NCH, t);
    } else {
      iak_bk_cf2_scaled(xmu, x, rt, &hm, &h1);
    }
    const double pref = sqrt(PI / (2.0 * x)) * exp(-x);
    rkmu = pref * hm;
    rk1 = pref * h1;
  }
  for (int i = 1; i <= nl; i++) { const double t = (xmu + i) * xi2 * rk1 + rkmu; rkmu = rk1; rk1 = t; }
  return rkmu;
}

// ---------------------------------------------------------------------------------------------
// Horizontal correlation phi_x(D) with c1 = 1 (spatialCovIAK3D, iaCovMatern.R:327-408)
// ---------------------------------------------------------------------------------------------
struct HxPrm {
  int modelx;            // 1 spherical, 2 matern, 3 wendland
  double a, nu;
  double s;              // sqrt(2 nu) / a                                  :358
  double lconst;         // -(nu - 1) log 2 - lgamma(nu)                    :363
  int half;              // 1,3,5 when nu is 0.5/1.5/2.5 (closed form), else 0
  // Wendland: blend of floor(nu) and ceil(nu) members (:384-386); beta coefficients per member
  int kf, kc; double wf, wc;
  double betaf[24], betac[24]; double scf, scc;
  BkFit bk;              // Matern with a general order: Chebyshev fit of the scaled K_mu, K_{mu+1} for x >= 2 (hx_init)
};

IAK_HD double iak_wend_eval(double r, int k, const double* beta) {  // wendland.eval, :1381-1390 (n = 2)
  const int l = 1 + k + 1;                                            // floor(2/2) + k + 1
  double phi = beta[0] * pow(1.0 - r, (double)(l + 2 * k));
  for (int m = 1; m <= k; m++) phi = phi + beta[m] * pow(r, (double)m) * pow(1.0 - r, (double)(l + 2 * k - m));
  return phi;
}
IAK_HD double iak_wendland(double D, double theta, int k, const double* beta, double sc) {   // :1358-1379
  double d = D;
  if (theta != 1.0) d = d / theta;
  if (k == 2) return (d < 1.0) ? (pow(1.0 - d, 6.0) * (35.0 * (d * d) + 18.0 * d + 3.0)) / 3.0 : 0.0;
  return (d < 1.0) ? iak_wend_eval(d, k, beta) / sc : 0.0;
}

IAK_HD double iak_phix(double D, const HxPrm* H, const double* rt = nullptr) {
  if (H->modelx == 1) {                                              // spherical :336-348
    const double r = D / H->a;
    if (r > 1.0) return 0.0;
    // R evaluates DOVERa^3 with libm pow() (correctly rounded); r*r*r would carry two roundings, and the
    // expression cancels near the range.  p + e == r*r exactly, so r3 is r^3 rounded (almost always) once.
    const double p = r * r, e = fma(r, r, -p);
    const double r3 = fma(p, r, e * r);
    return 1.0 - (1.5 * r - 0.5 * r3);
  }
  if (H->modelx == 2) {                                              // matern :353-377
    if (D == 0.0) return 1.0;
    const double x = H->s * D;
    if (H->half == 1) return exp(-x);
    if (H->half == 3) return (1.0 + x) * exp(-x);
    if (H->half == 5) return (1.0 + x + (x * x) / 3.0) * exp(-x);
    const double K = iak_besselk(H->nu, x, rt, &H->bk);
    const double v = exp(H->nu * log(x) + H->lconst + log(K));
    if (isinf(v)) return 1.0;                                        // :364 is.infinite -> c1
    return v;
  }
  if (H->modelx == 3) {                                              // wendland :378-401
    return H->wf * iak_wendland(D, H->a, H->kf, H->betaf, H->scf) + H->wc * iak_wendland(D, H->a, H->kc, H->betac, H->scc);
  }
  return 0.0;
}
