// tail.cu -- the closed-form end of a likelihood evaluation (fitIAK3D.R:821-880; compositeLikIAK3D.R:191-265) on the host:
// betahat = solve(XiAX, XiAz), the profiled variance and the REML / ML value from lndetA and WiAW = t([X z]) A^-1 [X z].
// O(p^3) with p <= a few dozen: it belongs to the host, but not to the interpreter -- at Edgeroi size (n = 1200) a NumPy / R
// tail costs 0.2 ms per evaluation against 0.65 ms of device time, and the optimiser calls it hundreds of times in sequence.
#include <cmath>
#include <vector>

#include "common.cuh"

// WiAW: q x q column-major (q = p + 1); only its UPPER triangle is read (forceSymmetric, fitIAK3D.R:811).
// n_for_cxd / n_for_lndet: <= 0 -> the defaults (n - p with REML, n without; n), else the composite likelihood's counts.
// ok = 0 (and nll = 9e99) where XiAX is not positive definite or a term is not finite -- the reference's badnll.
extern "C" int iak3d_reml_tail(const double* WiAW, double lndetA, int64_t n, int32_t p, int32_t useReml, double n_for_cxd, double n_for_lndet,
                               double* nll, double* betahat, double* cxdhat, double* lndetXiAX_out, double* resiAres_out, int32_t* ok) {
  if (!WiAW || !nll || !ok || p < 0) return IAK3D_ERR_ARG;
  const int q = p + 1;
  *nll = 9e99; *ok = 0;
  if (cxdhat) *cxdhat = NAN;
  if (lndetXiAX_out) *lndetXiAX_out = NAN;
  if (resiAres_out) *resiAres_out = NAN;
  auto G = [&](int i, int j) { return i <= j ? WiAW[(size_t)j * q + i] : WiAW[(size_t)i * q + j]; };
  for (int j = 0; j < q; j++)
    for (int i = 0; i <= j; i++)
      if (!std::isfinite(WiAW[(size_t)j * q + i])) return IAK3D_OK;
  if (!std::isfinite(lndetA)) return IAK3D_OK;
  // Cholesky of XiAX (lower, row-major scratch), log-determinant, two substitutions
  std::vector<double> L((size_t)p * p, 0.0), y(p), b(p);
  double lndetXiAX = 0.0;
  for (int j = 0; j < p; j++) {
    double d = G(j, j);
    for (int k = 0; k < j; k++) d -= L[(size_t)j * p + k] * L[(size_t)j * p + k];
    if (!(d > 0.0) || !std::isfinite(d)) return IAK3D_OK;
    const double ljj = std::sqrt(d);
    L[(size_t)j * p + j] = ljj;
    lndetXiAX += std::log(ljj);
    for (int i = j + 1; i < p; i++) {
      double s = G(i, j);
      for (int k = 0; k < j; k++) s -= L[(size_t)i * p + k] * L[(size_t)j * p + k];
      L[(size_t)i * p + j] = s / ljj;
    }
  }
  lndetXiAX *= 2.0;
  for (int i = 0; i < p; i++) {
    double s = G(i, p);
    for (int k = 0; k < i; k++) s -= L[(size_t)i * p + k] * y[k];
    y[i] = s / L[(size_t)i * p + i];
  }
  for (int i = p - 1; i >= 0; i--) {
    double s = y[i];
    for (int k = i + 1; k < p; k++) s -= L[(size_t)k * p + i] * b[k];
    b[i] = s / L[(size_t)i * p + i];
  }
  // resiAres = ziAz - 2 b'XiAz + b'XiAX b   (fitIAK3D.R:836)
  double bXz = 0.0, bXXb = 0.0;
  for (int i = 0; i < p; i++) bXz += b[i] * G(i, p);
  for (int i = 0; i < p; i++) {
    double s = 0.0;
    for (int k = 0; k < p; k++) s += G(i, k) * b[k];
    bXXb += b[i] * s;
  }
  const double res = G(p, p) - 2.0 * bXz + bXXb;
  const double nn = n_for_lndet > 0.0 ? n_for_lndet : (double)n;
  const double c = n_for_cxd > 0.0 ? res / n_for_cxd : (useReml ? res / (double)(n - p) : res / (double)n);
  const double lc = std::log(c);
  const double lndetC = lndetA + nn * lc;
  const double lndetXiCX = lndetXiAX - (double)p * lc;
  const double resiCres = c != 0.0 ? res / c : NAN;
  const double l2pi = std::log(2.0 * 3.14159265358979323846);
  const double v = useReml ? 0.5 * ((double)(n - p) * l2pi + lndetC + lndetXiCX + resiCres) : 0.5 * ((double)n * l2pi + lndetC + resiCres);
  if (betahat) for (int i = 0; i < p; i++) betahat[i] = b[i];
  if (cxdhat) *cxdhat = c;
  if (lndetXiAX_out) *lndetXiAX_out = lndetXiAX;
  if (resiAres_out) *resiAres_out = res;
  *ok = 1;                                   // the p x p system was solved; a non-finite value is still the reference's badnll
  *nll = std::isfinite(v) ? v : 9e99;
  return IAK3D_OK;
}

// nllIAK3D without rtnAll (fitIAK3D.R:730-880) in ONE call: covariance build + factorisation + bordered solves on the device,
// the tail above on the host.  status: IAK3D_OK / IAK3D_NOT_SPD / IAK3D_BAD_PARS like iak3d_nll_terms (nll = 9e99 then).
extern "C" int iak3d_nll_reml(iak3d_ds* ds, const iak3d_pars* pars, int32_t useReml, double* nll, double* betahat, double* cxdhat,
                              double* lndetA_out, double* WiAW_out, int32_t* status) {
  if (!ds || !pars || !nll || !status) return IAK3D_ERR_ARG;
  const int q = ds->q, p = q - 1;
  if (q < 1) { ds->ctx->err = "nll_reml: the dataset has no W = [X z]"; return IAK3D_ERR_ARG; }
  std::vector<double> G((size_t)q * q);
  double lnd = NAN;
  *nll = 9e99;
  if (cxdhat) *cxdhat = NAN;
  const int st = iak3d_nll_terms(ds, pars, &lnd, G.data(), nullptr);
  if (st < 0) return st;
  *status = st;
  if (st != 0) return IAK3D_OK;
  if (lndetA_out) *lndetA_out = lnd;
  if (WiAW_out) for (size_t e = 0; e < G.size(); e++) WiAW_out[e] = G[e];
  int32_t ok = 0;
  return iak3d_reml_tail(G.data(), lnd, ds->n, p, useReml, 0.0, 0.0, nll, betahat, cxdhat, nullptr, nullptr, &ok);
}
