// tail.cu -- the closed-form end of a likelihood evaluation (fitIAK3D.R:821-880; compositeLikIAK3D.R:191-265) on the host:
// betahat = solve(XiAX, XiAz), the profiled variance and the REML / ML value from lndetA and WiAW = t([X z]) A^-1 [X z].
// O(p^3) with p <= a few dozen: it belongs to the host, but not to the interpreter -- at Edgeroi size (n = 1200) a NumPy / R
// tail costs 0.2 ms per evaluation against 0.65 ms of device time, and the optimiser calls it hundreds of times in sequence.
#include <cmath>
#include <cstring>
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

// ================================================================================================================================
// readParsIAK3D (fitIAK3D.R:1332-1417; logitab :1587, tauTfm :1618, sdfdRead :1474) and the calls an optimiser makes, natively:
// the unconstrained vector goes in, the value (or the forward-difference gradient) comes out -- one library call per callback.
// ================================================================================================================================
namespace {
inline double logitab_inv(double z, const double* b) { return (1.0 / (1.0 + std::exp(-z))) * (b[1] - b[0]) + b[0]; }   // exp overflow -> +Inf -> lower bound, as in R
}

extern "C" int iak3d_read_pars(const iak3d_model* m, const double* pars, int32_t npar, iak3d_pars* out) {
  if (!m || !pars || !out || npar < 0) return IAK3D_ERR_ARG;
  if (m->modelx < 0 || m->modelx > 3) return IAK3D_ERR_ARG;
  std::memset(out, 0, sizeof(*out));
  int k = 0;
  bool short_ = false;
  auto take = [&]() -> double { if (k >= npar) { short_ = true; return NAN; } return pars[k++]; };
  auto ev = [&]() -> double { return 1e-8 + std::exp(take()); };               // variances: 1e-8 + exp(.) (:1353-1364)
  const bool nugget = m->modelx == 0, spherical = m->modelx == 1;
  out->modelx = m->modelx; out->cmeOpt = m->cmeOpt; out->nud = m->nud;
  out->quirk_cd1_unscaled = 1; out->quirk_scale = 1.0;
  out->ax = NAN; out->nux = NAN;
  if (!nugget) {
    out->ax = logitab_inv(take(), m->axBnds);
    if (!spherical) out->nux = logitab_inv(take(), m->nuxBnds);
  }
  out->ad = logitab_inv(take(), m->adBnds);
  if (m->prodSum) {
    out->cx0 = ev();
    if (!nugget) out->cx1 = ev();
  }
  out->cd1 = (m->sdfdType[0] != -9) ? ev() : 0.0;
  if (!m->lnTfmdData) {
    if (nugget) { out->cxd0 = 1.0; out->cxd1 = 0.0; }
    else { out->cxd1 = logitab_inv(take(), m->sxd1Bnds); out->cxd0 = 1.0 - out->cxd1; }
  } else {
    out->cxd0 = ev();
    out->cxd1 = nugget ? 0.0 : ev();
  }
  for (int c = 0; c < 3; c++) {
    const int t = m->sdfdType[c];
    out->sdfdType[c] = t;
    if (t == -1) {                                                           // tauTfm(invt = TRUE), :1618-1635
      const double v0 = take(), v1 = take();
      const double t2 = logitab_inv(v1, m->tau2Bnds);
      out->sdfdPars[c][0] = (1.0 - std::exp(v0)) / (1.0 - std::exp(-t2 * m->maxd));
      out->sdfdPars[c][1] = t2;
    } else if (t > 0) {
      if (t > IAK3D_MAX_SPL) return IAK3D_ERR_ARG;
      for (int e = 0; e < t; e++) out->sdfdSpl[c][e] = take();
    } else if (t != 0 && t != -9) {
      return IAK3D_ERR_ARG;                                                  // "sdfd option not yet programmed"
    }
  }
  out->cme = (m->cmeOpt == 1) ? std::exp(take()) : 0.0;
  if (short_) return IAK3D_ERR_ARG;
  const int rest = npar - k;                                                 // site-specific random effects: exp of what is left (:1413)
  if (rest > IAK3D_MAX_SSRE) return IAK3D_ERR_ARG;
  out->nssre = rest;
  for (int e = 0; e < rest; e++) out->cssre[e] = std::exp(pars[k + e]);
  return IAK3D_OK;
}

extern "C" int iak3d_nll_vector(iak3d_ds* ds, const iak3d_model* m, const double* pars, int32_t npar, int32_t useReml, double* nll) {
  if (!ds || !nll) return IAK3D_ERR_ARG;
  iak3d_pars P;
  const int st = iak3d_read_pars(m, pars, npar, &P);
  if (st != 0) { ds->ctx->err = "nll_vector: the parameter vector does not fit the model"; return st; }
  int32_t status = 0;
  *nll = 9e99;
  const int rc = iak3d_nll_reml(ds, &P, useReml, nll, nullptr, nullptr, nullptr, nullptr, &status);
  if (rc < 0) return rc;
  if (status != 0) *nll = 9e99;
  return IAK3D_OK;
}

// gradnllIAK3D (fitIAK3D.R:1776-1911): npar + 1 evaluations in ONE batched launch; grad[k] = (nll(pars + delta e_k) - nll0) / delta,
// a failed candidate counts as 9e99 and a non-finite difference as 0 (:1812-1813).
extern "C" int iak3d_gradnll_fd(iak3d_ds* ds, const iak3d_model* m, const double* pars, int32_t npar, double delta, int32_t useReml,
                                double* grad, double* nll0) {
  if (!ds || !m || !pars || !grad || npar < 1) return IAK3D_ERR_ARG;
  iak3d_ctx* ctx = ds->ctx;
  const int q = ds->q, p = q - 1, cnt = npar + 1;
  if (q < 1) { ctx->err = "gradnll_fd: the dataset has no W = [X z]"; return IAK3D_ERR_ARG; }
  std::vector<iak3d_pars> P((size_t)cnt);
  std::vector<double> v(pars, pars + npar);
  for (int c = 0; c < cnt; c++) {
    if (c > 0) v[c - 1] = pars[c - 1] + delta;
    const int st = iak3d_read_pars(m, v.data(), npar, &P[c]);
    if (c > 0) v[c - 1] = pars[c - 1];
    if (st != 0) { ctx->err = "gradnll_fd: the parameter vector does not fit the model"; return st; }
  }
  std::vector<iak3d_ds*> dss((size_t)cnt, ds);
  std::vector<double> lnd((size_t)cnt), G((size_t)cnt * q * q), val((size_t)cnt);
  std::vector<int32_t> st((size_t)cnt);
  const int rc = iak3d_nll_terms_batch(ctx, cnt, dss.data(), P.data(), lnd.data(), G.data(), st.data());
  if (rc < 0) return rc;
  for (int c = 0; c < cnt; c++) {
    val[c] = 9e99;
    if (st[c] != 0) continue;
    int32_t ok = 0;
    iak3d_reml_tail(G.data() + (size_t)c * q * q, lnd[c], ds->n, p, useReml, 0.0, 0.0, &val[c], nullptr, nullptr, nullptr, nullptr, &ok);
  }
  for (int kk = 0; kk < npar; kk++) {
    const double g = (val[kk + 1] - val[0]) / delta;
    grad[kk] = std::isfinite(g) ? g : 0.0;
  }
  if (nll0) *nll0 = val[0];
  return IAK3D_OK;
}
