/* r_shim.c -- .Call entry points that bind libiak3d.so (include/iak3d.h) into R.
 *
 * Build where R is installed (it is not in the image this repository is developed in):
 *     R CMD SHLIB -o iak3dR.so r/r_shim.c -I include -L iak3d_b200/lib -liak3d -Wl,-rpath,$PWD/iak3d_b200/lib
 * then  dyn.load("iak3dR.so"); source("r/iak3d_cuda.R")  AFTER the reference's own source() calls
 * (egEdgeroi.R:53-71).  Handles are R external pointers with finalizers; every numeric argument is an R double
 * vector / matrix (column-major, exactly what the ABI takes); status codes are returned to R, which maps
 * 1 / 2 to badnll = 9E99 (fitIAK3D.R:685,785-806) and negative values to stop().
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <string.h>

#include "iak3d.h"

static iak3d_ctx* g_ctx = NULL;

static iak3d_ctx* ctx_get(void) {
  if (g_ctx == NULL) {
    int st = iak3d_ctx_create(0, &g_ctx);
    if (st != IAK3D_OK) error("iak3d: no sm_100-class GPU / library initialisation failed (status %d); there is no CPU fallback", st);
  }
  return g_ctx;
}

static void ds_finalizer(SEXP p) {
  iak3d_ds* ds = (iak3d_ds*)R_ExternalPtrAddr(p);
  if (ds) { iak3d_dataset_destroy(ds); R_ClearExternalPtr(p); }
}
static void fit_finalizer(SEXP p) {
  iak3d_fit* f = (iak3d_fit*)R_ExternalPtrAddr(p);
  if (f) { iak3d_fit_destroy(f); R_ClearExternalPtr(p); }
}

/* parsBTfmd (list from readParsIAK3D) + model switches -> struct iak3d_pars.  pv = c(modelx, cmeOpt, sdfdType_cd1,
 * sdfdType_cxd0, sdfdType_cxd1, quirk, ax, nux, ad, nud, cx0, cx1, cd1, cxd0, cxd1, cme, tau(3x2 row-major), quirk_scale,
 * spline coefficients (3 x IAK3D_MAX_SPL row-major, zero padded), nssre, cssre (IAK3D_MAX_SSRE, zero padded)) */
#define IAK3D_PV_LEN (23 + 3 * IAK3D_MAX_SPL + 1 + IAK3D_MAX_SSRE)
static void pars_from_ptr(const double* v, iak3d_pars* P);
static void pars_from(SEXP pv, iak3d_pars* P) {
  if (LENGTH(pv) != IAK3D_PV_LEN) error("iak3d: parameter vector must have %d entries", IAK3D_PV_LEN);
  pars_from_ptr(REAL(pv), P);
}
static void pars_from_ptr(const double* v, iak3d_pars* P) {
  memset(P, 0, sizeof(*P));
  P->modelx = (int32_t)v[0]; P->cmeOpt = (int32_t)v[1];
  P->sdfdType[0] = (int32_t)v[2]; P->sdfdType[1] = (int32_t)v[3]; P->sdfdType[2] = (int32_t)v[4];
  P->quirk_cd1_unscaled = (int32_t)v[5];
  P->ax = v[6]; P->nux = v[7]; P->ad = v[8]; P->nud = v[9];
  P->cx0 = v[10]; P->cx1 = v[11]; P->cd1 = v[12]; P->cxd0 = v[13]; P->cxd1 = v[14]; P->cme = v[15];
  for (int c = 0; c < 3; c++) { P->sdfdPars[c][0] = v[16 + 2 * c]; P->sdfdPars[c][1] = v[17 + 2 * c]; }
  P->quirk_scale = v[22];
  for (int c = 0; c < 3; c++) for (int k = 0; k < IAK3D_MAX_SPL; k++) P->sdfdSpl[c][k] = v[23 + c * IAK3D_MAX_SPL + k];
  P->nssre = (int32_t)v[23 + 3 * IAK3D_MAX_SPL];            /* site-specific random effects: count, then the variances */
  for (int k = 0; k < IAK3D_MAX_SSRE; k++) P->cssre[k] = v[24 + 3 * IAK3D_MAX_SPL + k];
}

/* setupIAK3D: xData (n x 2), dIData (n x 2, already rounded, fitIAK3D.R:127), W = cbind(X, z) or NULL */
static void need_real(SEXP x, const char* what) {
  if (TYPEOF(x) != REALSXP) error("iak3d: %s must be a double matrix / vector (use storage.mode(x) <- 'double')", what);
}
SEXP iak3dR_dataset(SEXP xy, SEXP dI, SEXP W) {
  need_real(xy, "xData"); need_real(dI, "dIData");
  if (!isNull(W)) need_real(W, "W");
  if (ncols(xy) != 2 || ncols(dI) != 2 || nrows(dI) != nrows(xy) || (!isNull(W) && nrows(W) != nrows(xy)))
    error("iak3d: xData and dIData must be n x 2 and W n x q");
  const int64_t n = nrows(xy);
  iak3d_ds* ds = NULL;
  const int q = isNull(W) ? 0 : ncols(W);
  int st = iak3d_dataset_create(ctx_get(), n, REAL(xy), REAL(dI), isNull(W) ? NULL : REAL(W), q, &ds);
  if (st != IAK3D_OK) error("iak3d_dataset_create: %s", iak3d_last_error(g_ctx));
  SEXP p = PROTECT(R_MakeExternalPtr(ds, R_NilValue, R_NilValue));
  R_RegisterCFinalizerEx(p, ds_finalizer, TRUE);
  UNPROTECT(1);
  return p;
}

SEXP iak3dR_set_W(SEXP dsp, SEXP W) {
  need_real(W, "W");
  int64_t n = 0;
  iak3d_dataset_info((iak3d_ds*)R_ExternalPtrAddr(dsp), &n, NULL, NULL, NULL);
  if ((int64_t)nrows(W) != n) error("iak3d: W must have the dataset's %d rows", (int)n);
  int32_t same = 0;                                   /* optim() passes the same data to every callback: upload once */
  if (iak3d_dataset_same_W((iak3d_ds*)R_ExternalPtrAddr(dsp), REAL(W), ncols(W), &same) == IAK3D_OK && same) return R_NilValue;
  int st = iak3d_dataset_set_W((iak3d_ds*)R_ExternalPtrAddr(dsp), REAL(W), ncols(W));
  if (st != IAK3D_OK) error("iak3d_dataset_set_W: %s", iak3d_last_error(g_ctx));
  return R_NilValue;
}

/* spline sd functions: setupMats$XsdfdSplineU_{cd1,cxd0,cxd1} (ndIU x (sdfdType + 1)) of component comp (0, 1, 2) */
SEXP iak3dR_set_sdfd_spline(SEXP dsp, SEXP comp, SEXP XsplU) {
  int st = iak3d_dataset_set_sdfd_spline((iak3d_ds*)R_ExternalPtrAddr(dsp), asInteger(comp), ncols(XsplU), REAL(XsplU));
  if (st != IAK3D_OK) error("iak3d_dataset_set_sdfd_spline: %s", iak3d_last_error(g_ctx));
  return R_NilValue;
}
/* lognormal data: column (0-based) of W that the device overwrites with sigma2Vec - diag(C) at every evaluation; -1 = off */
SEXP iak3dR_set_lnN_column(SEXP dsp, SEXP col) {
  int st = iak3d_dataset_set_lnN_column((iak3d_ds*)R_ExternalPtrAddr(dsp), asInteger(col));
  if (st != IAK3D_OK) error("iak3d_dataset_set_lnN_column: %s", iak3d_last_error(g_ctx));
  return R_NilValue;
}
/* diag(setCIAK3D(pars)$C): list(status, diagC) */
SEXP iak3dR_cov_diag(SEXP dsp, SEXP pv) {
  iak3d_ds* ds = (iak3d_ds*)R_ExternalPtrAddr(dsp);
  iak3d_pars P; pars_from(pv, &P);
  int64_t n = 0; iak3d_dataset_info(ds, &n, NULL, NULL, NULL);
  SEXP d = PROTECT(allocVector(REALSXP, (R_xlen_t)n));
  int st = iak3d_cov_diag(ds, &P, REAL(d));
  if (st < 0) error("iak3d_cov_diag: %s", iak3d_last_error(g_ctx));
  SEXP out = PROTECT(allocVector(VECSXP, 2));
  SET_VECTOR_ELT(out, 0, ScalarInteger(st)); SET_VECTOR_ELT(out, 1, d);
  UNPROTECT(2);
  return out;
}

/* nllIAK3D(..., forCompLik = TRUE) terms for `count` parameter vectors on one dataset (count = 1: a function value;
 * count = npar + 1: a forward-difference gradient).  pmat: IAK3D_PV_LEN x count.  Returns list(status, lndetA, WiAW (q*q x count)). */
SEXP iak3dR_nll_terms(SEXP dsp, SEXP pmat) {
  iak3d_ds* ds = (iak3d_ds*)R_ExternalPtrAddr(dsp);
  const int count = ncols(pmat);
  int32_t q = 0; int64_t n = 0;
  iak3d_dataset_info(ds, &n, NULL, NULL, &q);
  iak3d_pars* P = (iak3d_pars*)R_alloc(count, sizeof(iak3d_pars));
  iak3d_ds** dss = (iak3d_ds**)R_alloc(count, sizeof(iak3d_ds*));
  if (nrows(pmat) != IAK3D_PV_LEN) error("iak3d: parameter matrix must have %d rows", IAK3D_PV_LEN);
  for (int i = 0; i < count; i++) { pars_from_ptr(REAL(pmat) + IAK3D_PV_LEN * (size_t)i, &P[i]); dss[i] = ds; }
  SEXP status = PROTECT(allocVector(INTSXP, count));
  SEXP lndet = PROTECT(allocVector(REALSXP, count));
  SEXP G = PROTECT(allocMatrix(REALSXP, q * q, count));
  int st = iak3d_nll_terms_batch(g_ctx, count, dss, P, REAL(lndet), REAL(G), (int32_t*)INTEGER(status));
  if (st < 0) error("iak3d_nll_terms_batch: %s", iak3d_last_error(g_ctx));
  SEXP out = PROTECT(allocVector(VECSXP, 3));
  SET_VECTOR_ELT(out, 0, status); SET_VECTOR_ELT(out, 1, lndet); SET_VECTOR_ELT(out, 2, G);
  UNPROTECT(4);
  return out;
}

/* nllIAK3D(..., rtnAll = FALSE) in one call: device terms + the closed-form REML / ML tail in the library (fitIAK3D.R:730-880).
 * Returns the value itself (9E99 where the reference returns badnll). */
SEXP iak3dR_nll_reml(SEXP dsp, SEXP pv, SEXP useReml) {
  iak3d_ds* ds = (iak3d_ds*)R_ExternalPtrAddr(dsp);
  iak3d_pars P; pars_from(pv, &P);
  double nll = 9e99; int32_t status = 0;
  int st = iak3d_nll_reml(ds, &P, asLogical(useReml) ? 1 : 0, &nll, NULL, NULL, NULL, NULL, &status);
  if (st < 0) error("iak3d_nll_reml: %s", iak3d_last_error(g_ctx));
  return ScalarReal(status != 0 ? 9e99 : nll);
}

/* the switches and bounds readParsIAK3D needs, as the numeric vector .iak3d_model() builds (r/iak3d_cuda.R):
 * modelx, cmeOpt, prodSum, lnTfmdData, sdfdType x 3, nud, axBnds, nuxBnds, adBnds, tau2Bnds, sxd1Bnds (2 each), maxd */
#define IAK3D_MV_LEN 19
static void model_from(SEXP mv, iak3d_model* M) {
  need_real(mv, "model vector");
  if (LENGTH(mv) != IAK3D_MV_LEN) error("iak3d: model vector must have %d entries", IAK3D_MV_LEN);
  const double* v = REAL(mv);
  memset(M, 0, sizeof(*M));
  M->modelx = (int32_t)v[0]; M->cmeOpt = (int32_t)v[1]; M->prodSum = (int32_t)v[2]; M->lnTfmdData = (int32_t)v[3];
  for (int c = 0; c < 3; c++) M->sdfdType[c] = (int32_t)v[4 + c];
  M->nud = v[7];
  for (int e = 0; e < 2; e++) {
    M->axBnds[e] = v[8 + e]; M->nuxBnds[e] = v[10 + e]; M->adBnds[e] = v[12 + e]; M->tau2Bnds[e] = v[14 + e]; M->sxd1Bnds[e] = v[16 + e];
  }
  M->maxd = v[18];
}

/* nllIAK3D(pars, ...) as optim() calls it: the unconstrained vector in, the value out (readParsIAK3D, device terms and the tail
 * inside the library; fitIAK3D.R:670-949 with rtnAll = FALSE) */
SEXP iak3dR_nll_vector(SEXP dsp, SEXP mv, SEXP pars, SEXP useReml) {
  iak3d_ds* ds = (iak3d_ds*)R_ExternalPtrAddr(dsp);
  iak3d_model M; model_from(mv, &M);
  need_real(pars, "pars");
  double nll = 9e99;
  int st = iak3d_nll_vector(ds, &M, REAL(pars), LENGTH(pars), asLogical(useReml) ? 1 : 0, &nll);
  if (st < 0) error("iak3d_nll_vector: %s", iak3d_last_error(g_ctx));
  return ScalarReal(nll);
}

/* gradnllIAK3D (fitIAK3D.R:1776-1911): forward differences, npar + 1 evaluations in one launch.  Returns list(grad, nll0). */
SEXP iak3dR_gradnll_fd(SEXP dsp, SEXP mv, SEXP pars, SEXP delta, SEXP useReml) {
  iak3d_ds* ds = (iak3d_ds*)R_ExternalPtrAddr(dsp);
  iak3d_model M; model_from(mv, &M);
  need_real(pars, "pars");
  const int npar = LENGTH(pars);
  SEXP g = PROTECT(allocVector(REALSXP, npar));
  double nll0 = 9e99;
  int st = iak3d_gradnll_fd(ds, &M, REAL(pars), npar, asReal(delta), asLogical(useReml) ? 1 : 0, REAL(g), &nll0);
  if (st < 0) error("iak3d_gradnll_fd: %s", iak3d_last_error(g_ctx));
  SEXP out = PROTECT(allocVector(VECSXP, 2));
  SET_VECTOR_ELT(out, 0, g); SET_VECTOR_ELT(out, 1, ScalarReal(nll0));
  UNPROTECT(2);
  return out;
}

/* one evaluation that also returns iAW = A^-1 W (n x q) for rtnAll / forCompLik callers */
SEXP iak3dR_nll_terms_iAW(SEXP dsp, SEXP pv) {
  iak3d_ds* ds = (iak3d_ds*)R_ExternalPtrAddr(dsp);
  iak3d_pars P; pars_from(pv, &P);
  int32_t q = 0; int64_t n = 0;
  iak3d_dataset_info(ds, &n, NULL, NULL, &q);
  SEXP lndet = PROTECT(allocVector(REALSXP, 1));
  SEXP G = PROTECT(allocMatrix(REALSXP, q, q));
  SEXP iAW = PROTECT(allocMatrix(REALSXP, (int)n, q));
  int st = iak3d_nll_terms(ds, &P, REAL(lndet), REAL(G), REAL(iAW));
  if (st < 0) error("iak3d_nll_terms: %s", iak3d_last_error(g_ctx));
  SEXP out = PROTECT(allocVector(VECSXP, 4));
  SET_VECTOR_ELT(out, 0, ScalarInteger(st)); SET_VECTOR_ELT(out, 1, lndet); SET_VECTOR_ELT(out, 2, G); SET_VECTOR_ELT(out, 3, iAW);
  UNPROTECT(4);
  return out;
}

/* setCIAK3D: list(status, C (n x n), sigma2Vec) */
SEXP iak3dR_cov_build(SEXP dsp, SEXP pv, SEXP wantC) {
  iak3d_ds* ds = (iak3d_ds*)R_ExternalPtrAddr(dsp);
  iak3d_pars P; pars_from(pv, &P);
  int64_t n = 0; iak3d_dataset_info(ds, &n, NULL, NULL, NULL);
  const int want = asLogical(wantC);
  SEXP C = PROTECT(want ? allocMatrix(REALSXP, (int)n, (int)n) : allocVector(REALSXP, 0));
  SEXP s2 = PROTECT(allocVector(REALSXP, (R_xlen_t)n));
  int st = iak3d_cov_build(ds, &P, want ? REAL(C) : NULL, REAL(s2));
  if (st < 0) error("iak3d_cov_build: %s", iak3d_last_error(g_ctx));
  SEXP out = PROTECT(allocVector(VECSXP, 3));
  SET_VECTOR_ELT(out, 0, ScalarInteger(st)); SET_VECTOR_ELT(out, 1, C); SET_VECTOR_ELT(out, 2, s2);
  UNPROTECT(3);
  return out;
}

/* setCIAK3D2: cross-covariance between rows (xy1, dI1) and the dataset */
SEXP iak3dR_cov_cross(SEXP dsp, SEXP pv, SEXP xy1, SEXP dI1) {
  iak3d_ds* ds = (iak3d_ds*)R_ExternalPtrAddr(dsp);
  iak3d_pars P; pars_from(pv, &P);
  int64_t n = 0; iak3d_dataset_info(ds, &n, NULL, NULL, NULL);
  const int n1 = nrows(xy1);
  SEXP out = PROTECT(allocMatrix(REALSXP, n1, (int)n));
  int st = iak3d_cov_cross(ds, &P, n1, REAL(xy1), REAL(dI1), REAL(out));
  if (st < 0) error("iak3d_cov_cross: %s", iak3d_last_error(g_ctx));
  UNPROTECT(1);
  return st == IAK3D_OK ? out : ScalarReal(NA_REAL);
}

/* setCApproxIAK3D2: the same with the spline bases of set 1 on its ROWS (list of 3: matrix n1 x (sdfdType + 1) or NULL) */
SEXP iak3dR_cov_cross_spl(SEXP dsp, SEXP pv, SEXP xy1, SEXP dI1, SEXP spl) {
  iak3d_ds* ds = (iak3d_ds*)R_ExternalPtrAddr(dsp);
  iak3d_pars P; pars_from(pv, &P);
  int64_t n = 0; iak3d_dataset_info(ds, &n, NULL, NULL, NULL);
  const int n1 = nrows(xy1);
  const double* sp[3];
  for (int c = 0; c < 3; c++) { SEXP e = VECTOR_ELT(spl, c); sp[c] = isNull(e) ? NULL : REAL(e); }
  SEXP out = PROTECT(allocMatrix(REALSXP, n1, (int)n));
  int st = iak3d_cov_cross_spl(ds, &P, n1, REAL(xy1), REAL(dI1), sp, REAL(out));
  if (st < 0) error("iak3d_cov_cross_spl: %s", iak3d_last_error(g_ctx));
  UNPROTECT(1);
  return st == IAK3D_OK ? out : ScalarReal(NA_REAL);
}

/* lndetANDinvCb(C, b): list(lndetC, invCb) with NA, NA on failure (optim_functions.R:292-294) */
SEXP iak3dR_lndet_invCb(SEXP C, SEXP b) {
  const int n = nrows(C);
  const int nrhs = isNull(b) ? 0 : ncols(b);
  SEXP lnd = PROTECT(allocVector(REALSXP, 1));
  SEXP inv = PROTECT(nrhs ? allocMatrix(REALSXP, n, nrhs) : allocMatrix(REALSXP, n, n));
  int st = iak3d_lndet_invCb(ctx_get(), n, REAL(C), nrhs, nrhs ? REAL(b) : NULL, REAL(lnd), REAL(inv));
  if (st < 0) error("iak3d_lndet_invCb: %s", iak3d_last_error(g_ctx));
  SEXP out = PROTECT(allocVector(VECSXP, 2));
  if (st == IAK3D_OK) { SET_VECTOR_ELT(out, 0, lnd); SET_VECTOR_ELT(out, 1, inv); }
  else { SET_VECTOR_ELT(out, 0, ScalarReal(NA_REAL)); SET_VECTOR_ELT(out, 1, ScalarReal(NA_REAL)); }
  UNPROTECT(3);
  return out;
}

/* kept fit (the big matrices of lmmFit) and kriging */
SEXP iak3dR_fit(SEXP dsp, SEXP pv, SEXP betahat, SEXP vbetahat) {
  iak3d_pars P; pars_from(pv, &P);
  iak3d_fit* f = NULL;
  int st = iak3d_fit_create((iak3d_ds*)R_ExternalPtrAddr(dsp), &P, REAL(betahat), REAL(vbetahat), &f);
  if (st != IAK3D_OK) error("iak3d_fit_create: status %d %s", st, iak3d_last_error(g_ctx));
  SEXP p = PROTECT(R_MakeExternalPtr(f, R_NilValue, dsp));      /* keeps the dataset alive */
  R_RegisterCFinalizerEx(p, fit_finalizer, TRUE);
  UNPROTECT(1);
  return p;
}
/* the same with the residual z - muhat given (lognormal data: muhat = setmuIAK3D(...)) */
SEXP iak3dR_fit_res(SEXP dsp, SEXP pv, SEXP betahat, SEXP vbetahat, SEXP z_muhat) {
  iak3d_pars P; pars_from(pv, &P);
  iak3d_fit* f = NULL;
  int st = iak3d_fit_create_res((iak3d_ds*)R_ExternalPtrAddr(dsp), &P, REAL(betahat), REAL(vbetahat), REAL(z_muhat), &f);
  if (st != IAK3D_OK) error("iak3d_fit_create_res: status %d %s", st, iak3d_last_error(g_ctx));
  SEXP p = PROTECT(R_MakeExternalPtr(f, R_NilValue, dsp));
  R_RegisterCFinalizerEx(p, fit_finalizer, TRUE);
  UNPROTECT(1);
  return p;
}
SEXP iak3dR_fit_option(SEXP fp, SEXP option, SEXP value) {
  int st = iak3d_fit_set_option((iak3d_fit*)R_ExternalPtrAddr(fp), asInteger(option), asInteger(value));
  if (st != IAK3D_OK) error("iak3d_fit_set_option: %s", iak3d_last_error(g_ctx));
  return R_NilValue;
}
SEXP iak3dR_fit_small(SEXP fp, SEXP n_, SEXP p_) {              /* iCX, iCz_muhat */
  const int n = asInteger(n_), p = asInteger(p_);
  SEXP iCX = PROTECT(allocMatrix(REALSXP, n, p));
  SEXP iCz = PROTECT(allocMatrix(REALSXP, n, 1));
  int st = iak3d_fit_get((iak3d_fit*)R_ExternalPtrAddr(fp), REAL(iCX), REAL(iCz), NULL, NULL);
  if (st != IAK3D_OK) error("iak3d_fit_get: %s", iak3d_last_error(g_ctx));
  SEXP out = PROTECT(allocVector(VECSXP, 2));
  SET_VECTOR_ELT(out, 0, iCX); SET_VECTOR_ELT(out, 1, iCz);
  UNPROTECT(3);
  return out;
}
SEXP iak3dR_predict(SEXP fp, SEXP xyMap, SEXP dIMap, SEXP XMap) {
  const int m = nrows(xyMap);
  SEXP z = PROTECT(allocVector(REALSXP, m));
  SEXP v = PROTECT(allocVector(REALSXP, m));
  int st = iak3d_predict((iak3d_fit*)R_ExternalPtrAddr(fp), m, REAL(xyMap), REAL(dIMap), REAL(XMap), REAL(z), REAL(v));
  if (st != IAK3D_OK) error("iak3d_predict: %s", iak3d_last_error(g_ctx));
  SEXP out = PROTECT(allocVector(VECSXP, 2));
  SET_VECTOR_ELT(out, 0, z); SET_VECTOR_ELT(out, 1, v);
  UNPROTECT(3);
  return out;
}

/* kriging with all its pieces: list(z, v, Ckk, sigma2Veck, CkhiCz_muhat, CkhiCChk[, CkhiCX]); spl = list of 3 spline
 * bases on the map supports (or NULL); wantCX needs iak3dR_fit_option(fit, 1, 1) beforehand */
SEXP iak3dR_predict_terms(SEXP fp, SEXP xyMap, SEXP dIMap, SEXP XMap, SEXP spl, SEXP wantCX) {
  iak3d_fit* f = (iak3d_fit*)R_ExternalPtrAddr(fp);
  const int m = nrows(xyMap), p = ncols(XMap);
  const double* sp[3] = {NULL, NULL, NULL};
  if (!isNull(spl)) for (int c = 0; c < 3; c++) { SEXP e = VECTOR_ELT(spl, c); sp[c] = isNull(e) ? NULL : REAL(e); }
  int st = iak3d_predict_stage_spl(f, m, REAL(xyMap), REAL(dIMap), REAL(XMap), isNull(spl) ? NULL : sp);
  if (st == IAK3D_OK) st = iak3d_predict_enqueue(f);
  if (st != IAK3D_OK) error("iak3d_predict: %s", iak3d_last_error(g_ctx));
  const int want = asLogical(wantCX);
  SEXP out = PROTECT(allocVector(VECSXP, want ? 7 : 6));
  for (int k = 0; k < 6; k++) SET_VECTOR_ELT(out, k, allocVector(REALSXP, m));
  st = iak3d_predict_fetch(f, REAL(VECTOR_ELT(out, 0)), REAL(VECTOR_ELT(out, 1)));
  if (st == IAK3D_OK) st = iak3d_predict_fetch_terms(f, REAL(VECTOR_ELT(out, 2)), REAL(VECTOR_ELT(out, 3)), REAL(VECTOR_ELT(out, 4)), REAL(VECTOR_ELT(out, 5)));
  if (st == IAK3D_OK && want) { SET_VECTOR_ELT(out, 6, allocMatrix(REALSXP, m, p)); st = iak3d_predict_fetch_CkhiCX(f, REAL(VECTOR_ELT(out, 6))); }
  if (st != IAK3D_OK) error("iak3d_predict_fetch: %s", iak3d_last_error(g_ctx));
  UNPROTECT(1);
  return out;
}

/* analytic gradient of the profiled objective: pmat = IAK3D_PV_LEN x (npar + 1) candidates (column 1 = the point itself),
 * delta (npar).  list(status, lndetA, WiAW (q x q), gsum (npar)); d nll / d theta = gsum / 2 */
SEXP iak3dR_nll_grad(SEXP dsp, SEXP pmat, SEXP delta, SEXP useReml) {
  iak3d_ds* ds = (iak3d_ds*)R_ExternalPtrAddr(dsp);
  const int count = ncols(pmat), npar = count - 1;
  int32_t q = 0; iak3d_dataset_info(ds, NULL, NULL, NULL, &q);
  if (nrows(pmat) != IAK3D_PV_LEN || LENGTH(delta) != npar) error("iak3d: nll_grad needs %d x (npar + 1) parameters and npar steps", IAK3D_PV_LEN);
  iak3d_pars* P = (iak3d_pars*)R_alloc(count, sizeof(iak3d_pars));
  for (int i = 0; i < count; i++) pars_from_ptr(REAL(pmat) + IAK3D_PV_LEN * (size_t)i, &P[i]);
  SEXP out = PROTECT(allocVector(VECSXP, 4));
  SET_VECTOR_ELT(out, 1, allocVector(REALSXP, 1)); SET_VECTOR_ELT(out, 2, allocMatrix(REALSXP, q, q)); SET_VECTOR_ELT(out, 3, allocVector(REALSXP, npar));
  int st = iak3d_nll_grad(ds, npar, P, REAL(delta), asLogical(useReml), REAL(VECTOR_ELT(out, 1)), REAL(VECTOR_ELT(out, 2)), REAL(VECTOR_ELT(out, 3)));
  if (st < 0) error("iak3d_nll_grad: %s", iak3d_last_error(g_ctx));
  SET_VECTOR_ELT(out, 0, ScalarInteger(st));
  UNPROTECT(1);
  return out;
}

/* gradHessIAK3D (nrUpdatesIAK3D.R:118-216): list(status, nll, grad, hess, fim, betahat, vbetahat) */
SEXP iak3dR_gradhess(SEXP dsp, SEXP pv, SEXP lnvPars) {
  iak3d_ds* ds = (iak3d_ds*)R_ExternalPtrAddr(dsp);
  iak3d_pars P; pars_from(pv, &P);
  int32_t q = 0; iak3d_dataset_info(ds, NULL, NULL, NULL, &q);
  const int m = LENGTH(lnvPars), p = q - 1;
  SEXP out = PROTECT(allocVector(VECSXP, 7));
  SET_VECTOR_ELT(out, 1, allocVector(REALSXP, 1)); SET_VECTOR_ELT(out, 2, allocVector(REALSXP, m));
  SET_VECTOR_ELT(out, 3, allocMatrix(REALSXP, m, m)); SET_VECTOR_ELT(out, 4, allocMatrix(REALSXP, m, m));
  SET_VECTOR_ELT(out, 5, allocMatrix(REALSXP, p, 1)); SET_VECTOR_ELT(out, 6, allocMatrix(REALSXP, p, p));
  int st = iak3d_gradhess(ds, &P, m, REAL(lnvPars), REAL(VECTOR_ELT(out, 1)), REAL(VECTOR_ELT(out, 2)), REAL(VECTOR_ELT(out, 3)),
                          REAL(VECTOR_ELT(out, 4)), REAL(VECTOR_ELT(out, 5)), REAL(VECTOR_ELT(out, 6)));
  if (st < 0) error("iak3d_gradhess: %s", iak3d_last_error(g_ctx));
  SET_VECTOR_ELT(out, 0, ScalarInteger(st));
  UNPROTECT(1);
  return out;
}

/* ---- device-resident matrices (iak3d_mat_*): the algebra of predMatsIAK3D_CLEV_ByBlock, compositeLikIAK3D.R:609-798 ---- */
static void mat_finalizer(SEXP p) {
  iak3d_mat* M = (iak3d_mat*)R_ExternalPtrAddr(p);
  if (M) { iak3d_mat_destroy(M); R_ClearExternalPtr(p); }
}
static SEXP mat_wrap(iak3d_mat* M) {
  SEXP p = PROTECT(R_MakeExternalPtr(M, R_NilValue, R_NilValue));
  R_RegisterCFinalizerEx(p, mat_finalizer, TRUE);
  UNPROTECT(1);
  return p;
}
#define MAT(p) ((iak3d_mat*)R_ExternalPtrAddr(p))
/* matrix(0, rows, cols) on the device, or as.matrix(init) when init is not NULL */
SEXP iak3dR_mat(SEXP rows, SEXP cols, SEXP init) {
  iak3d_mat* M = NULL;
  const int64_t nr = isNull(init) ? (int64_t)asReal(rows) : nrows(init), nc = isNull(init) ? (int64_t)asReal(cols) : ncols(init);
  int st = iak3d_mat_create(ctx_get(), nr, nc, &M);
  if (st == IAK3D_OK && !isNull(init)) st = iak3d_mat_set(M, 0, 0, nr, nc, REAL(init), nr);
  if (st != IAK3D_OK) { if (M) iak3d_mat_destroy(M); error("iak3d_mat_create: %s", iak3d_last_error(g_ctx)); }
  return mat_wrap(M);
}
/* as.matrix(M[r0 + 1:nr, c0 + 1:nc]) */
SEXP iak3dR_mat_get(SEXP Mp, SEXP blk) {
  const double* b = REAL(blk);
  const int64_t r0 = (int64_t)b[0], c0 = (int64_t)b[1], nr = (int64_t)b[2], nc = (int64_t)b[3];
  SEXP out = PROTECT(allocMatrix(REALSXP, (int)nr, (int)nc));
  int st = iak3d_mat_get(MAT(Mp), r0, c0, nr, nc, REAL(out), nr);
  if (st != IAK3D_OK) error("iak3d_mat_get: %s", iak3d_last_error(g_ctx));
  UNPROTECT(1);
  return out;
}
/* D[dr0.., dc0..] <- beta * D[..] + alpha * S[sr0.., sc0..] (t(.) when tr); blk = c(dr0, dc0, sr0, sc0, nr, nc, tr, alpha, beta) */
SEXP iak3dR_mat_copy(SEXP Dp, SEXP Sp, SEXP blk) {
  const double* b = REAL(blk);
  int st = iak3d_mat_copy(MAT(Dp), (int64_t)b[0], (int64_t)b[1], MAT(Sp), (int64_t)b[2], (int64_t)b[3], (int64_t)b[4], (int64_t)b[5],
                          (int32_t)b[6], b[7], b[8]);
  if (st != IAK3D_OK) error("iak3d_mat_copy: %s", iak3d_last_error(g_ctx));
  return R_NilValue;
}
/* C <- (acc ? C : 0) + alpha * A %*% t(B) */
SEXP iak3dR_mat_gemm_nt(SEXP Cp, SEXP alpha, SEXP Ap, SEXP Bp, SEXP acc) {
  int st = iak3d_mat_gemm_nt(MAT(Cp), asReal(alpha), MAT(Ap), MAT(Bp), asInteger(acc));
  if (st != IAK3D_OK) error("iak3d_mat_gemm_nt: %s", iak3d_last_error(g_ctx));
  return R_NilValue;
}
/* M <- chol2inv(chol(M)); stops like chol() when M is not positive definite */
SEXP iak3dR_mat_spd_inverse(SEXP Mp) {
  int32_t status = 0;
  int st = iak3d_mat_spd_inverse(MAT(Mp), &status);
  if (st < 0) error("iak3d_mat_spd_inverse: %s", iak3d_last_error(g_ctx));
  if (status != 0) error("the leading minor is not positive definite");
  return R_NilValue;
}
/* setCIAK3D(...)$C of the dataset as a device matrix; NULL for invalid parameters (R: C = NA) */
SEXP iak3dR_mat_cov_build(SEXP dsp, SEXP pv) {
  iak3d_ds* ds = (iak3d_ds*)R_ExternalPtrAddr(dsp);
  iak3d_pars P; pars_from(pv, &P);
  int64_t n = 0;
  iak3d_dataset_info(ds, &n, NULL, NULL, NULL);
  iak3d_mat* M = NULL;
  int st = iak3d_mat_create(ctx_get(), n, n, &M);
  if (st == IAK3D_OK) st = iak3d_mat_cov_build(ds, &P, M);
  if (st == IAK3D_BAD_PARS) { iak3d_mat_destroy(M); return R_NilValue; }
  if (st != IAK3D_OK) { if (M) iak3d_mat_destroy(M); error("iak3d_mat_cov_build: %s", iak3d_last_error(g_ctx)); }
  return mat_wrap(M);
}
/* end of a prediction call: return the cached matrix memory to the driver */
SEXP iak3dR_mat_trim(void) {
  if (iak3d_mat_trim(ctx_get()) != IAK3D_OK) error("iak3d_mat_trim: %s", iak3d_last_error(g_ctx));
  return R_NilValue;
}
/* colSums(A * B) */
SEXP iak3dR_mat_coldot(SEXP Ap, SEXP Bp) {
  int64_t nc = 0;
  iak3d_mat_shape(MAT(Ap), NULL, &nc);
  SEXP out = PROTECT(allocVector(REALSXP, (R_xlen_t)nc));
  int st = iak3d_mat_coldot(MAT(Ap), MAT(Bp), REAL(out));
  if (st != IAK3D_OK) error("iak3d_mat_coldot: %s", iak3d_last_error(g_ctx));
  UNPROTECT(1);
  return out;
}

/* ---- composite likelihood: every unit (block pair / block) of compositeLikIAK3D.R:49-155 in ONE batched launch ----
 * dslist: list of dataset handles (one per unit, W already set); pv: the parameter vector they share.
 * Returns list(status (count), lndetA (count), WiAW (q*q x count)). */
SEXP iak3dR_nll_terms_units(SEXP dslist, SEXP pv) {
  const int count = LENGTH(dslist);
  if (count < 1) error("iak3d: empty unit list");
  iak3d_pars P0; pars_from(pv, &P0);
  iak3d_pars* P = (iak3d_pars*)R_alloc(count, sizeof(iak3d_pars));
  iak3d_ds** dss = (iak3d_ds**)R_alloc(count, sizeof(iak3d_ds*));
  for (int i = 0; i < count; i++) { P[i] = P0; dss[i] = (iak3d_ds*)R_ExternalPtrAddr(VECTOR_ELT(dslist, i)); }
  int32_t q = 0; int64_t n = 0;
  iak3d_dataset_info(dss[0], &n, NULL, NULL, &q);
  SEXP status = PROTECT(allocVector(INTSXP, count));
  SEXP lndet = PROTECT(allocVector(REALSXP, count));
  SEXP G = PROTECT(allocMatrix(REALSXP, q * q, count));
  int st = iak3d_nll_terms_batch(ctx_get(), count, dss, P, REAL(lndet), REAL(G), (int32_t*)INTEGER(status));
  if (st < 0) error("iak3d_nll_terms_batch: %s", iak3d_last_error(g_ctx));
  SEXP out = PROTECT(allocVector(VECSXP, 3));
  SET_VECTOR_ELT(out, 0, status); SET_VECTOR_ELT(out, 1, lndet); SET_VECTOR_ELT(out, 2, G);
  UNPROTECT(4);
  return out;
}

/* ---- makeXvX on the device (makeXvX.R:1; include/iak3d.h: iak3d_design_*) ---- */
static void design_finalizer(SEXP p) {
  iak3d_design* g = (iak3d_design*)R_ExternalPtrAddr(p);
  if (g) { iak3d_design_destroy(g); R_ClearExternalPtr(p); }
}
SEXP iak3dR_design(SEXP ispec, SEXP dspec) {     /* integer and double parts of the flat specification */
  iak3d_design* g = NULL;
  int st = iak3d_design_create(ctx_get(), (const int32_t*)INTEGER(ispec), LENGTH(ispec), REAL(dspec), LENGTH(dspec), &g);
  if (st != IAK3D_OK) error("iak3d_design_create: %s", iak3d_last_error(g_ctx));
  SEXP p = PROTECT(R_MakeExternalPtr(g, R_NilValue, R_NilValue));
  R_RegisterCFinalizerEx(p, design_finalizer, TRUE);
  UNPROTECT(1);
  return p;
}
SEXP iak3dR_design_eval(SEXP gp, SEXP covs, SEXP dI) {          /* covs: m x ncov (or NULL), dI: m x 2 -> X (m x p) */
  iak3d_design* g = (iak3d_design*)R_ExternalPtrAddr(gp);
  int32_t p = 0, ncov = 0;
  iak3d_design_info(g, &p, &ncov);
  const int m = nrows(dI);
  SEXP X = PROTECT(allocMatrix(REALSXP, m, p));
  int st = iak3d_design_eval(g, m, isNull(covs) ? NULL : REAL(covs), REAL(dI), REAL(X), NULL, NULL);
  if (st != IAK3D_OK) error("iak3d_design_eval: %s", iak3d_last_error(g_ctx));
  UNPROTECT(1);
  return X;
}
/* the whole map of predictIAK3D in one staging: list(zMap, vMap), each ndIMap x nxMap like the reference's */
SEXP iak3dR_predict_grid(SEXP fp, SEXP gp, SEXP xyCell, SEXP covCell, SEXP dIInt, SEXP nxPerBatch) {
  iak3d_fit* f = (iak3d_fit*)R_ExternalPtrAddr(fp);
  iak3d_design* g = (iak3d_design*)R_ExternalPtrAddr(gp);
  const int ncell = nrows(xyCell), nint = nrows(dIInt);
  int st = iak3d_predict_stage_grid(f, g, ncell, REAL(xyCell), isNull(covCell) ? NULL : REAL(covCell), nint, REAL(dIInt), (int64_t)asReal(nxPerBatch));
  if (st == IAK3D_OK) st = iak3d_predict_enqueue(f);
  if (st != IAK3D_OK) error("iak3d_predict_stage_grid: %s", iak3d_last_error(g_ctx));
  SEXP zt = PROTECT(allocMatrix(REALSXP, ncell, nint));           /* support (i, c) is element i * ncell + c: t(zMap) */
  SEXP vt = PROTECT(allocMatrix(REALSXP, ncell, nint));
  st = iak3d_predict_fetch(f, REAL(zt), REAL(vt));
  if (st != IAK3D_OK) error("iak3d_predict_fetch: %s", iak3d_last_error(g_ctx));
  SEXP out = PROTECT(allocVector(VECSXP, 2));
  SET_VECTOR_ELT(out, 0, zt); SET_VECTOR_ELT(out, 1, vt);
  UNPROTECT(3);
  return out;
}

/* ---- block construction (compositeLikIAK3D.R:803-1133) ---- */
SEXP iak3dR_nearest_centroid(SEXP xy, SEXP vcentres, SEXP wantCounts) {   /* 1-based block of every location [, counts] */
  iak3d_points* P = NULL;
  const int m = nrows(xy), nB = nrows(vcentres);
  int st = iak3d_points_create(ctx_get(), m, REAL(xy), &P);
  if (st != IAK3D_OK) error("iak3d_points_create: %s", iak3d_last_error(g_ctx));
  SEXP blk = PROTECT(allocVector(INTSXP, m));
  SEXP cnt = PROTECT(allocVector(REALSXP, nB));
  int64_t* c64 = (int64_t*)R_alloc(nB, sizeof(int64_t));
  st = iak3d_points_nearest(P, nB, REAL(vcentres), (int32_t*)INTEGER(blk), asLogical(wantCounts) ? c64 : NULL);
  iak3d_points_destroy(P);
  if (st != IAK3D_OK) error("iak3d_points_nearest: %s", iak3d_last_error(g_ctx));
  for (int i = 0; i < m; i++) INTEGER(blk)[i] += 1;
  for (int b = 0; b < nB; b++) REAL(cnt)[b] = asLogical(wantCounts) ? (double)c64[b] : 0.0;
  SEXP out = PROTECT(allocVector(VECSXP, 2));
  SET_VECTOR_ELT(out, 0, blk); SET_VECTOR_ELT(out, 1, cnt);
  UNPROTECT(3);
  return out;
}
SEXP iak3dR_voronoi_seed(SEXP xU, SEXP nPerBlock) {                 /* vcentres (nBlocks x 2) */
  const int nU = nrows(xU), npb = asInteger(nPerBlock);
  const int nB = nU / (npb > 0 ? npb : 1);
  if (nB < 1) error("iak3d: fewer unique locations than nPerBlock");
  SEXP vc = PROTECT(allocMatrix(REALSXP, nB, 2));
  int32_t got = 0;
  int st = iak3d_voronoi_seed(nU, REAL(xU), npb, nB, &got, REAL(vc));
  if (st != IAK3D_OK) error("iak3d_voronoi_seed failed (status %d)", st);
  UNPROTECT(1);
  return vc;
}
SEXP iak3dR_dirichlet_adjacency(SEXP vcentres) {                    /* subsetPairsAdj: npairs x 2, 1-based, i < j */
  const int nB = nrows(vcentres);
  const int cap = 3 * nB + 3;
  int32_t* pr = (int32_t*)R_alloc(2 * (size_t)cap, sizeof(int32_t));
  int64_t np = 0;
  int st = iak3d_dirichlet_adjacency(nB, REAL(vcentres), pr, cap, &np);
  if (st != IAK3D_OK) error("iak3d_dirichlet_adjacency failed (status %d)", st);
  SEXP out = PROTECT(allocMatrix(INTSXP, (int)np, 2));
  for (int k = 0; k < (int)np; k++) { INTEGER(out)[k] = pr[2 * k] + 1; INTEGER(out)[(int)np + k] = pr[2 * k + 1] + 1; }
  UNPROTECT(1);
  return out;
}

/* ---- site-specific random effects (siteIDData given; iaCovMatern.R:440-460): site id (integer codes, equal = same site) and
 * the columns XData[, colnames4ssre] of every row ---- */
SEXP iak3dR_set_ssre(SEXP dsp, SEXP siteid, SEXP Xs) {
  need_real(Xs, "XData[, colnames4ssre]");
  if (TYPEOF(siteid) != INTSXP) error("iak3d: site ids must be integer codes");
  int st = iak3d_dataset_set_ssre((iak3d_ds*)R_ExternalPtrAddr(dsp), ncols(Xs), (const int32_t*)INTEGER(siteid), REAL(Xs));
  if (st != IAK3D_OK) error("iak3d_dataset_set_ssre: %s", iak3d_last_error(g_ctx));
  return R_NilValue;
}
SEXP iak3dR_cov_cross_ssre(SEXP dsp, SEXP pv, SEXP xy1, SEXP dI1, SEXP siteid1, SEXP Xs1) {
  iak3d_ds* ds = (iak3d_ds*)R_ExternalPtrAddr(dsp);
  iak3d_pars P; pars_from(pv, &P);
  int64_t n = 0; iak3d_dataset_info(ds, &n, NULL, NULL, NULL);
  const int n1 = nrows(xy1);
  SEXP out = PROTECT(allocMatrix(REALSXP, n1, (int)n));
  int st = iak3d_cov_cross_ssre(ds, &P, n1, REAL(xy1), REAL(dI1), (const int32_t*)INTEGER(siteid1), REAL(Xs1), REAL(out));
  if (st < 0) error("iak3d_cov_cross_ssre: %s", iak3d_last_error(g_ctx));
  UNPROTECT(1);
  return st == 0 ? out : R_NilValue;
}
/* kriging of map supports that carry site ids: list(z, v) */
SEXP iak3dR_predict_ssre(SEXP fp, SEXP xyMap, SEXP dIMap, SEXP XMap, SEXP siteid, SEXP Xs) {
  iak3d_fit* f = (iak3d_fit*)R_ExternalPtrAddr(fp);
  const int m = nrows(xyMap);
  int st = iak3d_predict_stage_ssre(f, m, REAL(xyMap), REAL(dIMap), REAL(XMap), (const int32_t*)INTEGER(siteid), REAL(Xs));
  if (st == IAK3D_OK) st = iak3d_predict_enqueue(f);
  if (st != IAK3D_OK) error("iak3d_predict_stage_ssre: %s", iak3d_last_error(g_ctx));
  SEXP z = PROTECT(allocVector(REALSXP, m));
  SEXP v = PROTECT(allocVector(REALSXP, m));
  st = iak3d_predict_fetch(f, REAL(z), REAL(v));
  if (st != IAK3D_OK) error("iak3d_predict_fetch: %s", iak3d_last_error(g_ctx));
  SEXP out = PROTECT(allocVector(VECSXP, 2));
  SET_VECTOR_ELT(out, 0, z); SET_VECTOR_ELT(out, 1, v);
  UNPROTECT(3);
  return out;
}

static const R_CallMethodDef call_methods[] = {
  {"iak3dR_dataset", (DL_FUNC)&iak3dR_dataset, 3},       {"iak3dR_set_W", (DL_FUNC)&iak3dR_set_W, 2},
  {"iak3dR_nll_reml", (DL_FUNC)&iak3dR_nll_reml, 3},
  {"iak3dR_nll_vector", (DL_FUNC)&iak3dR_nll_vector, 4}, {"iak3dR_gradnll_fd", (DL_FUNC)&iak3dR_gradnll_fd, 5},
  {"iak3dR_nll_terms", (DL_FUNC)&iak3dR_nll_terms, 2},   {"iak3dR_nll_terms_iAW", (DL_FUNC)&iak3dR_nll_terms_iAW, 2},
  {"iak3dR_cov_build", (DL_FUNC)&iak3dR_cov_build, 3},   {"iak3dR_cov_cross", (DL_FUNC)&iak3dR_cov_cross, 4},
  {"iak3dR_lndet_invCb", (DL_FUNC)&iak3dR_lndet_invCb, 2}, {"iak3dR_fit", (DL_FUNC)&iak3dR_fit, 4},
  {"iak3dR_fit_small", (DL_FUNC)&iak3dR_fit_small, 3},   {"iak3dR_predict", (DL_FUNC)&iak3dR_predict, 4},
  {"iak3dR_set_sdfd_spline", (DL_FUNC)&iak3dR_set_sdfd_spline, 3}, {"iak3dR_set_lnN_column", (DL_FUNC)&iak3dR_set_lnN_column, 2},
  {"iak3dR_cov_diag", (DL_FUNC)&iak3dR_cov_diag, 2},     {"iak3dR_cov_cross_spl", (DL_FUNC)&iak3dR_cov_cross_spl, 5},
  {"iak3dR_fit_res", (DL_FUNC)&iak3dR_fit_res, 5},       {"iak3dR_fit_option", (DL_FUNC)&iak3dR_fit_option, 3},
  {"iak3dR_predict_terms", (DL_FUNC)&iak3dR_predict_terms, 6}, {"iak3dR_gradhess", (DL_FUNC)&iak3dR_gradhess, 3},
  {"iak3dR_nll_grad", (DL_FUNC)&iak3dR_nll_grad, 4},
  {"iak3dR_mat", (DL_FUNC)&iak3dR_mat, 3},               {"iak3dR_mat_get", (DL_FUNC)&iak3dR_mat_get, 2},
  {"iak3dR_mat_copy", (DL_FUNC)&iak3dR_mat_copy, 3},     {"iak3dR_mat_gemm_nt", (DL_FUNC)&iak3dR_mat_gemm_nt, 5},
  {"iak3dR_mat_spd_inverse", (DL_FUNC)&iak3dR_mat_spd_inverse, 1}, {"iak3dR_mat_cov_build", (DL_FUNC)&iak3dR_mat_cov_build, 2},
  {"iak3dR_mat_coldot", (DL_FUNC)&iak3dR_mat_coldot, 2}, {"iak3dR_mat_trim", (DL_FUNC)&iak3dR_mat_trim, 0},
  {"iak3dR_nll_terms_units", (DL_FUNC)&iak3dR_nll_terms_units, 2},
  {"iak3dR_design", (DL_FUNC)&iak3dR_design, 2},         {"iak3dR_design_eval", (DL_FUNC)&iak3dR_design_eval, 3},
  {"iak3dR_predict_grid", (DL_FUNC)&iak3dR_predict_grid, 6},
  {"iak3dR_nearest_centroid", (DL_FUNC)&iak3dR_nearest_centroid, 3}, {"iak3dR_voronoi_seed", (DL_FUNC)&iak3dR_voronoi_seed, 2},
  {"iak3dR_dirichlet_adjacency", (DL_FUNC)&iak3dR_dirichlet_adjacency, 1},
  {"iak3dR_set_ssre", (DL_FUNC)&iak3dR_set_ssre, 3},     {"iak3dR_cov_cross_ssre", (DL_FUNC)&iak3dR_cov_cross_ssre, 6},
  {"iak3dR_predict_ssre", (DL_FUNC)&iak3dR_predict_ssre, 6},
  {NULL, NULL, 0}};

void R_init_iak3dR(DllInfo* dll) {
  R_registerRoutines(dll, NULL, call_methods, NULL, NULL);
  R_useDynamicSymbols(dll, FALSE);
}
