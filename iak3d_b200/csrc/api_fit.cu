// api_fit.cu -- the kept fit (big matrices of lmmFit, fitIAK3D.R:919-936) and block kriging
// (per-batch body of predictIAK3D, predictIAK3D.R:179-255) behind the C ABI.
//
// Kriging never forms iC.  With C = L L' kept on the device and Yw = L^-1 [X, z - X betahat] kept in the bordered
// rows of the factor buffer, a panel of cross-covariance rows Ckh is pushed through the same tile engine as the
// factorisation (mode 1: X = Ckh L^-T, DMMA), and
//     Ckh iC X = X Yw(:,1:p),   Ckh iC (z - mu) = X Yw(:,p+1),   rowSums((Ckh iC) * Ckh) = rowSums(X * X)
// come out of the tile epilogues -- n^2 flops per prediction instead of the reference's 2 n^2.
#include <math.h>
#include <string.h>

#include <algorithm>

#include "common.cuh"

// from api.cu
void iak_unique_first(int64_t n, const double* c0, const double* c1, std::vector<int32_t>& id, std::vector<double>& U);
int iak_make_plan(iak3d_ctx* ctx, const iak3d_pars* p, bool square, EvalPlan* pl);
int iak_alloc_slot(iak3d_ctx* ctx, int64_t n, int32_t q, int64_t nxU, int64_t ndIU, Slot** out, bool tables_only = false);
void iak_free_slot(Slot* s);
struct TableCache;
int iak_square_tables(iak3d_ctx* ctx, const iak3d_ds* ds, const EvalPlan& pl, Slot* s, CovArgs* a, TableCache* cache = nullptr, TableQueue* tq = nullptr);
int iak_factor_slot(iak3d_ctx* ctx, Slot* s, int32_t* status_out, double* lndet_out, double* G_out);
int iak_backsolve_to_host(iak3d_ctx* ctx, Slot* s, double* out_host, double* d_out_opt);
int iak_inverse_from_factor(iak3d_ctx* ctx, Slot* s, double* d_iC);
int iak_plan_check_avsd(iak3d_ctx* ctx, const iak3d_ds* ds, EvalPlan* pl);
int iak_avsd_host(iak3d_ctx* ctx, const EvalPlan& pl, int k, const double* dIU, int64_t nd, const double* XsplU, int ncol,
                  std::vector<double>& s);
int iak_cross_depth_tables(iak3d_ctx* ctx, const iak3d_ds* ds2, const EvalPlan& pl, const double* d_d1, const double* dIU1,
                           int64_t nd1, const std::vector<double>* spl1U, double* d_T, double* d_s, const double** Tout);
void iak_spl_unique_rows(const double* Xrow, int64_t n1, int ncol, const std::vector<int32_t>& did1, int64_t nd1, std::vector<double>& out);

// Prediction rows per panel: two 128-row blocks per persistent CTA (37888 rows on a B200).  Each row block is a serial
// chain of column tasks (X_rj needs X_r,0..j-1), so the panel should hold more chains than CTAs; a taller panel also
// amortises the panel build and the end-of-panel tail.  Bounded so that the panel stays below ~24 GB for large fits.
static int64_t fit_chunk_rows(const iak3d_ctx* ctx, int64_t Npad) {
  int64_t rows = (int64_t)2 * ctx->sm_count * TM;
  while (rows > 16384 && (double)rows * (double)Npad * 8.0 > 24e9) rows -= (int64_t)ctx->sm_count * TM / 2;
  return rows < 16384 ? 16384 : rows;
}

struct iak3d_fit {
  iak3d_ds* ds = nullptr;
  iak3d_ctx* ctx = nullptr;
  iak3d_pars pars;
  EvalPlan plan_sq, plan_x, plan_kk;   // C of the fit, cross-covariance Ckh, and Ckk = diag(setCIAK3D(map)) plans
  Slot* slot = nullptr;          // L, inv diag blocks, Yw rows
  int32_t p = 0, q = 0;
  double* d_W2 = nullptr;        // n x q: [X, z - X betahat]
  double* d_beta = nullptr;      // p
  double* d_vbeta = nullptr;     // p x p
  // ---- staged prediction job ----
  int64_t m = 0;
  int64_t nd1 = 0;
  double* d_xy = nullptr;        // per chunk: rows x 2 column-major blocks
  double* d_X = nullptr;         // m x p column-major
  int32_t* d_did1 = nullptr;     // m
  int32_t* d_sidMap = nullptr; double* d_XsMap = nullptr; int64_t ssre_cap = 0; bool ssre_staged = false;   // site ids / ssre columns of the map supports
  double* d_dIU1 = nullptr;      // nd1 x 2
  std::vector<double> h_dIU1;    // host copy
  std::vector<double> spl1U[3];  // spline sd basis of the map supports' unique intervals (nd1 x ncol), setCApproxIAK3D2
  double* d_s = nullptr;         // averaged-sd vectors: 3 x (nd1cap + ndIU2) cross + 3 x nd1cap self
  double* d_z = nullptr; double* d_v = nullptr; double* d_ckk = nullptr;
  double* d_cz = nullptr; double* d_qf = nullptr; double* d_s2k = nullptr;   // CkhiCz_muhat, rowSums(CkhiC * Ckh), sigma2Veck
  double* d_avself = nullptr;    // NTAB_BUF x nd1cap: avVarVec of the map supports' self tables
  double* d_cx = nullptr; int64_t cx_cap = 0;   // m x p: Ckh iC X, kept when opt_keep_cx
  int32_t opt_keep_cx = 0, opt_cross_square = 0;
  int64_t cap_m = 0;
  // ---- panel workspace (one chunk) ----
  int64_t ldP = 0;
  double* d_P = nullptr;         // ldP x Npad
  double* d_G = nullptr;         // ldP x (q+1)
  int32_t* d_pflags = nullptr;   // (ldP/TM) x nCB
  int32_t pepoch = 0;
  double* d_phix = nullptr; uint8_t* d_same = nullptr;   // ldP x nxU2
  double* d_T = nullptr;         // 3 x nd1cap x ndIU2
  double* d_Tself = nullptr;     // 3 x nd1cap x nd1cap
  int64_t nd1cap = 0;
  int32_t* d_iota = nullptr;     // 0..ldP-1
  ProblemDesc* d_prob = nullptr;
  int4* d_tasks = nullptr; int32_t ntasks = 0; int64_t tasks_rows = 0;
  int32_t* d_ticket = nullptr; int32_t* d_status = nullptr; int32_t* h_status = nullptr;
  bool staged = false, enqueued = false;
};

static void fit_free_job(iak3d_fit* f) {
  cudaFree(f->d_xy); cudaFree(f->d_X); cudaFree(f->d_did1); cudaFree(f->d_dIU1); cudaFree(f->d_z); cudaFree(f->d_v); cudaFree(f->d_ckk);
  cudaFree(f->d_cz); cudaFree(f->d_qf); cudaFree(f->d_s2k); cudaFree(f->d_cx); f->d_cx = nullptr; f->cx_cap = 0;
  cudaFree(f->d_sidMap); cudaFree(f->d_XsMap); f->d_sidMap = nullptr; f->d_XsMap = nullptr; f->ssre_cap = 0; f->ssre_staged = false;
  f->d_xy = f->d_X = f->d_dIU1 = f->d_z = f->d_v = f->d_ckk = f->d_cz = f->d_qf = f->d_s2k = nullptr; f->d_did1 = nullptr; f->cap_m = 0;
}
static void fit_free_panel(iak3d_fit* f) {
  cudaFree(f->d_P); cudaFree(f->d_G); cudaFree(f->d_pflags); cudaFree(f->d_phix); cudaFree(f->d_same); cudaFree(f->d_T);
  cudaFree(f->d_Tself); cudaFree(f->d_iota); cudaFree(f->d_tasks); cudaFree(f->d_s); cudaFree(f->d_avself);
  f->d_P = f->d_G = f->d_phix = f->d_T = f->d_Tself = f->d_s = f->d_avself = nullptr; f->d_pflags = nullptr; f->d_same = nullptr; f->d_iota = nullptr;
  f->d_tasks = nullptr; f->ldP = 0; f->nd1cap = 0; f->tasks_rows = 0;
}

extern "C" int iak3d_fit_destroy(iak3d_fit* f) {
  if (!f) return IAK3D_OK;
  cudaSetDevice(f->ctx->device);
  cudaStreamSynchronize(f->ctx->stream);
  fit_free_job(f); fit_free_panel(f);
  iak_free_slot(f->slot);
  cudaFree(f->d_W2); cudaFree(f->d_beta); cudaFree(f->d_vbeta); cudaFree(f->d_prob); cudaFree(f->d_ticket); cudaFree(f->d_status);
  if (f->h_status) cudaFreeHost(f->h_status);
  delete f;
  return IAK3D_OK;
}

extern "C" int iak3d_fit_create(iak3d_ds* ds, const iak3d_pars* pars, const double* betahat, const double* vbetahat, iak3d_fit** out) {
  return iak3d_fit_create_res(ds, pars, betahat, vbetahat, nullptr, out);
}

extern "C" int iak3d_fit_create_res(iak3d_ds* ds, const iak3d_pars* pars, const double* betahat, const double* vbetahat,
                                    const double* z_muhat, iak3d_fit** out) {
  if (!ds || !pars || !out) return IAK3D_ERR_ARG;
  *out = nullptr;
  iak3d_ctx* ctx = ds->ctx;
  if (ds->q < 2 || !betahat || !vbetahat) { ctx->err = "fit_create: the dataset needs W = [X z] and betahat / vbetahat"; return IAK3D_ERR_ARG; }
  cudaSetDevice(ctx->device);
  iak3d_fit* f = new iak3d_fit();
  f->ds = ds; f->ctx = ctx; f->pars = *pars; f->q = ds->q; f->p = ds->q - 1;
  int rc = iak_make_plan(ctx, pars, true, &f->plan_sq);
  if (rc == 0) rc = iak_make_plan(ctx, pars, false, &f->plan_x);
  if (rc == 0) { iak3d_pars pk = *pars; pk.quirk_scale = 1.0; rc = iak_make_plan(ctx, &pk, true, &f->plan_kk); }
  if (rc == 0) rc = iak_plan_check_avsd(ctx, ds, &f->plan_sq);
  if (rc == 0 && f->plan_sq.status != 0) rc = f->plan_sq.status;
  if (rc != 0) { delete f; return rc; }
  const int64_t n = ds->n; const int p = f->p, q = f->q;
  // W2 = [X, z - X betahat]  (predictIAK3D.R:111,134)
  std::vector<double> W2(ds->h_W, ds->h_W + (size_t)n * q);
  if (z_muhat != nullptr) memcpy(&W2[(size_t)p * n], z_muhat, (size_t)n * 8);
  else for (int64_t i = 0; i < n; i++) {
    double mu = 0.0;
    for (int a = 0; a < p; a++) mu = mu + ds->h_W[(size_t)a * n + i] * betahat[a];
    W2[(size_t)p * n + i] = ds->h_W[(size_t)p * n + i] - mu;
  }
  do {
    cudaError_t e = cudaMalloc(&f->d_W2, (size_t)n * q * 8);
    if (e == cudaSuccess) e = cudaMalloc(&f->d_beta, (size_t)p * 8);
    if (e == cudaSuccess) e = cudaMalloc(&f->d_vbeta, (size_t)p * p * 8);
    if (e == cudaSuccess) e = cudaMalloc(&f->d_prob, sizeof(ProblemDesc));
    if (e == cudaSuccess) e = cudaMalloc(&f->d_ticket, 2 * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMalloc(&f->d_status, 2 * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMallocHost(&f->h_status, 2 * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMemcpyAsync(f->d_W2, W2.data(), (size_t)n * q * 8, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(f->d_beta, betahat, (size_t)p * 8, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(f->d_vbeta, vbetahat, (size_t)p * p * 8, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { ctx->err = std::string("fit_create: ") + cudaGetErrorString(e); rc = IAK3D_ERR_CUDA; cudaGetLastError(); break; }
    rc = iak_alloc_slot(ctx, n, q, ds->nxU, ds->ndIU, &f->slot);
    if (rc != 0) break;
    CovArgs a;
    rc = iak_square_tables(ctx, ds, f->plan_sq, f->slot, &a);
    if (rc != 0) break;
    Slot* s = f->slot;
    rc = launch_cov_factor_layout(ctx, a, f->d_W2, q, s->n, s->Npad, s->ld, s->nCB, s->d_L);
    if (rc != 0) break;
    int32_t status = 0;
    rc = iak_factor_slot(ctx, s, &status, nullptr, nullptr);
    if (rc != 0) break;
    if (status != 0) { rc = status; break; }
  } while (0);
  if (rc != 0) { iak3d_fit_destroy(f); return rc; }
  *out = f;
  return IAK3D_OK;
}

extern "C" int iak3d_fit_get(iak3d_fit* f, double* iCX, double* iCz_muhat, double* C, double* iC) {
  if (!f) return IAK3D_ERR_ARG;
  iak3d_ctx* ctx = f->ctx;
  cudaSetDevice(ctx->device);
  const int64_t n = f->ds->n; const int p = f->p, q = f->q;
  int rc = IAK3D_OK;
  if (iCX || iCz_muhat) {
    std::vector<double> tmp((size_t)n * q);
    rc = iak_backsolve_to_host(ctx, f->slot, tmp.data(), nullptr);
    if (rc != 0) return rc;
    if (iCX) memcpy(iCX, tmp.data(), (size_t)n * p * 8);
    if (iCz_muhat) memcpy(iCz_muhat, tmp.data() + (size_t)n * p, (size_t)n * 8);
  }
  if (C || iC) {
    double* d = nullptr;
    if (cudaMalloc(&d, (size_t)n * n * 8) != cudaSuccess) { ctx->err = "fit_get: allocation failed"; cudaGetLastError(); return IAK3D_ERR_CUDA; }
    do {
      if (C) {
        // the tables of the slot are still those of the fit parameters
        CovArgs a;
        rc = iak_square_tables(ctx, f->ds, f->plan_sq, f->slot, &a);
        if (rc != 0) break;
        rc = launch_cov_rect(ctx, a, d, n, n, n, 1);
        if (rc != 0) break;
        cudaMemcpyAsync(C, d, (size_t)n * n * 8, cudaMemcpyDeviceToHost, ctx->stream);
        if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { ctx->err = "fit_get: copy failed"; rc = IAK3D_ERR_CUDA; break; }
      }
      if (iC) {
        rc = iak_inverse_from_factor(ctx, f->slot, d);
        if (rc != 0) break;
        cudaMemcpyAsync(iC, d, (size_t)n * n * 8, cudaMemcpyDeviceToHost, ctx->stream);
        if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { ctx->err = "fit_get: copy failed"; rc = IAK3D_ERR_CUDA; break; }
      }
    } while (0);
    cudaFree(d);
  }
  return rc;
}

// ---- panel workspace ----------------------------------------------------------------------------
static int fit_reserve_panel(iak3d_fit* f, int64_t rows, int64_t nd1) {
  iak3d_ctx* ctx = f->ctx;
  Slot* s = f->slot;
  const int64_t ldP = round_up(rows, TM);
  if (ldP > f->ldP) {
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(f->d_P); cudaFree(f->d_G); cudaFree(f->d_pflags); cudaFree(f->d_phix); cudaFree(f->d_same); cudaFree(f->d_iota);
    f->d_P = f->d_G = f->d_phix = nullptr; f->d_pflags = nullptr; f->d_same = nullptr; f->d_iota = nullptr; f->ldP = 0;
    CUDA_TRY(ctx, iak_malloc(ctx, (void**)&f->d_P, (size_t)ubuf_doubles(ldP, s->Npad) * 8));
    CUDA_TRY(ctx, cudaMalloc(&f->d_G, (size_t)ldP * (f->q + 1) * 8));
    const size_t nfl = (size_t)(ldP / TM) * s->nCB;
    CUDA_TRY(ctx, cudaMalloc(&f->d_pflags, nfl * 4));
    CUDA_TRY(ctx, cudaMemsetAsync(f->d_pflags, 0, nfl * 4, ctx->stream));
    f->pepoch = 0;
    CUDA_TRY(ctx, cudaMalloc(&f->d_phix, (size_t)ldP * f->ds->nxU * 8));
    CUDA_TRY(ctx, cudaMalloc(&f->d_same, (size_t)ldP * f->ds->nxU));
    std::vector<int32_t> iota((size_t)ldP);
    for (int64_t i = 0; i < ldP; i++) iota[(size_t)i] = (int32_t)i;
    CUDA_TRY(ctx, cudaMalloc(&f->d_iota, (size_t)ldP * 4));
    CUDA_TRY(ctx, cudaMemcpyAsync(f->d_iota, iota.data(), (size_t)ldP * 4, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    f->ldP = ldP;
  }
  if (nd1 > f->nd1cap) {
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(f->d_T); cudaFree(f->d_Tself); cudaFree(f->d_s); cudaFree(f->d_avself); f->d_T = f->d_Tself = f->d_s = f->d_avself = nullptr;
    CUDA_TRY(ctx, cudaMalloc(&f->d_avself, (size_t)NTAB_BUF * nd1 * 8));
    CUDA_TRY(ctx, cudaMalloc(&f->d_T, (size_t)NTAB_BUF * nd1 * f->ds->ndIU * 8));
    CUDA_TRY(ctx, cudaMalloc(&f->d_Tself, (size_t)NTAB_BUF * nd1 * nd1 * 8));
    CUDA_TRY(ctx, cudaMalloc(&f->d_s, (size_t)(3 * (nd1 + f->ds->ndIU) + 3 * nd1) * 8));
    f->nd1cap = nd1;
  }
  return 0;
}

static int fit_panel_tasks(iak3d_fit* f, int64_t rowsP) {
  iak3d_ctx* ctx = f->ctx;
  if (f->tasks_rows == rowsP && f->d_tasks) return 0;
  const int nRBp = (int)(rowsP / TM), nCB = f->slot->nCB;
  std::vector<int4> tasks;
  tasks.reserve((size_t)nRBp * nCB);
  for (int j = 0; j < nCB; j++)
    for (int r = 0; r < nRBp; r++) tasks.push_back(make_int4(0, r, j, 0));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  cudaFree(f->d_tasks); f->d_tasks = nullptr;
  CUDA_TRY(ctx, cudaMalloc(&f->d_tasks, tasks.size() * sizeof(int4)));
  CUDA_TRY(ctx, cudaMemcpyAsync(f->d_tasks, tasks.data(), tasks.size() * sizeof(int4), cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  f->ntasks = (int32_t)tasks.size();
  f->tasks_rows = rowsP;
  return 0;
}

// per-support buffers of a staged job
static int fit_reserve_job(iak3d_fit* f, int64_t m) {
  iak3d_ctx* ctx = f->ctx;
  const int p = f->p;
  if (m > f->cap_m) {
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    fit_free_job(f);
    CUDA_TRY(ctx, cudaMalloc(&f->d_xy, (size_t)m * 16));
    CUDA_TRY(ctx, cudaMalloc(&f->d_X, (size_t)m * p * 8));
    CUDA_TRY(ctx, cudaMalloc(&f->d_did1, (size_t)m * 4));
    CUDA_TRY(ctx, cudaMalloc(&f->d_z, (size_t)m * 8));
    CUDA_TRY(ctx, cudaMalloc(&f->d_v, (size_t)m * 8));
    CUDA_TRY(ctx, cudaMalloc(&f->d_ckk, (size_t)m * 8));
    CUDA_TRY(ctx, cudaMalloc(&f->d_cz, (size_t)m * 8));
    CUDA_TRY(ctx, cudaMalloc(&f->d_qf, (size_t)m * 8));
    CUDA_TRY(ctx, cudaMalloc(&f->d_s2k, (size_t)m * 8));
    f->cap_m = m;
  }
  if (f->opt_keep_cx && (int64_t)m * p > f->cx_cap) {
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(f->d_cx); f->d_cx = nullptr;
    CUDA_TRY(ctx, cudaMalloc(&f->d_cx, (size_t)m * p * 8));
    f->cx_cap = (int64_t)m * p;
  }
  return 0;
}

extern "C" int iak3d_predict_stage(iak3d_fit* f, int64_t m, const double* xyMap, const double* dIMap, const double* XMap) {
  return iak3d_predict_stage_spl(f, m, xyMap, dIMap, XMap, nullptr);
}

extern "C" int iak3d_predict_stage_ssre(iak3d_fit* f, int64_t m, const double* xyMap, const double* dIMap, const double* XMap,
                                        const int32_t* siteidMap, const double* XsMap) {
  if (!f) return IAK3D_ERR_ARG;
  const int ns = f->plan_sq.nssre;
  if (ns > 0 && m > 0 && (!siteidMap || !XsMap)) { f->ctx->err = "predict_stage_ssre: the fit has site-specific random effects: siteidMap and XsMap are needed"; return IAK3D_ERR_ARG; }
  int rc = iak3d_predict_stage_spl(f, m, xyMap, dIMap, XMap, nullptr);
  if (rc != 0 || ns == 0 || m == 0) return rc;
  iak3d_ctx* ctx = f->ctx;
  f->staged = false;
  if (m > f->ssre_cap) {
    cudaFree(f->d_sidMap); cudaFree(f->d_XsMap); f->d_sidMap = nullptr; f->d_XsMap = nullptr; f->ssre_cap = 0;
    CUDA_TRY(ctx, cudaMalloc(&f->d_sidMap, (size_t)m * 4));
    CUDA_TRY(ctx, cudaMalloc(&f->d_XsMap, (size_t)m * IAK3D_MAX_SSRE * 8));
    f->ssre_cap = m;
  }
  CUDA_TRY(ctx, cudaMemcpyAsync(f->d_sidMap, siteidMap, (size_t)m * 4, cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(ctx, cudaMemcpyAsync(f->d_XsMap, XsMap, (size_t)m * ns * 8, cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  f->ssre_staged = true;
  f->staged = true;
  return IAK3D_OK;
}

extern "C" int iak3d_predict_stage_spl(iak3d_fit* f, int64_t m, const double* xyMap, const double* dIMap, const double* XMap,
                                       const double* const* Xspl1) {
  if (!f || m < 0 || (m > 0 && (!xyMap || !dIMap || !XMap))) return IAK3D_ERR_ARG;
  iak3d_ctx* ctx = f->ctx;
  cudaSetDevice(ctx->device);
  f->staged = false; f->enqueued = false; f->ssre_staged = false;
  f->m = m;
  if (m == 0) { f->staged = true; return IAK3D_OK; }
  const int p = f->p;
  { const int rcj = fit_reserve_job(f, m); if (rcj != 0) return rcj; }
  // unique depth intervals of the map supports (usually one per call: predictIAK3D.R:150)
  std::vector<int32_t> did1; std::vector<double> dIU1;
  iak_unique_first(m, dIMap, dIMap + m, did1, dIU1);
  f->nd1 = (int64_t)dIU1.size() / 2;
  f->h_dIU1 = dIU1;
  for (int c = 0; c < 3; c++) {
    f->spl1U[c].clear();
    if (f->pars.sdfdType[c] > 0 && Xspl1 && Xspl1[c]) iak_spl_unique_rows(Xspl1[c], m, f->pars.sdfdType[c] + 1, did1, f->nd1, f->spl1U[c]);
  }
  cudaFree(f->d_dIU1); f->d_dIU1 = nullptr;
  CUDA_TRY(ctx, cudaMalloc(&f->d_dIU1, dIU1.size() * 8));
  CUDA_TRY(ctx, cudaMemcpyAsync(f->d_dIU1, dIU1.data(), dIU1.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(ctx, cudaMemcpyAsync(f->d_did1, did1.data(), (size_t)m * 4, cudaMemcpyHostToDevice, ctx->stream));
  // coordinates: one (rows x 2) column-major block per chunk
  const int64_t PRED_CHUNK = fit_chunk_rows(ctx, f->slot->Npad);
  for (int64_t r0 = 0; r0 < m; r0 += PRED_CHUNK) {
    const int64_t rows = std::min(PRED_CHUNK, m - r0);
    CUDA_TRY(ctx, cudaMemcpyAsync(f->d_xy + 2 * r0, xyMap + r0, (size_t)rows * 8, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(f->d_xy + 2 * r0 + rows, xyMap + m + r0, (size_t)rows * 8, cudaMemcpyHostToDevice, ctx->stream));
  }
  CUDA_TRY(ctx, cudaMemcpyAsync(f->d_X, XMap, (size_t)m * p * 8, cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  int rc = fit_reserve_panel(f, std::min(PRED_CHUNK, m), f->nd1);
  if (rc != 0) return rc;
  f->staged = true;
  return IAK3D_OK;
}

// predictIAK3D's loops over depths and batches (predictIAK3D.R:144-155) for a whole grid in one staging: supports are
// (interval i, cell c) at row i * ncell + c -- the layout of zMap / vMap (ndIMap x nxMap, row-major here) -- and their design
// rows are made ON THE DEVICE by makeXvX (design.cu) from the per-cell covariates: 8 (2 + ncov) bytes per cell travel instead
// of 8 (4 + p) per support.
extern "C" int iak3d_predict_stage_grid(iak3d_fit* f, iak3d_design* g, int64_t ncell, const double* xyCell, const double* covCell,
                                        int64_t nint, const double* dIInt, int64_t nxPerBatch) {
  if (!f || !g || ncell < 0 || nint < 0) return IAK3D_ERR_ARG;
  iak3d_ctx* ctx = f->ctx;
  if (design_ctx(g) != ctx) { ctx->err = "predict_stage_grid: the design belongs to another context"; return IAK3D_ERR_ARG; }
  if (design_p(g) != f->p) { ctx->err = "predict_stage_grid: the design has a different number of columns than the fit's X"; return IAK3D_ERR_ARG; }
  for (int c = 0; c < 3; c++)
    if (f->pars.sdfdType[c] > 0) { ctx->err = "predict_stage_grid: spline sd functions need iak3d_predict_stage_spl"; return IAK3D_ERR_ARG; }
  const int64_t m = ncell * nint;
  const int ncov = design_ncov(g);
  if (m > 0 && (!xyCell || !dIInt || (ncov > 0 && !covCell))) return IAK3D_ERR_ARG;
  cudaSetDevice(ctx->device);
  f->staged = false; f->enqueued = false;
  f->m = m;
  if (m == 0) { f->staged = true; return IAK3D_OK; }
  { const int rcj = fit_reserve_job(f, m); if (rcj != 0) return rcj; }
  f->nd1 = nint;
  f->h_dIU1.assign(dIInt, dIInt + 2 * nint);
  for (int c = 0; c < 3; c++) f->spl1U[c].clear();
  cudaFree(f->d_dIU1); f->d_dIU1 = nullptr;
  CUDA_TRY(ctx, cudaMalloc(&f->d_dIU1, (size_t)nint * 16));
  CUDA_TRY(ctx, cudaMemcpyAsync(f->d_dIU1, dIInt, (size_t)nint * 16, cudaMemcpyHostToDevice, ctx->stream));
  double *d_xyc = nullptr, *d_cov = nullptr;
  int rc = IAK3D_OK;
  do {
    cudaError_t e = iak_pool_alloc(ctx, (void**)&d_xyc, (size_t)ncell * 16);
    if (e == cudaSuccess && ncov > 0) e = iak_pool_alloc(ctx, (void**)&d_cov, (size_t)ncell * ncov * 8);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_xyc, xyCell, (size_t)ncell * 16, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess && ncov > 0) e = cudaMemcpyAsync(d_cov, covCell, (size_t)ncell * ncov * 8, cudaMemcpyHostToDevice, ctx->stream);
    if (e != cudaSuccess) { ctx->err = std::string("predict_stage_grid: ") + cudaGetErrorString(e); cudaGetLastError(); rc = IAK3D_ERR_CUDA; break; }
    const int64_t PRED_CHUNK = fit_chunk_rows(ctx, f->slot->Npad);
    rc = launch_grid_supports(ctx, d_xyc, ncell, m, PRED_CHUNK, f->d_xy, f->d_did1);
    if (rc != 0) break;
    rc = launch_design_rows(ctx, g, m, ncell, nxPerBatch, d_cov, ncell, f->d_dIU1, nint, f->d_X, m, nullptr, nullptr);
    if (rc != 0) break;
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { ctx->err = "predict_stage_grid: staging failed"; cudaGetLastError(); rc = IAK3D_ERR_CUDA; break; }
    rc = fit_reserve_panel(f, std::min(PRED_CHUNK, m), f->nd1);
  } while (0);
  iak_pool_free(ctx, d_xyc); iak_pool_free(ctx, d_cov);
  if (rc != 0) return rc;
  f->staged = true;
  return IAK3D_OK;
}

// the design rows of the staged job as the predictor will use them (m x p column-major); parity tests
extern "C" int iak3d_predict_fetch_X(iak3d_fit* f, double* X) {
  if (!f || !X) return IAK3D_ERR_ARG;
  iak3d_ctx* ctx = f->ctx;
  if (!f->staged) { ctx->err = "predict_fetch_X: nothing staged"; return IAK3D_ERR_ARG; }
  cudaSetDevice(ctx->device);
  if (f->m == 0) return IAK3D_OK;
  CUDA_TRY(ctx, cudaMemcpyAsync(X, f->d_X, (size_t)f->m * f->p * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return IAK3D_OK;
}

// depth tables of the staged job -- cross (nd1 x nd2) and self (nd1 x nd1, for Ckk) -- Ckk, sigma2Veck, and the part of
// the panel's CovArgs that does not change from chunk to chunk
static int fit_prepare_job(iak3d_fit* f, CovArgs* ap) {
  iak3d_ctx* ctx = f->ctx;
  const int64_t m = f->m;
  iak3d_ds* ds = f->ds;
  const int64_t nd1 = f->nd1, nd2 = ds->ndIU;
  // compositeLikIAK3D.R:540-552 builds Ckh as a block of setCIAK3D of the joined supports: the square plan (cd1 quirk, :1052)
  const EvalPlan& px = f->opt_cross_square ? f->plan_kk : f->plan_x; const EvalPlan& ps = f->plan_kk;
  int rc;
  CovArgs& a = *ap;
  const double* Tself[3] = {nullptr, nullptr, nullptr};
  const double* avself[3] = {nullptr, nullptr, nullptr};
  rc = iak_cross_depth_tables(ctx, ds, px, f->d_dIU1, f->h_dIU1.data(), nd1, f->spl1U, f->d_T, f->d_s, a.T);
  if (rc != 0) return rc;
  if (!ps.approx) {
    for (int k = 0; k < ps.ntab; k++) {
      double* Ts = f->d_Tself + (size_t)k * nd1 * nd1;
      rc = launch_iacov_table(ctx, ps.prm[k], f->d_dIU1, nd1, f->d_dIU1, nd1, Ts, f->d_avself + (size_t)k * nd1);
      if (rc != 0) return rc;
      Tself[k] = Ts; avself[k] = f->d_avself + (size_t)k * nd1;
    }
  } else {
    // Ckk = diag(setCApproxIAK3D(map supports)) (predictIAK3D.R:183; fitIAK3D.R:959-962)
    int ksrc = -1;
    for (int k = 0; k < ps.ntab; k++) if (ps.sd_type[k] == 0) ksrc = k;
    const int bsrc = ksrc >= 0 ? ksrc : NTAB_BUF - 1;
    double* src = f->d_Tself + (size_t)bsrc * nd1 * nd1;
    rc = launch_iacov_table(ctx, ps.prm[0], f->d_dIU1, nd1, f->d_dIU1, nd1, src, f->d_avself + (size_t)bsrc * nd1);
    if (rc != 0) return rc;
    if (ksrc >= 0) { Tself[ksrc] = src; avself[ksrc] = f->d_avself + (size_t)bsrc * nd1; }
    std::vector<double> s1;
    for (int k = 0; k < ps.ntab; k++) {
      if (ps.sd_type[k] == 0) continue;
      const int c = ps.sd_comp[k];
      const int ncol1 = (!f->spl1U[c].empty() && nd1 > 0) ? (int)(f->spl1U[c].size() / (size_t)nd1) : 0;
      rc = iak_avsd_host(ctx, ps, k, f->h_dIU1.data(), nd1, f->spl1U[c].empty() ? nullptr : f->spl1U[c].data(), ncol1, s1);
      if (rc != 0) return rc;
      double* dsv = f->d_s + (size_t)3 * (nd1 + nd2) + (size_t)k * nd1;
      CUDA_TRY(ctx, cudaMemcpyAsync(dsv, s1.data(), (size_t)nd1 * 8, cudaMemcpyHostToDevice, ctx->stream));
      double* Ts = f->d_Tself + (size_t)k * nd1 * nd1;
      rc = launch_scale_table(ctx, src, dsv, nd1, dsv, nd1, 1, Ts, f->d_avself + (size_t)k * nd1);
      if (rc != 0) return rc;
      Tself[k] = Ts; avself[k] = f->d_avself + (size_t)k * nd1;
    }
  }
  rc = launch_ckk(ctx, ps, Tself, nd1, f->d_did1, m, f->d_ckk);
  if (rc != 0) return rc;
  if (ps.nssre > 0) {     // Ckk = diag(setCIAK3D(map supports)) with their own site structures (predictIAK3D.R:179-189)
    if (!f->ssre_staged) { ctx->err = "kriging: the fit has site-specific random effects: stage with iak3d_predict_stage_ssre"; return IAK3D_ERR_ARG; }
    rc = launch_ssre_diag(ctx, ps.nssre, ps.cssre, f->d_XsMap, m, m, f->d_ckk);
    if (rc != 0) return rc;
  }
  rc = launch_sigma2(ctx, ps, avself, f->d_did1, m, f->d_s2k);
  if (rc != 0) return rc;
  for (int c = 0; c < 3; c++) a.tidx[c] = px.tidx[c];
  a.ntab = px.ntab;
  a.cx0 = px.cx0; a.cx1 = px.cx1; a.kd1 = px.kd1; a.cxd0 = px.cxd0; a.cxd1 = px.cxd1; a.cme = 0.0;
  a.xid_c = ds->d_xid; a.did_c = ds->d_did; a.nc = ds->n; a.ndIU_r = nd1;
  a.nssre = px.nssre;
  for (int k = 0; k < px.nssre; k++) a.cssre[k] = px.cssre[k];
  a.sid_c = ds->d_sid; a.Xs_c = ds->d_Xs; a.ldXs_c = ds->n;
  return 0;
}

// the cross-covariance panel Ckh of map rows [r0, r0 + rows) in the blocked layout (rowsP x Npad, zero padded)
static int fit_build_panel(iak3d_fit* f, CovArgs& a, int64_t r0, int64_t rows, int64_t rowsP) {
  iak3d_ctx* ctx = f->ctx;
  iak3d_ds* ds = f->ds;
  const EvalPlan& px = f->opt_cross_square ? f->plan_kk : f->plan_x;
  // horizontal tables of this chunk: every map row is its own location
  int rc = launch_phix_table(ctx, px.H, f->d_xy + 2 * r0, rows, ds->d_xU, ds->nxU, px.has_phix ? f->d_phix : nullptr, f->d_same);
  if (rc != 0) return rc;
  a.phix = px.has_phix ? f->d_phix : nullptr;
  a.same = f->d_same;
  a.xid_r = f->d_iota; a.did_r = f->d_did1 + r0; a.nr = rows; a.nxU_r = rows;
  if (a.nssre > 0) { a.sid_r = f->d_sidMap + r0; a.Xs_r = f->d_XsMap + r0; a.ldXs_r = f->m; }
  return launch_cov_panel(ctx, a, f->d_P, rowsP, f->slot->Npad);
}

extern "C" int iak3d_predict_panel_fetch(iak3d_fit* f, int64_t chunk, double* Ckh, int64_t cap_rows, int64_t* rows_out) {
  if (!f || !Ckh || chunk < 0) return IAK3D_ERR_ARG;
  iak3d_ctx* ctx = f->ctx;
  if (!f->staged) { ctx->err = "predict_panel_fetch: nothing staged"; return IAK3D_ERR_ARG; }
  cudaSetDevice(ctx->device);
  const int64_t PRED_CHUNK = fit_chunk_rows(ctx, f->slot->Npad);
  const int64_t r0 = chunk * PRED_CHUNK;
  if (r0 >= f->m) { ctx->err = "predict_panel_fetch: no such chunk"; return IAK3D_ERR_ARG; }
  const int64_t rows = std::min(PRED_CHUNK, f->m - r0), rowsP = round_up(rows, TM), n = f->ds->n;
  if (rows_out) *rows_out = rows;
  if (cap_rows < rows) { ctx->err = "predict_panel_fetch: output too small"; return IAK3D_ERR_ARG; }
  CovArgs a;
  int rc = fit_prepare_job(f, &a);
  if (rc != 0) return rc;
  rc = fit_build_panel(f, a, r0, rows, rowsP);
  if (rc != 0) return rc;
  double* d_out = nullptr;
  if (iak_pool_alloc(ctx, (void**)&d_out, (size_t)rows * n * 8) != cudaSuccess) { ctx->err = "predict_panel_fetch: allocation failed"; cudaGetLastError(); return IAK3D_ERR_CUDA; }
  rc = launch_unblock_rect(ctx, f->d_P, rowsP / UR, 0, rows, n, 0, d_out, rows);
  if (rc == 0) {
    cudaMemcpyAsync(Ckh, d_out, (size_t)rows * n * 8, cudaMemcpyDeviceToHost, ctx->stream);
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { ctx->err = "predict_panel_fetch: copy failed"; rc = IAK3D_ERR_CUDA; }
  }
  iak_pool_free(ctx, d_out);
  return rc;
}

extern "C" int iak3d_predict_enqueue(iak3d_fit* f) {
  if (!f) return IAK3D_ERR_ARG;
  iak3d_ctx* ctx = f->ctx;
  if (!f->staged) { ctx->err = "predict_enqueue: nothing staged"; return IAK3D_ERR_ARG; }
  cudaSetDevice(ctx->device);
  f->enqueued = true;
  const int64_t m = f->m;
  if (m == 0) return IAK3D_OK;
  Slot* s = f->slot;
  CovArgs a;
  int rc = fit_prepare_job(f, &a);
  if (rc != 0) return rc;
  CUDA_TRY(ctx, cudaMemsetAsync(f->d_status, 0, 2 * sizeof(int32_t), ctx->stream));
  const int64_t PRED_CHUNK = fit_chunk_rows(ctx, s->Npad);
  for (int64_t r0 = 0; r0 < m; r0 += PRED_CHUNK) {
    const int64_t rows = std::min(PRED_CHUNK, m - r0);
    const int64_t rowsP = round_up(rows, TM);
    rc = fit_build_panel(f, a, r0, rows, rowsP);
    if (rc != 0) return rc;
    rc = fit_panel_tasks(f, rowsP);
    if (rc != 0) return rc;
    ProblemDesc pd;
    pd.L = s->d_L; pd.nRUL = s->ld / UR; pd.P = f->d_P; pd.nRUP = rowsP / UR; pd.invd = s->d_invd;
    pd.doneL = nullptr; pd.doneP = f->d_pflags; pd.invdone = nullptr;
    pd.nCB = s->nCB; pd.nRB = (int32_t)(rowsP / TM); pd.mode = 1; pd.epoch = ++f->pepoch;
    pd.w0 = (int32_t)s->Npad; pd.q = f->q; pd.G = f->d_G; pd.ldG = rowsP; pd.logsum = nullptr; pd.status = f->d_status;
    pd.prof = nullptr; pd.nlive = 0; pd.A = nullptr; pd.nRUA = 0; pd.kchunks = 0; pd.use_cin = 0; pd.alpha = 0.0;
    if (getenv("IAK3D_PROF") != nullptr) {      // diagnostics: per-task timestamps of the last panel (iak3d_last_task_profile)
      if (f->ntasks > ctx->prof_cap) {
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(ctx->d_prof); ctx->d_prof = nullptr;
        CUDA_TRY(ctx, cudaMalloc(&ctx->d_prof, (size_t)f->ntasks * 8 * sizeof(long long)));
        ctx->prof_cap = f->ntasks;
      }
      ctx->prof_ntasks = f->ntasks; ctx->d_prof_tasks = f->d_tasks;
      pd.prof = ctx->d_prof;
    }
    CUDA_TRY(ctx, cudaMemcpyAsync(f->d_prob, &pd, sizeof(pd), cudaMemcpyHostToDevice, ctx->stream));
    rc = launch_tile_tasks(ctx, f->d_prob, f->d_tasks, f->ntasks, f->d_ticket, 1, 0, (int)(rowsP / TM));
    if (rc != 0) return rc;
    rc = launch_krige_finish(ctx, f->d_G, rowsP, f->q, f->d_X + r0, m, rows, f->d_ckk + r0, f->d_beta, f->d_vbeta, f->d_z + r0,
                             f->d_v + r0, f->d_cz + r0, f->d_qf + r0, (f->opt_keep_cx && f->d_cx) ? f->d_cx + r0 : nullptr, m);
    if (rc != 0) return rc;
  }
  CUDA_TRY(ctx, cudaMemcpyAsync(f->h_status, f->d_status, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  return IAK3D_OK;
}

extern "C" int iak3d_predict_fetch(iak3d_fit* f, double* z, double* v) {
  if (!f) return IAK3D_ERR_ARG;
  iak3d_ctx* ctx = f->ctx;
  if (!f->enqueued) { ctx->err = "predict_fetch: nothing enqueued"; return IAK3D_ERR_ARG; }
  cudaSetDevice(ctx->device);
  f->enqueued = false;
  if (f->m == 0) return IAK3D_OK;
  if (z) CUDA_TRY(ctx, cudaMemcpyAsync(z, f->d_z, (size_t)f->m * 8, cudaMemcpyDeviceToHost, ctx->stream));
  if (v) CUDA_TRY(ctx, cudaMemcpyAsync(v, f->d_v, (size_t)f->m * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  if (f->h_status[1] != 0) { ctx->err = "kriging: dependency wait timed out (internal error)"; return IAK3D_ERR_TIMEOUT; }
  return IAK3D_OK;
}

extern "C" int iak3d_predict_fetch_terms(iak3d_fit* f, double* Ckk, double* sigma2Veck, double* CkhiCz_muhat, double* CkhiCChk) {
  if (!f) return IAK3D_ERR_ARG;
  iak3d_ctx* ctx = f->ctx;
  if (!f->staged) { ctx->err = "predict_fetch_terms: nothing staged"; return IAK3D_ERR_ARG; }
  cudaSetDevice(ctx->device);
  if (f->m == 0) return IAK3D_OK;
  const size_t bytes = (size_t)f->m * 8;
  if (Ckk) CUDA_TRY(ctx, cudaMemcpyAsync(Ckk, f->d_ckk, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  if (sigma2Veck) CUDA_TRY(ctx, cudaMemcpyAsync(sigma2Veck, f->d_s2k, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  if (CkhiCz_muhat) CUDA_TRY(ctx, cudaMemcpyAsync(CkhiCz_muhat, f->d_cz, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  if (CkhiCChk) CUDA_TRY(ctx, cudaMemcpyAsync(CkhiCChk, f->d_qf, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return IAK3D_OK;
}

extern "C" int iak3d_fit_set_option(iak3d_fit* f, int32_t option, int32_t value) {
  if (!f) return IAK3D_ERR_ARG;
  if (option == IAK3D_FIT_KEEP_CKHICX) f->opt_keep_cx = value ? 1 : 0;
  else if (option == IAK3D_FIT_CROSS_SQUARE) f->opt_cross_square = value ? 1 : 0;
  else { f->ctx->err = "fit_set_option: unknown option"; return IAK3D_ERR_ARG; }
  return IAK3D_OK;
}

extern "C" int iak3d_predict_fetch_CkhiCX(iak3d_fit* f, double* CkhiCX) {
  if (!f || !CkhiCX) return IAK3D_ERR_ARG;
  iak3d_ctx* ctx = f->ctx;
  if (!f->staged || !f->opt_keep_cx || !f->d_cx) { ctx->err = "predict_fetch_CkhiCX: set IAK3D_FIT_KEEP_CKHICX before staging"; return IAK3D_ERR_ARG; }
  cudaSetDevice(ctx->device);
  if (f->m == 0) return IAK3D_OK;
  CUDA_TRY(ctx, cudaMemcpyAsync(CkhiCX, f->d_cx, (size_t)f->m * f->p * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return IAK3D_OK;
}

extern "C" int iak3d_predict(iak3d_fit* f, int64_t m, const double* xyMap, const double* dIMap, const double* XMap, double* z,
                             double* v) {
  int rc = iak3d_predict_stage(f, m, xyMap, dIMap, XMap);
  if (rc != 0) return rc;
  rc = iak3d_predict_enqueue(f);
  if (rc != 0) return rc;
  return iak3d_predict_fetch(f, z, v);
}
