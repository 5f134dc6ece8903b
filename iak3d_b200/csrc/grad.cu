// grad.cu -- analytic gradient of the profiled REML / ML objective: replaces the npar + 1 likelihood
// evaluations of the forward-difference gradient (gradnllIAK3D, fitIAK3D.R:1865-1886) by ONE factorisation, the explicit
// inverse (triangular-aware, gemm.cu) and one pass over the lower triangle.
//
//   nll(theta) = 1/2 [ const + log|A| + (n - p) log s2 + log|X'A^-1 X| ]      (REML; ML: p -> 0 in the last two terms)
//   d nll / d theta_k = 1/2 sum_ij w_ij (dA/dtheta_k)_ij ,
//   w = A^-1 - [REML] A^-1 X (X'A^-1 X)^-1 X'A^-1 - u u' / s2 ,   u = A^-1 (z - X betahat),  s2 = cxdhat
//
// dA/dtheta_k is taken ENTRY-WISE from the unique-pair tables of the perturbed parameter set (theta + delta_k e_k through
// the host's transforms, exactly the candidates of the forward-difference gradient): (A_k - A_0)_ij / delta_k costs a few
// table loads per entry and no second factorisation; sets share tables (a perturbed variance changes no table at all),
// so an entry loads each DISTINCT table once.  All parameters are handled in the same pass over A^-1.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>

#include "common.cuh"

struct TableCache;
int iak_make_plan(iak3d_ctx* ctx, const iak3d_pars* p, bool square, EvalPlan* pl);
int iak_plan_check_avsd(iak3d_ctx* ctx, const iak3d_ds* ds, EvalPlan* pl);
int iak_get_slot(iak3d_ds* ds, size_t k, Slot** out);
int iak_get_table_slot(iak3d_ds* ds, size_t k, Slot** out);
TableCache* iak_table_cache_new();
void iak_table_cache_free(TableCache* c);
int iak_square_tables(iak3d_ctx* ctx, const iak3d_ds* ds, const EvalPlan& pl, Slot* s, CovArgs* a, TableCache* cache = nullptr, TableQueue* tq = nullptr);
int iak_factor_slot(iak3d_ctx* ctx, Slot* s, int32_t* status_out, double* lndet_out, double* G_out);
int iak_backsolve_to_host(iak3d_ctx* ctx, Slot* s, double* out_host, double* d_out_opt);
int iak_inverse_blocked(iak3d_ctx* ctx, Slot* s, double** out, int64_t* rowsU_out, double* U_keep, double* iC_keep);
int launch_panel_product(iak3d_ctx* ctx, const double* d_M, int64_t nRU, int64_t n, const double* d_W, int q, double* d_out);

namespace {

constexpr int GR_MAXSET = 20;    // parameter sets: the base point + up to 19 perturbed ones
constexpr int GR_MAXT = 12;      // distinct depth tables among them
constexpr int GR_MAXP = 6;       // distinct horizontal tables

struct GradSet { int itd, it0, it1, ipx; double cx0, cx1, kd1, cxd0, cxd1, cme, inv_delta; };
struct GradArgs {
  const double* Tu[GR_MAXT]; const double* Pu[GR_MAXP];
  int nTu, nPu, nset;
  GradSet set[GR_MAXSET];
  const int32_t* xid; const int32_t* did;
  int64_t n, nd, nx;
  const double* iC; int64_t nRU;      // blocked explicit inverse, lower triangle
  const double* Zt; int pz;           // pz x n: (A^-1 X) chol(X'A^-1 X)^-T transposed (REML), pz = 0 for ML
  const double* u; double invc;       // A^-1 (z - X betahat), 1 / cxdhat
  int nRB;
};

__global__ void __launch_bounds__(256) grad_pass_kernel(const __grid_constant__ GradArgs g, double* __restrict__ partial) {
  __shared__ double s_val[GR_MAXT + GR_MAXP][256];      // this thread's table values of the current entry (dynamic table index)
  __shared__ double s_red[8][GR_MAXSET];
  // tile (r, j) of the lower triangle, folded grid as in cov_factor_kernel is not worth it here: 1-D list, decode by search
  const int nRB = g.nRB;
  int j, r;
  {
    const int64_t t = blockIdx.x;
    auto before = [&](int64_t c) -> int64_t { const int64_t h = c / 2; return c * (int64_t)nRB - (h * (h - 1) + (c & 1) * h); };
    int lo = 0, hi = (int)((g.n + TN - 1) / TN) - 1;
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (before(mid) <= t) lo = mid; else hi = mid - 1; }
    j = lo; r = (int)(t - before(j)) + j / 2;
  }
  const int tid = threadIdx.x, rp = tid & 63, cg = tid >> 6;
  const int64_t R0 = (int64_t)r * TM, C0 = (int64_t)j * TN, n = g.n;
  const int64_t i0 = R0 + 2 * rp;
  double acc[GR_MAXSET];
#pragma unroll
  for (int s = 0; s < GR_MAXSET; s++) acc[s] = 0.0;
  int xi[2], di[2]; double ui[2];
#pragma unroll
  for (int e = 0; e < 2; e++) {
    const int64_t i = i0 + e;
    xi[e] = i < n ? g.xid[i] : -1; di[e] = i < n ? g.did[i] : 0; ui[e] = i < n ? g.u[i] : 0.0;
  }
  for (int cc = 0; cc < 16; cc++) {
    const int64_t c = C0 + cg * 16 + cc;
    if (c >= n) break;
    const int xj = g.xid[c], dj = g.did[c];
    const double uj = g.u[c];
#pragma unroll
    for (int e = 0; e < 2; e++) {
      const int64_t i = i0 + e;
      if (i >= n || i < c) continue;                                   // lower triangle incl. the diagonal
      // weight of this entry
      double w = g.iC[uidx(i, c, g.nRU)];
      for (int a = 0; a < g.pz; a++) w = fma(-__ldg(g.Zt + (int64_t)a * n + i), __ldg(g.Zt + (int64_t)a * n + c), w);
      w = fma(-ui[e] * uj, g.invc, w);
      if (i != c) w *= 2.0;                                            // the mirrored entry
      // the distinct table values of this entry
#pragma unroll
      for (int k = 0; k < GR_MAXT; k++) if (k < g.nTu) s_val[k][tid] = __ldg(g.Tu[k] + (int64_t)dj * g.nd + di[e]);
#pragma unroll
      for (int k = 0; k < GR_MAXP; k++) if (k < g.nPu) s_val[GR_MAXT + k][tid] = __ldg(g.Pu[k] + (int64_t)xj * g.nx + xi[e]);
      const bool same = (xi[e] == xj), diag = (i == c);
      double v0 = 0.0;
#pragma unroll
      for (int s = 0; s < GR_MAXSET; s++) {
        if (s < g.nset) {
          const GradSet& S = g.set[s];
          double v = diag ? S.cme : 0.0;
          if (same) v += S.cx0 + S.cxd0 * s_val[S.it0][tid];
          if (S.itd >= 0) v += S.kd1 * s_val[S.itd][tid];
          if (S.ipx >= 0) v += s_val[GR_MAXT + S.ipx][tid] * (S.cx1 + S.cxd1 * s_val[S.it1][tid]);
          if (s == 0) v0 = v; else acc[s] = fma(w * S.inv_delta, v - v0, acc[s]);
        }
      }
    }
  }
  // deterministic block reduction: warp shuffle tree, then the 8 warps in order
  const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int s = 1; s < GR_MAXSET; s++) {
    if (s < g.nset) {
      double v = acc[s];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
      if (lane == 0) s_red[warp][s] = v;
    }
  }
  __syncthreads();
  if (tid < g.nset && tid >= 1) {
    double v = 0.0;
    for (int wv = 0; wv < 8; wv++) v += s_red[wv][tid];
    partial[(int64_t)blockIdx.x * (GR_MAXSET - 1) + tid - 1] = v;
  }
}

bool host_chol_lower(int p, const double* A, double* L) {
  for (int jj = 0; jj < p; jj++) {
    for (int i = jj; i < p; i++) {
      double s = A[(size_t)jj * p + i];
      for (int k = 0; k < jj; k++) s -= L[(size_t)k * p + i] * L[(size_t)k * p + jj];
      if (i == jj) { if (!(s > 0.0) || !isfinite(s)) return false; L[(size_t)jj * p + jj] = sqrt(s); }
      else L[(size_t)jj * p + i] = s / L[(size_t)jj * p + jj];
    }
    for (int i = 0; i < jj; i++) L[(size_t)jj * p + i] = 0.0;
  }
  return true;
}

}  // namespace

extern "C" int iak3d_nll_grad(iak3d_ds* ds, int32_t npar, const iak3d_pars* pars, const double* delta, int32_t useReml,
                              double* lndetA, double* WiAW, double* gsum) {
  if (!ds || !pars || !delta || !gsum || npar < 1) return IAK3D_ERR_ARG;
  iak3d_ctx* ctx = ds->ctx;
  if (npar + 1 > GR_MAXSET) { ctx->err = "nll_grad: too many parameters"; return IAK3D_ERR_ARG; }
  if (ds->q < 2) { ctx->err = "nll_grad: the dataset needs W = [X z]"; return IAK3D_ERR_ARG; }
  cudaSetDevice(ctx->device);
  if (ctx->batch_pending()) { ctx->err = "nll_grad: a batch is pending"; return IAK3D_ERR_ARG; }
  const int64_t n = ds->n;
  const int q = ds->q, p = q - 1;
  const double nanv = nan("");
  for (int k = 0; k < npar; k++) gsum[k] = nanv;
  const bool prof = getenv("IAK3D_GRADPROF") != nullptr;
  auto tnow = [&]() { if (prof) cudaStreamSynchronize(ctx->stream); return std::chrono::steady_clock::now(); };
  auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
  const auto t0 = tnow();
  // ---- plans and tables of the npar + 1 parameter sets (tables are shared wherever the parameters allow) ----
  std::vector<EvalPlan> plans((size_t)npar + 1);
  for (int k = 0; k <= npar; k++) {
    int st = iak_make_plan(ctx, &pars[k], true, &plans[(size_t)k]);
    if (st == 0) st = iak_plan_check_avsd(ctx, ds, &plans[(size_t)k]);
    if (st != 0) return st;
    if (plans[(size_t)k].nssre > 0) { ctx->err = "nll_grad: the analytic gradient does not cover site-specific random effects (use the forward-difference batch)"; return IAK3D_ERR_ARG; }
  }
  if (plans[0].status != 0) return plans[0].status;
  Slot* s0 = nullptr;
  int rc = iak_get_slot(ds, 0, &s0);
  if (rc != 0) return rc;
  std::vector<CovArgs> cargs((size_t)npar + 1);
  TableCache* cache = iak_table_cache_new();
  for (int k = 0; k <= npar && rc == 0; k++) {
    if (plans[(size_t)k].status != 0) continue;                         // that component of the gradient stays NaN
    Slot* sk = s0;
    if (k > 0) rc = iak_get_table_slot(ds, (size_t)k - 1, &sk);
    if (rc == 0) rc = iak_square_tables(ctx, ds, plans[(size_t)k], sk, &cargs[(size_t)k], cache);
  }
  iak_table_cache_free(cache);
  if (rc != 0) return rc;
  const auto t1 = tnow();
  // ---- base point: build, factorise, A^-1 W ----
  Slot* s = s0;
  rc = launch_cov_factor_layout(ctx, cargs[0], ds->d_W, q, s->n, s->Npad, s->ld, s->nCB, s->d_L);
  if (rc != 0) return rc;
  int32_t status = 0; double lnd = 0.0;
  std::vector<double> G((size_t)q * q);
  rc = iak_factor_slot(ctx, s, &status, &lnd, G.data());
  if (rc != 0) return rc;
  if (status != 0) return status;
  if (lndetA) *lndetA = lnd;
  if (WiAW) memcpy(WiAW, G.data(), (size_t)q * q * sizeof(double));
  const auto t2 = tnow();
  // ---- explicit inverse (blocked, symmetric after the mirror); its two big buffers are kept with the dataset ----
  {
    const int64_t rowsU0 = round_up(s->Npad, TM);
    const size_t bytes = (size_t)ubuf_doubles(rowsU0, s->Npad) * sizeof(double);
    if (ds->grad_bytes != bytes) {
      CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
      cudaFree(ds->d_gradU); cudaFree(ds->d_gradiC); ds->d_gradU = ds->d_gradiC = nullptr; ds->grad_bytes = 0;
      CUDA_TRY(ctx, iak_malloc(ctx, (void**)&ds->d_gradU, bytes));
      CUDA_TRY(ctx, iak_malloc(ctx, (void**)&ds->d_gradiC, bytes));
      ds->grad_bytes = bytes;
    }
  }
  double* iCb = nullptr; int64_t rowsU = 0;
  rc = iak_inverse_blocked(ctx, s, &iCb, &rowsU, ds->d_gradU, ds->d_gradiC);
  if (rc != 0) return rc;
  const auto t3 = tnow();
  // ---- A^-1 W = iC W: one pass over the inverse instead of a block back-substitution ----
  std::vector<double> iAW((size_t)n * q);
  double* d_iAW = nullptr;
  if (iak_pool_alloc(ctx, (void**)&d_iAW, (size_t)n * q * 8) != cudaSuccess) { ctx->err = "nll_grad: allocation failed"; cudaGetLastError(); return IAK3D_ERR_CUDA; }
  rc = launch_panel_product(ctx, iCb, rowsU / UR, n, ds->d_W, q, d_iAW);
  if (rc == 0) {
    cudaMemcpyAsync(iAW.data(), d_iAW, (size_t)n * q * 8, cudaMemcpyDeviceToHost, ctx->stream);
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { ctx->err = "nll_grad: product failed"; rc = IAK3D_ERR_CUDA; }
  }
  iak_pool_free(ctx, d_iAW);
  if (rc != 0) return rc;
  // ---- p x p algebra on the host: betahat, cxdhat, Z = (A^-1 X) chol(M)^-T, u = A^-1 z - (A^-1 X) betahat ----
  std::vector<double> M((size_t)p * p), Lm((size_t)p * p), b((size_t)p), beta((size_t)p);
  for (int c = 0; c < p; c++) for (int r2 = 0; r2 < p; r2++) M[(size_t)c * p + r2] = (r2 <= c) ? G[(size_t)c * q + r2] : G[(size_t)r2 * q + c];   // upper triangle kept (forceSymmetric)
  for (int r2 = 0; r2 < p; r2++) b[(size_t)r2] = G[(size_t)p * q + r2];
  const double ziAz = G[(size_t)p * q + p];
  if (!host_chol_lower(p, M.data(), Lm.data())) return IAK3D_NOT_SPD;
  {
    std::vector<double> y((size_t)p);
    for (int i = 0; i < p; i++) { double t = b[(size_t)i]; for (int k = 0; k < i; k++) t -= Lm[(size_t)k * p + i] * y[(size_t)k]; y[(size_t)i] = t / Lm[(size_t)i * p + i]; }
    for (int i = p - 1; i >= 0; i--) { double t = y[(size_t)i]; for (int k = i + 1; k < p; k++) t -= Lm[(size_t)i * p + k] * beta[(size_t)k]; beta[(size_t)i] = t / Lm[(size_t)i * p + i]; }
  }
  double bXz = 0.0, bXXb = 0.0;
  for (int r2 = 0; r2 < p; r2++) { bXz += beta[(size_t)r2] * b[(size_t)r2]; double t = 0.0; for (int c = 0; c < p; c++) t += M[(size_t)c * p + r2] * beta[(size_t)c]; bXXb += beta[(size_t)r2] * t; }
  const double res = ziAz - 2.0 * bXz + bXXb;
  const double cxd = useReml ? res / (double)(n - p) : res / (double)n;
  if (!(cxd > 0.0) || !isfinite(cxd)) return IAK3D_NOT_SPD;
  const int pz = useReml ? p : 0;
  std::vector<double> Zt((size_t)std::max(pz, 1) * n), u((size_t)n);
  for (int64_t i = 0; i < n; i++) {
    double t = iAW[(size_t)p * n + i];
    for (int a = 0; a < p; a++) t -= iAW[(size_t)a * n + i] * beta[(size_t)a];
    u[(size_t)i] = t;
    // row i of Z solves Z_i Lm' = Y_i  (forward substitution over the columns of Lm)
    for (int a = 0; a < pz; a++) {
      double zv = iAW[(size_t)a * n + i];
      for (int k = 0; k < a; k++) zv -= Lm[(size_t)k * p + a] * Zt[(size_t)k * n + i];
      Zt[(size_t)a * n + i] = zv / Lm[(size_t)a * p + a];
    }
  }
  const auto t4 = tnow();
  const auto t5 = t4;
  // ---- distinct tables and the per-set coefficients ----
  GradArgs g;
  memset(&g, 0, sizeof(g));
  auto findT = [&](const double* ptr) -> int {
    if (ptr == nullptr) return -1;
    for (int k = 0; k < g.nTu; k++) if (g.Tu[k] == ptr) return k;
    if (g.nTu >= GR_MAXT) return -2;
    g.Tu[g.nTu] = ptr; return g.nTu++;
  };
  auto findP = [&](const double* ptr) -> int {
    if (ptr == nullptr) return -1;
    for (int k = 0; k < g.nPu; k++) if (g.Pu[k] == ptr) return k;
    if (g.nPu >= GR_MAXP) return -2;
    g.Pu[g.nPu] = ptr; return g.nPu++;
  };
  std::vector<int> setpar;                                              // parameter index of set s >= 1
  bool overflow = false;
  for (int k = 0; k <= npar; k++) {
    if (plans[(size_t)k].status != 0) continue;
    const CovArgs& a = cargs[(size_t)k];
    GradSet S;
    S.itd = a.tidx[0] >= 0 ? findT(a.T[a.tidx[0]]) : -1;
    S.it0 = a.tidx[1] >= 0 ? findT(a.T[a.tidx[1]]) : -1;
    S.it1 = a.tidx[2] >= 0 ? findT(a.T[a.tidx[2]]) : -1;
    S.ipx = findP(a.phix);
    if (S.itd == -2 || S.it0 == -2 || S.it1 == -2 || S.ipx == -2) { overflow = true; break; }
    if (S.it1 < 0) { S.it1 = 0; }                                       // never read: ipx < 0 or cxd1 * 0
    S.cx0 = a.cx0; S.cx1 = a.cx1; S.kd1 = a.kd1; S.cxd0 = a.cxd0; S.cxd1 = (a.tidx[2] >= 0) ? a.cxd1 : 0.0; S.cme = a.cme;
    if (S.it0 < 0) { S.it0 = 0; S.cxd0 = 0.0; }
    S.inv_delta = (k == 0) ? 0.0 : 1.0 / delta[k - 1];
    g.set[g.nset++] = S;
    if (k > 0) setpar.push_back(k - 1);
  }
  if (overflow) { ctx->err = "nll_grad: more distinct tables than the pass holds"; return IAK3D_ERR_ARG; }
  double *dZt = nullptr, *du = nullptr, *dpart = nullptr;
  const int nRB = (int)((n + TM - 1) / TM), nCB = (int)((n + TN - 1) / TN);
  int64_t tiles = 0;
  for (int c = 0; c < nCB; c++) tiles += nRB - c / 2;
  do {
    if (g.nset <= 1) break;
    if (iak_pool_alloc(ctx, (void**)&dZt, Zt.size() * 8) != cudaSuccess || iak_pool_alloc(ctx, (void**)&du, (size_t)n * 8) != cudaSuccess ||
        iak_pool_alloc(ctx, (void**)&dpart, (size_t)tiles * (GR_MAXSET - 1) * 8) != cudaSuccess) { ctx->err = "nll_grad: allocation failed"; rc = IAK3D_ERR_CUDA; cudaGetLastError(); break; }
    cudaMemcpyAsync(dZt, Zt.data(), Zt.size() * 8, cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync(du, u.data(), (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream);
    g.xid = ds->d_xid; g.did = ds->d_did; g.n = n; g.nd = ds->ndIU; g.nx = ds->nxU;
    g.iC = iCb; g.nRU = rowsU / UR; g.Zt = dZt; g.pz = pz; g.u = du; g.invc = 1.0 / cxd; g.nRB = nRB;
    grad_pass_kernel<<<(unsigned)tiles, 256, 0, ctx->stream>>>(g, dpart);
    ctx->launches++;
    std::vector<double> part((size_t)tiles * (GR_MAXSET - 1));
    cudaMemcpyAsync(part.data(), dpart, part.size() * 8, cudaMemcpyDeviceToHost, ctx->stream);
    const cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { ctx->err = std::string("nll_grad: ") + cudaGetErrorString(e); rc = IAK3D_ERR_CUDA; break; }
    for (int sidx = 1; sidx < g.nset; sidx++) {
      double t = 0.0;
      for (int64_t b2 = 0; b2 < tiles; b2++) t += part[(size_t)b2 * (GR_MAXSET - 1) + sidx - 1];   // fixed order
      gsum[setpar[(size_t)sidx - 1]] = t;
    }
  } while (0);
  const auto t6 = tnow();
  iak_pool_free(ctx, dZt); iak_pool_free(ctx, du); iak_pool_free(ctx, dpart);
  if (prof) fprintf(stderr, "[gradprof] n=%lld: tables %.2f, build+factorise %.2f, inverse %.2f, iC W + host algebra %.2f, pass %.2f, free %.2f ms\n",
                    (long long)n, ms(t0, t1), ms(t1, t2), ms(t2, t3), ms(t3, t4), ms(t5, t6), ms(t6, tnow()));
  return rc;
}
