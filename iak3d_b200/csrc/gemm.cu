// gemm.cu -- dense FP64 products behind the explicit inverse (chol2inv / solve(C): optim_functions.R:297,
// fitIAK3D.R:931) and the Newton-Raphson terms on log-variances (gradHessIAK3D, nrUpdatesIAK3D.R:118-216).
//
// Every O(n^3) product runs through the tile engine of chol.cu (DMMA K-loop, TMA-staged blocked operands):
//   U  = L^-T                 mode 1 with an identity panel
//   iC = U U'                 mode 2 (plain product  Out = use_cin * Out + alpha * A B')
//   T  = iC - (iCX V) iCX'    mode 2, K = 32 (p padded)
//   M_i = T dC_i              mode 2 (dC_i symmetric, so T dC_i = T dC_i')
// The O(n^2) pieces (traces, tr(M_i M_j), matrix-vector products) are plain kernels with fixed-order reductions.
#include <math.h>
#include <string.h>

#include <algorithm>

#include "common.cuh"

// from api.cu
struct TableCache;
int iak_make_plan(iak3d_ctx* ctx, const iak3d_pars* p, bool square, EvalPlan* pl);
int iak_plan_check_avsd(iak3d_ctx* ctx, const iak3d_ds* ds, EvalPlan* pl);
int iak_alloc_slot(iak3d_ctx* ctx, int64_t n, int32_t q, int64_t nxU, int64_t ndIU, Slot** out, bool tables_only = false);
void iak_free_slot(Slot* s);
int iak_square_tables(iak3d_ctx* ctx, const iak3d_ds* ds, const EvalPlan& pl, Slot* s, CovArgs* a, TableCache* cache = nullptr, TableQueue* tq = nullptr);
int iak_factor_slot(iak3d_ctx* ctx, Slot* s, int32_t* status_out, double* lndet_out, double* G_out);
int iak_backsolve_to_host(iak3d_ctx* ctx, Slot* s, double* out_host, double* d_out_opt);

namespace {

__global__ void identity_blocked_kernel(double* __restrict__ P, int64_t nRU, int64_t rows, int64_t cols) {
  // one thread per (unit column c, row i): zero, except the diagonal
  const int64_t c = blockIdx.y;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < rows; i += (int64_t)gridDim.x * blockDim.x)
    P[uidx(i, c, nRU)] = (i == c && c < cols) ? 1.0 : 0.0;
}
__global__ void unblock_kernel(const double* __restrict__ B, int64_t nRU, int64_t n, double* __restrict__ out) {
  const int64_t c = blockIdx.y;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[c * n + i] = B[uidx(i, c, nRU)];
}
// column-major n x k (k <= 32) -> blocked rows x 32 panel, zero padded
__global__ void block_panel_kernel(const double* __restrict__ src, int64_t n, int k, double* __restrict__ dst, int64_t nRU, int64_t rows) {
  const int c = blockIdx.y;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < rows; i += (int64_t)gridDim.x * blockDim.x)
    dst[uidx(i, c, nRU)] = (i < n && c < k) ? src[(int64_t)c * n + i] : 0.0;
}
// y = M x over the leading n x n part (thread per row, columns in order: deterministic)
__global__ void matvec_rows_kernel(const double* __restrict__ M, int64_t nRU, int64_t n, const double* __restrict__ x, double* __restrict__ y) {
  const int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n) return;
  double s = 0.0;
  for (int64_t b = 0; b < n; b++) s = fma(M[uidx(a, b, nRU)], __ldg(x + b), s);
  y[a] = s;
}
// Sharding of the Newton-Raphson trace terms over GPUs: the n x n products M_i = T dC_i are cut into 128 x 128
// super-tiles; the unordered pair {R, C} of super-tile indices -- i.e. tile (R, C) together with its transpose position (C, R),
// which tr(M_i M_j) = sum_ab M_i[a,b] M_j[b,a] needs side by side -- belongs to rank (k mod world), k its index in the packed
// lower triangle.  A rank computes and reads only its own tiles; the partial traces are summed by one all-reduce.
__host__ __device__ __forceinline__ int tile_owner(int64_t R, int64_t Cc, int world) {
  const int64_t hi = R > Cc ? R : Cc, lo = R > Cc ? Cc : R;
  return (int)((hi * (hi + 1) / 2 + lo) % world);
}
// partial[blockIdx.x] = sum over a in this block's 32 rows, all b < n in super-tiles owned by `rank`, of Mi(a,b) * Mj(b,a)
// (Mj == nullptr: trace of Mi, diagonal super-tiles only)
__global__ void __launch_bounds__(256) trace_pair_kernel(const double* __restrict__ Mi, const double* __restrict__ Mj, int64_t nRU,
                                                         int64_t n, double* __restrict__ partial, int rank, int world) {
  __shared__ double sj[32][33];
  __shared__ double red[256];
  const int64_t a0 = (int64_t)blockIdx.x * 32;
  const int tf = threadIdx.x & 31, ts = threadIdx.x >> 5;      // fast index 0..31, slow index 0..7
  double acc = 0.0;
  if (Mj == nullptr) {
    if (threadIdx.x < 32 && tile_owner(a0 / TM, a0 / TM, world) == rank) { const int64_t a = a0 + threadIdx.x; if (a < n) acc = Mi[uidx(a, a, nRU)]; }
  } else {
    for (int64_t b0 = 0; b0 < n; b0 += 32) {
      if (tile_owner(a0 / TM, b0 / TM, world) != rank) continue;      // uniform over the block
      // Mj block rows b0.. (fast), cols a0.. (slow) -> shared, then read transposed
      for (int s = ts; s < 32; s += 8) {
        const int64_t b = b0 + tf, a = a0 + s;
        sj[s][tf] = (b < n && a < n) ? Mj[uidx(b, a, nRU)] : 0.0;
      }
      __syncthreads();
      for (int s = ts; s < 32; s += 8) {
        const int64_t a = a0 + tf, b = b0 + s;
        if (a < n && b < n) acc = fma(Mi[uidx(a, b, nRU)], sj[tf][s], acc);
      }
      __syncthreads();
    }
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o]; __syncthreads(); }
  if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}

struct DevBuf {                       // stream-ordered scratch (iak_pool_alloc); a kept buffer (ctx == nullptr) is not released
  void* p = nullptr;
  iak3d_ctx* ctx = nullptr;
  ~DevBuf() { if (p && ctx) iak_pool_free(ctx, p); }
  cudaError_t alloc(iak3d_ctx* c, size_t bytes) { ctx = c; return iak_pool_alloc(c, &p, bytes); }
  template <class T> T* as() { return (T*)p; }
};

// (r, j) tiles of a rows x cols output, column-major task order (mode 1 needs it; mode 2 does not care).
// tri = 0: all tiles, K-loop from 0.  tri = 1 (mode 1, identity panel: U = L^-T is upper triangular): only the tiles with
// j >= 2r, K-loop from the row block's own first column (X(r,k) = 0 before it).  tri = 2 (mode 2, U U' with U upper
// triangular, lower triangle of the symmetric result): only the tiles with r >= j/2, K-loop from the row block's first column.
int run_tiles(iak3d_ctx* ctx, const ProblemDesc& pd, int nRB, int nCB, int tri = 0, int rank = 0, int world = 1) {
  std::vector<int4> tasks;
  tasks.reserve((size_t)nRB * nCB);
  for (int j = 0; j < nCB; j++)
    for (int r = 0; r < nRB; r++) {
      if (tri == 1 && j < 2 * r) continue;
      if (tri == 2 && r < j / 2) continue;
      if (world > 1 && tile_owner(r, j / 2, world) != rank) continue;      // 128 x 64 task (r, j) lies in super-tile (r, j / 2)
      tasks.push_back(make_int4(0, r, j, tri ? r * (TM / UC) : 0));
    }
  if (tasks.empty()) return 0;
  DevBuf d_tasks, d_prob, d_ticket;
  CUDA_TRY(ctx, d_tasks.alloc(ctx, tasks.size() * sizeof(int4)));
  CUDA_TRY(ctx, d_prob.alloc(ctx, sizeof(ProblemDesc)));
  CUDA_TRY(ctx, d_ticket.alloc(ctx, 2 * sizeof(int32_t)));
  CUDA_TRY(ctx, cudaMemcpyAsync(d_tasks.p, tasks.data(), tasks.size() * sizeof(int4), cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(ctx, cudaMemcpyAsync(d_prob.p, &pd, sizeof(pd), cudaMemcpyHostToDevice, ctx->stream));
  const int st = launch_tile_tasks(ctx, d_prob.as<ProblemDesc>(), d_tasks.as<int4>(), (int32_t)tasks.size(), d_ticket.as<int32_t>(), pd.mode, 0,
                                   pd.mode == 2 ? (int)std::min<size_t>(tasks.size(), 1u << 30) : nRB);
  if (st != 0) return st;
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}

// Out(rowsOut x colsOut, blocked) = use_cin * Out + alpha * A(rowsOut x K) B(colsOut x K)'   (all blocked, K multiple of 32)
int gemm_nt(iak3d_ctx* ctx, const double* A, int64_t rowsA, const double* B, int64_t rowsB, double* Out, int64_t rowsOut,
            int64_t colsOut, int64_t K, double alpha, int use_cin, int32_t* d_status, int tri = 0, int rank = 0, int world = 1) {
  ProblemDesc pd;
  memset(&pd, 0, sizeof(pd));
  pd.L = const_cast<double*>(B); pd.nRUL = rowsB / UR;
  pd.P = Out; pd.nRUP = rowsOut / UR;
  pd.A = A; pd.nRUA = rowsA / UR;
  pd.nCB = (int32_t)(colsOut / TN); pd.nRB = (int32_t)(rowsOut / TM);
  pd.mode = 2; pd.kchunks = (int32_t)(K / UC); pd.use_cin = use_cin; pd.alpha = alpha;
  pd.status = d_status;
  return run_tiles(ctx, pd, pd.nRB, pd.nCB, tri, rank, world);
}

// upper triangle := transpose of the lower triangle (blocked layout)
__global__ void mirror_lower_kernel(double* __restrict__ M, int64_t nRU, int64_t n) {
  const int64_t c = blockIdx.y;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < c; i += (int64_t)gridDim.x * blockDim.x)
    M[uidx(i, c, nRU)] = M[uidx(c, i, nRU)];
}

}  // namespace

// mat.cu: the plain product on caller-owned blocked matrices
int iak_gemm_nt(iak3d_ctx* ctx, const double* A, int64_t rowsA, const double* B, int64_t rowsB, double* Out, int64_t rowsOut,
                int64_t colsOut, int64_t K, double alpha, int use_cin, int32_t* d_status) {
  return gemm_nt(ctx, A, rowsA, B, rowsB, Out, rowsOut, colsOut, K, alpha, use_cin, d_status, 0);
}

// out (n x q, column-major) = M W for the leading n x n part of a blocked (full, symmetric) matrix.  CTA = 128 rows x a
// chunk of PP_CH columns (the chunk of W is staged in shared memory and read as a broadcast); the chunks' partial
// sums are added in chunk order by a second kernel: deterministic.
constexpr int PP_CH = 256;
__global__ void __launch_bounds__(128) panel_product_kernel(const double* __restrict__ M, int64_t nRU, int64_t n, const double* __restrict__ W,
                                                            int q, double* __restrict__ part /* [chunk][q][n] */) {
  extern __shared__ double sw[];                     // PP_CH x q (column c of the chunk: sw[c * q + w])
  const int64_t c0 = (int64_t)blockIdx.y * PP_CH;
  const int cn = (int)((n - c0) < PP_CH ? (n - c0) : PP_CH);
  for (int idx = threadIdx.x; idx < cn * q; idx += blockDim.x) { const int w = idx / cn, c = idx % cn; sw[c * q + w] = W[(int64_t)w * n + c0 + c]; }
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int w0 = 0; w0 < q; w0 += 16) {
    double acc[16];
#pragma unroll
    for (int w = 0; w < 16; w++) acc[w] = 0.0;
    for (int c = 0; c < cn; c++) {
      const double m = M[uidx(i, c0 + c, nRU)];
#pragma unroll
      for (int w = 0; w < 16; w++) if (w0 + w < q) acc[w] = fma(m, sw[c * q + w0 + w], acc[w]);
    }
#pragma unroll
    for (int w = 0; w < 16; w++) if (w0 + w < q) part[((int64_t)blockIdx.y * q + w0 + w) * n + i] = acc[w];
  }
}
__global__ void panel_reduce_kernel(const double* __restrict__ part, int nchunk, int64_t nq, double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq) return;
  double s = 0.0;
  for (int k = 0; k < nchunk; k++) s += part[(int64_t)k * nq + i];
  out[i] = s;
}
int launch_panel_product(iak3d_ctx* ctx, const double* d_M, int64_t nRU, int64_t n, const double* d_W, int q, double* d_out) {
  if (n == 0 || q == 0) return 0;
  const int nchunk = (int)((n + PP_CH - 1) / PP_CH);
  double* d_part = nullptr;
  if (iak_pool_alloc(ctx, (void**)&d_part, (size_t)nchunk * q * n * sizeof(double)) != cudaSuccess) { ctx->err = "panel product: allocation failed"; cudaGetLastError(); return IAK3D_ERR_CUDA; }
  panel_product_kernel<<<dim3((unsigned)((n + 127) / 128), (unsigned)nchunk), 128, (size_t)PP_CH * q * sizeof(double), ctx->stream>>>(d_M, nRU, n, d_W, q, d_part);
  panel_reduce_kernel<<<(unsigned)(((int64_t)n * q + 255) / 256), 256, 0, ctx->stream>>>(d_part, nchunk, (int64_t)n * q, d_out);
  ctx->launches += 2;
  const cudaError_t e = cudaStreamSynchronize(ctx->stream);
  iak_pool_free(ctx, d_part);
  CUDA_TRY(ctx, e);
  return 0;
}

// iC (blocked, rowsU x Npad with rowsU = round_up(Npad, 128)) from a finished factor.  With U_keep / iC_keep (both of
// ubuf_doubles(rowsU, Npad) doubles, owned by the caller) nothing is allocated for the two big buffers and *out = iC_keep;
// otherwise the caller frees *out with cudaFree.
int iak_inverse_blocked(iak3d_ctx* ctx, Slot* s, double** out, int64_t* rowsU_out, double* U_keep, double* iC_keep) {
  const int64_t Npad = s->Npad, rowsU = round_up(Npad, TM), nRU = rowsU / UR;
  const size_t bytes = (size_t)ubuf_doubles(rowsU, Npad) * sizeof(double);
  DevBuf U, G, flags, status;
  double* iC = iC_keep;
  if (U_keep != nullptr) { U.p = U_keep; }
  else CUDA_TRY(ctx, U.alloc(ctx, bytes));
  struct Release { DevBuf& b; bool keep; ~Release() { if (keep) b.p = nullptr; } } rel{U, U_keep != nullptr};
  CUDA_TRY(ctx, G.alloc(ctx, (size_t)rowsU * sizeof(double)));
  const size_t nfl = (size_t)(rowsU / TM) * s->nCB;
  CUDA_TRY(ctx, flags.alloc(ctx, nfl * sizeof(int32_t)));
  CUDA_TRY(ctx, status.alloc(ctx, 2 * sizeof(int32_t)));
  CUDA_TRY(ctx, cudaMemsetAsync(flags.p, 0, nfl * sizeof(int32_t), ctx->stream));
  CUDA_TRY(ctx, cudaMemsetAsync(status.p, 0, 2 * sizeof(int32_t), ctx->stream));
  {
    dim3 grid((unsigned)((rowsU + 255) / 256), (unsigned)Npad);
    if (Npad > 65535) { ctx->err = "inverse: more than 65535 columns"; return IAK3D_ERR_ARG; }
    identity_blocked_kernel<<<grid, 256, 0, ctx->stream>>>(U.as<double>(), nRU, rowsU, Npad);
    ctx->launches++;
  }
  // U = I L^-T  (mode 1 against the finished factor)
  ProblemDesc pd;
  memset(&pd, 0, sizeof(pd));
  pd.L = s->d_L; pd.nRUL = s->ld / UR; pd.P = U.as<double>(); pd.nRUP = nRU; pd.invd = s->d_invd;
  pd.doneP = flags.as<int32_t>(); pd.nCB = s->nCB; pd.nRB = (int32_t)(rowsU / TM); pd.mode = 1; pd.epoch = 1;
  pd.w0 = (int32_t)s->Npad; pd.q = 0; pd.G = G.as<double>(); pd.ldG = rowsU; pd.status = status.as<int32_t>();
  // U is upper triangular: only its non-zero tiles are computed, each K-loop starts at the tile's own row block
  // (n^3/3 flops instead of n^3); the tiles below stay as the identity panel initialised them (zero)
  int rc = run_tiles(ctx, pd, pd.nRB, pd.nCB, 1);
  if (rc != 0) return rc;
  // iC = U U': lower triangle only, K from the row block on (n^3/3 flops instead of 2 n^3), then mirrored
  if (iC == nullptr && iak_malloc(ctx, (void**)&iC, bytes) != cudaSuccess) { ctx->err = "inverse: allocation failed"; cudaGetLastError(); return IAK3D_ERR_CUDA; }
  rc = gemm_nt(ctx, U.as<double>(), rowsU, U.as<double>(), rowsU, iC, rowsU, Npad, Npad, 1.0, 0, status.as<int32_t>(), 2);
  if (rc != 0) { if (iC_keep == nullptr) cudaFree(iC); return rc; }
  {
    dim3 grid((unsigned)std::min<int64_t>((Npad + 255) / 256, 64), (unsigned)Npad);
    mirror_lower_kernel<<<grid, 256, 0, ctx->stream>>>(iC, nRU, Npad);
    ctx->launches++;
  }
  int32_t hs[2] = {0, 0};
  CUDA_TRY(ctx, cudaMemcpy(hs, status.p, sizeof(hs), cudaMemcpyDeviceToHost));
  if (hs[1] != 0) { if (iC_keep == nullptr) cudaFree(iC); ctx->err = "inverse: dependency wait timed out (internal error)"; return IAK3D_ERR_TIMEOUT; }
  *out = iC; *rowsU_out = rowsU;
  return 0;
}

int iak_inverse_from_factor(iak3d_ctx* ctx, Slot* s, double* d_iC /* n x n column-major */) {
  double* iCb = nullptr; int64_t rowsU = 0;
  int rc = iak_inverse_blocked(ctx, s, &iCb, &rowsU, nullptr, nullptr);
  if (rc != 0) return rc;
  dim3 grid((unsigned)((s->n + 255) / 256), (unsigned)s->n);
  unblock_kernel<<<grid, 256, 0, ctx->stream>>>(iCb, rowsU / UR, s->n, d_iC);
  ctx->launches++;
  const cudaError_t e = cudaStreamSynchronize(ctx->stream);
  cudaFree(iCb);
  CUDA_TRY(ctx, e);
  return 0;
}

// ---- small dense helpers on the host (p x p, p <= 63) -------------------------------------------------
static bool host_chol(int p, const double* A /* col-major */, double* L) {
  for (int j = 0; j < p; j++) {
    for (int i = j; i < p; i++) {
      double s = A[(size_t)j * p + i];
      for (int k = 0; k < j; k++) s -= L[(size_t)k * p + i] * L[(size_t)k * p + j];
      if (i == j) { if (!(s > 0.0) || !isfinite(s)) return false; L[(size_t)j * p + j] = sqrt(s); }
      else L[(size_t)j * p + i] = s / L[(size_t)j * p + j];
    }
    for (int i = 0; i < j; i++) L[(size_t)j * p + i] = 0.0;
  }
  return true;
}
static void host_chol_solve(int p, const double* L, const double* b, double* x) {   // (L L') x = b
  std::vector<double> y((size_t)p);
  for (int i = 0; i < p; i++) { double s = b[i]; for (int k = 0; k < i; k++) s -= L[(size_t)k * p + i] * y[k]; y[i] = s / L[(size_t)i * p + i]; }
  for (int i = p - 1; i >= 0; i--) { double s = y[i]; for (int k = i + 1; k < p; k++) s -= L[(size_t)i * p + k] * x[k]; x[i] = s / L[(size_t)i * p + i]; }
}

// The rank's share of gradHessIAK3D: everything O(n^2) and the factorisation / inverse are computed by every rank (T must
// be whole wherever a tile of T dC_i is formed); the six n x n x n products and the 21 pair traces -- 12 n^3 of the 13 n^3
// flops -- are cut by super-tile ownership.  tr_part = { this rank's part of tr(M_i) (m), of tr(M_i M_j) (m x m, both
// triangles filled) }: summed over the ranks (one all-reduce of m + m^2 doubles) they give the traces of :197-210.
extern "C" int iak3d_gradhess_shard(iak3d_ds* ds, const iak3d_pars* pars, int32_t ncomp, const double* lnvPars, int32_t rank, int32_t world,
                                    double* nll, double* betahat, double* vbetahat, double* zTdCTz_out, double* qij_out, double* tr_part) {
  if (!ds || !pars || !lnvPars || !nll || !zTdCTz_out || !qij_out || !tr_part || world < 1 || rank < 0 || rank >= world) return IAK3D_ERR_ARG;
  iak3d_ctx* ctx = ds->ctx;
  if (ncomp != 5 && ncomp != 6) { ctx->err = "gradhess: ncomp must be 5 or 6 (cx0, cx1, cd1, cxd0, cxd1[, cme])"; return IAK3D_ERR_ARG; }
  if (ds->q < 2) { ctx->err = "gradhess: the dataset needs W = [X z]"; return IAK3D_ERR_ARG; }
  cudaSetDevice(ctx->device);
  const int64_t n = ds->n;
  const int q = ds->q, p = q - 1, m = ncomp;
  // C = sum_i exp(theta_i) dA_i  (nrUpdatesIAK3D.R:130-138): setCIAK3D with these variances, no quirk
  iak3d_pars P = *pars;
  double var[6] = {0, 0, 0, 0, 0, 0};
  for (int i = 0; i < m; i++) var[i] = exp(lnvPars[i]);
  P.cx0 = var[0]; P.cx1 = var[1]; P.cd1 = var[2]; P.cxd0 = var[3]; P.cxd1 = var[4]; P.cme = var[5];
  P.quirk_cd1_unscaled = 0;
  P.nssre = 0;                 // nrUpdatesIAK3D's components are the six terms of setCIAK3D (nrUpdatesIAK3D.R:130-138)
  EvalPlan pl;
  int rc = iak_make_plan(ctx, &P, true, &pl);
  if (rc == 0) rc = iak_plan_check_avsd(ctx, ds, &pl);
  if (rc != 0) return rc;
  if (pl.status != 0) return pl.status;
  const double nanv = nan("");
  Slot* s = nullptr;
  rc = iak_alloc_slot(ctx, n, q, ds->nxU, ds->ndIU, &s);
  if (rc != 0) return rc;
  double* iCb = nullptr;
  std::vector<double*> Mi((size_t)m, nullptr);
  do {
    CovArgs a;
    rc = iak_square_tables(ctx, ds, pl, s, &a);
    if (rc != 0) break;
    rc = launch_cov_factor_layout(ctx, a, ds->d_W, q, s->n, s->Npad, s->ld, s->nCB, s->d_L);
    if (rc != 0) break;
    int32_t status = 0; double lndetC = 0.0;
    std::vector<double> G((size_t)q * q);
    rc = iak_factor_slot(ctx, s, &status, &lndetC, G.data());
    if (rc != 0) break;
    if (status != 0) { *nll = nanv; rc = status; break; }
    std::vector<double> iCW((size_t)n * q);
    rc = iak_backsolve_to_host(ctx, s, iCW.data(), nullptr);
    if (rc != 0) break;
    // ---- p x p algebra on the host (nrUpdatesIAK3D.R:160-190) ----
    std::vector<double> XiCX((size_t)p * p), XiCz((size_t)p), Lp((size_t)p * p), beta((size_t)p), iXiCX((size_t)p * p);
    for (int c = 0; c < p; c++) for (int r = 0; r < p; r++) XiCX[(size_t)c * p + r] = 0.5 * (G[(size_t)c * q + r] + G[(size_t)r * q + c]);
    for (int r = 0; r < p; r++) XiCz[r] = 0.5 * (G[(size_t)p * q + r] + G[(size_t)r * q + p]);
    const double ziCz = G[(size_t)p * q + p];
    if (!host_chol(p, XiCX.data(), Lp.data())) { *nll = nanv; rc = IAK3D_NOT_SPD; break; }
    host_chol_solve(p, Lp.data(), XiCz.data(), beta.data());
    for (int c = 0; c < p; c++) { std::vector<double> e((size_t)p, 0.0); e[c] = 1.0; host_chol_solve(p, Lp.data(), e.data(), &iXiCX[(size_t)c * p]); }
    double lndetXiCX = 0.0;
    for (int i = 0; i < p; i++) lndetXiCX += 2.0 * log(Lp[(size_t)i * p + i]);
    double bXz = 0.0, bXXb = 0.0;
    for (int r = 0; r < p; r++) { bXz += beta[r] * XiCz[r]; double t = 0.0; for (int c = 0; c < p; c++) t += XiCX[(size_t)c * p + r] * beta[c]; bXXb += beta[r] * t; }
    const double zTz = ziCz - 2.0 * bXz + bXXb;
    std::vector<double> Tz((size_t)n), A1((size_t)n * p);
    for (int64_t i = 0; i < n; i++) {
      double t = iCW[(size_t)p * n + i];
      for (int c = 0; c < p; c++) t -= iCW[(size_t)c * n + i] * beta[c];
      Tz[(size_t)i] = t;
      for (int c = 0; c < p; c++) { double v = 0.0; for (int k = 0; k < p; k++) v += iCW[(size_t)k * n + i] * iXiCX[(size_t)c * p + k]; A1[(size_t)c * n + i] = v; }
    }
    // ---- T = iC - iCX (X'iCX)^-1 iCX'  (:183) ----
    int64_t rowsU = 0;
    rc = iak_inverse_blocked(ctx, s, &iCb, &rowsU, nullptr, nullptr);
    if (rc != 0) break;
    const int64_t nRU = rowsU / UR, Npad = s->Npad;
    const size_t mbytes = (size_t)ubuf_doubles(rowsU, Npad) * sizeof(double);
    DevBuf dA1, dB1, pA, pB, dTz, du, dv, dpart, dstatus, dC;
    const size_t pbytes = (size_t)ubuf_doubles(rowsU, TN) * sizeof(double);     // rowsU x 64 panel (K = 32 uses the first unit column)
    cudaError_t e = dA1.alloc(ctx, (size_t)n * p * 8);
    if (e == cudaSuccess) e = dB1.alloc(ctx, (size_t)n * p * 8);
    if (e == cudaSuccess) e = pA.alloc(ctx, pbytes);
    if (e == cudaSuccess) e = pB.alloc(ctx, pbytes);
    if (e == cudaSuccess) e = dTz.alloc(ctx, (size_t)n * 8);
    if (e == cudaSuccess) e = du.alloc(ctx, (size_t)n * 8);
    if (e == cudaSuccess) e = dv.alloc(ctx, (size_t)n * 8);
    const int nblk = (int)((n + 31) / 32);
    if (e == cudaSuccess) e = dpart.alloc(ctx, (size_t)nblk * 8);
    if (e == cudaSuccess) e = dstatus.alloc(ctx, 2 * sizeof(int32_t));
    if (e == cudaSuccess) e = dC.alloc(ctx, mbytes);
    if (e != cudaSuccess) { ctx->err = std::string("gradhess: allocation failed: ") + cudaGetErrorString(e); cudaGetLastError(); rc = IAK3D_ERR_CUDA; break; }
    cudaMemsetAsync(dstatus.p, 0, 2 * sizeof(int32_t), ctx->stream);
    cudaMemcpyAsync(dA1.p, A1.data(), (size_t)n * p * 8, cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync(dB1.p, iCW.data(), (size_t)n * p * 8, cudaMemcpyHostToDevice, ctx->stream);      // iCX = first p columns
    cudaMemcpyAsync(dTz.p, Tz.data(), (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream);
    {
      dim3 grid((unsigned)((rowsU + 255) / 256), TN);
      block_panel_kernel<<<grid, 256, 0, ctx->stream>>>(dA1.as<double>(), n, p, pA.as<double>(), nRU, rowsU);
      block_panel_kernel<<<grid, 256, 0, ctx->stream>>>(dB1.as<double>(), n, p, pB.as<double>(), nRU, rowsU);
      ctx->launches += 2;
    }
    if (p > 32) {   // K = 64: both unit columns of the panels
      rc = gemm_nt(ctx, pA.as<double>(), rowsU, pB.as<double>(), rowsU, iCb, rowsU, Npad, 64, -1.0, 1, dstatus.as<int32_t>());
    } else {
      rc = gemm_nt(ctx, pA.as<double>(), rowsU, pB.as<double>(), rowsU, iCb, rowsU, Npad, 32, -1.0, 1, dstatus.as<int32_t>());
    }
    if (rc != 0) break;
    double* T = iCb;
    // ---- per component: dC_i, u_i = dC_i Tz, M_i = T dC_i, tr M_i, v_i = M_i Tz, r_i = M_i' z ----
    std::vector<double> trM((size_t)m, 0.0), zTdCTz((size_t)m, 0.0);
    std::vector<std::vector<double>> vv((size_t)m, std::vector<double>((size_t)n)), rr((size_t)m, std::vector<double>((size_t)n));
    std::vector<double> part((size_t)nblk), hu((size_t)n);
    const double cv[6][6] = {{1, 0, 0, 0, 0, 0}, {0, 1, 0, 0, 0, 0}, {0, 0, 1, 0, 0, 0}, {0, 0, 0, 1, 0, 0}, {0, 0, 0, 0, 1, 0}, {0, 0, 0, 0, 0, 1}};
    const unsigned rb = (unsigned)((n + 127) / 128);
    for (int i = 0; i < m && rc == 0; i++) {
      CovArgs ai = a;
      ai.cx0 = cv[i][0] * pl.cx0; ai.cx1 = cv[i][1] * pl.cx1; ai.kd1 = cv[i][2] * pl.cd1; ai.cxd0 = cv[i][3] * pl.cxd0;
      ai.cxd1 = cv[i][4] * pl.cxd1; ai.cme = cv[i][5] * pl.cme;
      rc = launch_cov_rect(ctx, ai, dC.as<double>(), rowsU, rowsU, Npad, 1, /*blocked=*/1);
      if (rc != 0) break;
      matvec_rows_kernel<<<rb, 128, 0, ctx->stream>>>(dC.as<double>(), nRU, n, dTz.as<double>(), du.as<double>());
      ctx->launches++;
      cudaMemcpyAsync(hu.data(), du.p, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream);
      if (iak_malloc(ctx, (void**)&Mi[i], mbytes) != cudaSuccess) { ctx->err = "gradhess: allocation failed (M_i)"; cudaGetLastError(); rc = IAK3D_ERR_CUDA; break; }
      // v_i = M_i Tz = T (dC_i Tz) = T u_i and r_i = M_i' z = dC_i (T z) = u_i: two matrix-vector products on whole matrices,
      // so that M_i itself is only needed where its traces are taken
      matvec_rows_kernel<<<rb, 128, 0, ctx->stream>>>(T, nRU, n, du.as<double>(), dv.as<double>());
      ctx->launches++;
      cudaMemcpyAsync(vv[i].data(), dv.p, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream);
      rc = gemm_nt(ctx, T, rowsU, dC.as<double>(), rowsU, Mi[i], rowsU, Npad, Npad, 1.0, 0, dstatus.as<int32_t>(), 0, rank, world);
      if (rc != 0) break;
      if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { ctx->err = "gradhess: kernel failure"; rc = IAK3D_ERR_CUDA; break; }   // hu, vv[i]
      double t = 0.0;
      for (int64_t k = 0; k < n; k++) t += Tz[(size_t)k] * hu[(size_t)k];
      zTdCTz[i] = t;
      rr[i] = hu;
      trace_pair_kernel<<<nblk, 256, 0, ctx->stream>>>(Mi[i], nullptr, nRU, n, dpart.as<double>(), rank, world);
      ctx->launches++;
      cudaMemcpyAsync(part.data(), dpart.p, (size_t)nblk * 8, cudaMemcpyDeviceToHost, ctx->stream);
      if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { ctx->err = "gradhess: kernel failure"; rc = IAK3D_ERR_CUDA; break; }
      t = 0.0;
      for (int k = 0; k < nblk; k++) t += part[(size_t)k];
      trM[i] = t;
    }
    if (rc != 0) break;
    // ---- this rank's share of the traces; the quadratic forms (:197-210) ----
    for (int i = 0; i < m; i++) { tr_part[i] = trM[i]; zTdCTz_out[i] = zTdCTz[i]; }
    for (int i = 0; i < m && rc == 0; i++)
      for (int j = i; j < m; j++) {
        trace_pair_kernel<<<nblk, 256, 0, ctx->stream>>>(Mi[i], Mi[j], nRU, n, dpart.as<double>(), rank, world);
        ctx->launches++;
        // the context stream is non-blocking: a plain cudaMemcpy would not wait for the kernel
        cudaMemcpyAsync(part.data(), dpart.p, (size_t)nblk * 8, cudaMemcpyDeviceToHost, ctx->stream);
        if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { ctx->err = "gradhess: kernel failure"; rc = IAK3D_ERR_CUDA; break; }
        double tr = 0.0;
        for (int k = 0; k < nblk; k++) tr += part[(size_t)k];
        double qij = 0.0;                                            // z' T dC_i T dC_j T z = (M_i' z) . (M_j Tz)
        for (int64_t k = 0; k < n; k++) qij += rr[i][(size_t)k] * vv[j][(size_t)k];
        qij_out[(size_t)i * m + j] = qij_out[(size_t)j * m + i] = qij;
        tr_part[m + (size_t)i * m + j] = tr_part[m + (size_t)j * m + i] = tr;
      }
    if (rc != 0) break;
    int32_t hs[2] = {0, 0};
    cudaMemcpy(hs, dstatus.p, sizeof(hs), cudaMemcpyDeviceToHost);
    if (hs[1] != 0) { ctx->err = "gradhess: dependency wait timed out (internal error)"; rc = IAK3D_ERR_TIMEOUT; break; }
    *nll = 0.5 * ((double)(n - p) * log(2.0 * 3.14159265358979323846) + lndetC + lndetXiCX + zTz);
    if (betahat) memcpy(betahat, beta.data(), (size_t)p * 8);
    if (vbetahat) memcpy(vbetahat, iXiCX.data(), (size_t)p * p * 8);
  } while (0);
  cudaStreamSynchronize(ctx->stream);
  for (double* mptr : Mi) if (mptr) cudaFree(mptr);
  if (iCb) cudaFree(iCb);
  iak_free_slot(s);
  return rc;
}

// gradHessIAK3D on one GPU: the whole share, then the assembly of nrUpdatesIAK3D.R:197-210
extern "C" int iak3d_gradhess(iak3d_ds* ds, const iak3d_pars* pars, int32_t ncomp, const double* lnvPars, double* nll, double* grad,
                              double* hess, double* fim, double* betahat, double* vbetahat) {
  if (!ds || !pars || !lnvPars || !nll || !grad || !hess || !fim) return IAK3D_ERR_ARG;
  if (ncomp != 5 && ncomp != 6) { ds->ctx->err = "gradhess: ncomp must be 5 or 6 (cx0, cx1, cd1, cxd0, cxd1[, cme])"; return IAK3D_ERR_ARG; }
  const int m = ncomp;
  double zq[6], qij[36], tr[6 + 36];
  const int rc = iak3d_gradhess_shard(ds, pars, ncomp, lnvPars, 0, 1, nll, betahat, vbetahat, zq, qij, tr);
  if (rc != 0) return rc;
  for (int i = 0; i < m; i++) grad[i] = 0.5 * (tr[i] - zq[i]);
  for (int i = 0; i < m; i++)
    for (int j = i; j < m; j++) {
      const double t = tr[m + i * m + j], q2 = qij[i * m + j];
      if (i == j) hess[(size_t)i * m + j] = 0.5 * (tr[i] - t - zq[i] + 2.0 * q2);
      else hess[(size_t)i * m + j] = hess[(size_t)j * m + i] = 0.5 * (-t + 2.0 * q2);
      fim[(size_t)i * m + j] = fim[(size_t)j * m + i] = 0.5 * t;
    }
  return IAK3D_OK;
}
