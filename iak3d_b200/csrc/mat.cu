// Device-resident dense matrices in the blocked layout of the tile engine (common.cuh: 64 x 32 units, column stride 68),
// for the callers of the hot path whose algebra is made of matrices as large as the data blocks: the Eidsvik
// composite-likelihood block predictor (compositeLikIAK3D.R:609-798) is a chain of products, SPD inverses and sub-block
// moves on (map supports + block data) x (block data) matrices.  Products run on the tile engine (mode 2, DMMA), SPD
// inverses are the factorisation + the triangular-aware explicit inverse of gemm.cu; everything else is a byte mover.
// Rows are padded to 128 and columns to 64 with zeros, and every operation keeps the padding zero.
#include <string.h>

#include <algorithm>
#include <vector>

#include "common.cuh"

int iak_make_plan(iak3d_ctx* ctx, const iak3d_pars* p, bool square, EvalPlan* pl);
int iak_plan_check_avsd(iak3d_ctx* ctx, const iak3d_ds* ds, EvalPlan* pl);
int iak_alloc_slot(iak3d_ctx* ctx, int64_t n, int32_t q, int64_t nxU, int64_t ndIU, Slot** out, bool tables_only = false);
void iak_free_slot(Slot* s);
int iak_get_slot(iak3d_ds* ds, size_t k, Slot** out);
struct TableCache;
int iak_square_tables(iak3d_ctx* ctx, const iak3d_ds* ds, const EvalPlan& pl, Slot* s, CovArgs* a, TableCache* cache = nullptr, TableQueue* tq = nullptr);
int iak_factor_slot(iak3d_ctx* ctx, Slot* s, int32_t* status_out, double* lndet_out, double* G_out);
int iak_inverse_blocked(iak3d_ctx* ctx, Slot* s, double** out, int64_t* rowsU_out, double* U_keep, double* iC_keep);
int iak_gemm_nt(iak3d_ctx* ctx, const double* A, int64_t rowsA, const double* B, int64_t rowsB, double* Out, int64_t rowsOut,
                int64_t colsOut, int64_t K, double alpha, int use_cin, int32_t* d_status);

struct iak3d_mat {
  iak3d_ctx* ctx = nullptr;
  int64_t rows = 0, cols = 0;      // logical shape
  int64_t prow = 0, pcol = 0;      // padded: multiples of 128 and 64
  double* d = nullptr;
  int64_t nRU() const { return prow / UR; }
  size_t bytes() const { return (size_t)ubuf_doubles(prow, pcol) * sizeof(double); }
};

namespace {

// host column-major (ld) <-> sub-block of a blocked matrix
__global__ void mat_set_kernel(double* __restrict__ M, int64_t nRU, int64_t r0, int64_t c0, int64_t nr, int64_t nc,
                               const double* __restrict__ src, int64_t ld) {
  const int64_t j = blockIdx.y;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nr; i += (int64_t)gridDim.x * blockDim.x)
    M[uidx(r0 + i, c0 + j, nRU)] = src[j * ld + i];
}
__global__ void mat_get_kernel(const double* __restrict__ M, int64_t nRU, int64_t r0, int64_t c0, int64_t nr, int64_t nc,
                               double* __restrict__ dst, int64_t ld) {
  const int64_t j = blockIdx.y;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nr; i += (int64_t)gridDim.x * blockDim.x)
    dst[j * ld + i] = M[uidx(r0 + i, c0 + j, nRU)];
}
// D[dr0+i, dc0+j] = beta * D[..] + alpha * S[sr0+i, sc0+j]   (transpose: S[sr0+j, sc0+i]), 32 x 32 tiles through shared
// memory so that both sides are read and written along their rows
__global__ void __launch_bounds__(256) mat_copy_kernel(double* __restrict__ D, int64_t nRUd, int64_t dr0, int64_t dc0,
                                                       const double* __restrict__ S, int64_t nRUs, int64_t sr0, int64_t sc0,
                                                       int64_t nr, int64_t nc, int transpose, double alpha, double beta) {
  __shared__ double t[32][33];
  const int64_t i0 = (int64_t)blockIdx.x * 32, j0 = (int64_t)blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  if (!transpose) {
    for (int jj = ty; jj < 32; jj += 8) {
      const int64_t i = i0 + tx, j = j0 + jj;
      if (i < nr && j < nc) {
        double* p = D + uidx(dr0 + i, dc0 + j, nRUd);
        const double v = alpha * S[uidx(sr0 + i, sc0 + j, nRUs)];
        *p = (beta == 0.0) ? v : beta * (*p) + v;
      }
    }
    return;
  }
  // source tile: rows j0.. (of S) x cols i0..; read along S's rows (tx = row of S)
  for (int cc = ty; cc < 32; cc += 8) {
    const int64_t sj = j0 + tx, si = i0 + cc;            // S[sr0 + sj, sc0 + si] = D element (i = si, j = sj)
    t[cc][tx] = (sj < nc && si < nr) ? S[uidx(sr0 + sj, sc0 + si, nRUs)] : 0.0;
  }
  __syncthreads();
  for (int jj = ty; jj < 32; jj += 8) {
    const int64_t i = i0 + tx, j = j0 + jj;
    if (i < nr && j < nc) {
      double* p = D + uidx(dr0 + i, dc0 + j, nRUd);
      const double v = alpha * t[tx][jj];
      *p = (beta == 0.0) ? v : beta * (*p) + v;
    }
  }
}
// factor buffer (same geometry as the matrix when q = 0): lower triangle of M, identity on the padding
__global__ void mat_pack_factor_kernel(const double* __restrict__ M, double* __restrict__ L, int64_t nRU, int64_t n, int64_t prow) {
  const int64_t c = blockIdx.y;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < prow; i += (int64_t)gridDim.x * blockDim.x) {
    double v;
    if (c < n) v = (i < n && i >= c) ? M[uidx(i, c, nRU)] : 0.0;
    else v = (i == c) ? 1.0 : 0.0;
    L[uidx(i, c, nRU)] = v;
  }
}
__global__ void mat_zero_pad_diag_kernel(double* __restrict__ M, int64_t nRU, int64_t n, int64_t pcol) {
  const int64_t c = n + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c < pcol) M[uidx(c, c, nRU)] = 0.0;
}
// out[j] = sum_i A[i,j] B[i,j]: one warp per column, lanes stride the rows, fixed shuffle tree (deterministic)
__global__ void mat_coldot_kernel(const double* __restrict__ A, int64_t nRUa, const double* __restrict__ B, int64_t nRUb, int64_t nr,
                                  int64_t nc, double* __restrict__ out) {
  const int64_t j = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (j >= nc) return;
  double s = 0.0;
  for (int64_t i = lane; i < nr; i += 32) s = fma(A[uidx(i, j, nRUa)], B[uidx(i, j, nRUb)], s);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if (lane == 0) out[j] = s;
}

bool in_range(const iak3d_mat* M, int64_t r0, int64_t c0, int64_t nr, int64_t nc) {
  return r0 >= 0 && c0 >= 0 && nr >= 0 && nc >= 0 && r0 + nr <= M->rows && c0 + nc <= M->cols;
}

}  // namespace

extern "C" int iak3d_mat_create(iak3d_ctx* ctx, int64_t rows, int64_t cols, iak3d_mat** out) {
  if (!ctx || !out || rows <= 0 || cols <= 0) return IAK3D_ERR_ARG;
  cudaSetDevice(ctx->device);
  iak3d_mat* M = new iak3d_mat();
  M->ctx = ctx; M->rows = rows; M->cols = cols;
  M->prow = round_up(rows, TM); M->pcol = round_up(cols, TN);
  if (iak_pool_alloc(ctx, (void**)&M->d, M->bytes()) != cudaSuccess) {
    cudaGetLastError(); delete M;
    ctx->err = "mat_create: allocation failed";
    return IAK3D_ERR_CUDA;
  }
  const cudaError_t e = cudaMemsetAsync(M->d, 0, M->bytes(), ctx->stream);
  if (e != cudaSuccess) { iak_pool_free(ctx, M->d); delete M; ctx->err = std::string("mat_create: ") + cudaGetErrorString(e); return IAK3D_ERR_CUDA; }
  *out = M;
  return IAK3D_OK;
}

extern "C" int iak3d_mat_destroy(iak3d_mat* M) {
  if (!M) return IAK3D_OK;
  cudaSetDevice(M->ctx->device);
  iak_pool_free(M->ctx, M->d);                  // stream-ordered: pending kernels on the stream finish first
  delete M;
  return IAK3D_OK;
}

// Hands the cached matrix memory back to the driver (end of a prediction call: the factor buffers of a large fit are
// plain allocations and must be able to use it).
extern "C" int iak3d_mat_trim(iak3d_ctx* ctx) {
  if (!ctx) return IAK3D_ERR_ARG;
  cudaSetDevice(ctx->device);
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  cudaMemPool_t pool;
  CUDA_TRY(ctx, cudaDeviceGetDefaultMemPool(&pool, ctx->device));
  CUDA_TRY(ctx, cudaMemPoolTrimTo(pool, 0));
  return IAK3D_OK;
}

extern "C" int iak3d_mat_shape(const iak3d_mat* M, int64_t* rows, int64_t* cols) {
  if (!M) return IAK3D_ERR_ARG;
  if (rows) *rows = M->rows;
  if (cols) *cols = M->cols;
  return IAK3D_OK;
}

extern "C" int iak3d_mat_set(iak3d_mat* M, int64_t r0, int64_t c0, int64_t nr, int64_t nc, const double* src, int64_t ld) {
  if (!M || !src || ld < nr) return IAK3D_ERR_ARG;
  iak3d_ctx* ctx = M->ctx;
  if (!in_range(M, r0, c0, nr, nc)) { ctx->err = "mat_set: sub-block out of range"; return IAK3D_ERR_ARG; }
  if (nr == 0 || nc == 0) return IAK3D_OK;
  cudaSetDevice(ctx->device);
  double* stage = nullptr;
  if (iak_pool_alloc(ctx, (void**)&stage, (size_t)nr * nc * 8) != cudaSuccess) { cudaGetLastError(); ctx->err = "mat_set: allocation failed"; return IAK3D_ERR_CUDA; }
  cudaError_t e = cudaMemcpy2DAsync(stage, (size_t)nr * 8, src, (size_t)ld * 8, (size_t)nr * 8, (size_t)nc, cudaMemcpyHostToDevice, ctx->stream);
  for (int64_t j0 = 0; j0 < nc && e == cudaSuccess; j0 += 65535) {
    const int64_t ncj = std::min<int64_t>(65535, nc - j0);
    dim3 grid((unsigned)std::min<int64_t>((nr + 255) / 256, 1024), (unsigned)ncj);
    mat_set_kernel<<<grid, 256, 0, ctx->stream>>>(M->d, M->nRU(), r0, c0 + j0, nr, ncj, stage + (size_t)j0 * nr, nr);
    ctx->launches++;
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  iak_pool_free(ctx, stage);
  CUDA_TRY(ctx, e);
  return IAK3D_OK;
}

extern "C" int iak3d_mat_get(const iak3d_mat* M, int64_t r0, int64_t c0, int64_t nr, int64_t nc, double* dst, int64_t ld) {
  if (!M || !dst || ld < nr) return IAK3D_ERR_ARG;
  iak3d_ctx* ctx = M->ctx;
  if (!in_range(M, r0, c0, nr, nc)) { ctx->err = "mat_get: sub-block out of range"; return IAK3D_ERR_ARG; }
  if (nr == 0 || nc == 0) return IAK3D_OK;
  cudaSetDevice(ctx->device);
  double* stage = nullptr;
  if (iak_pool_alloc(ctx, (void**)&stage, (size_t)nr * nc * 8) != cudaSuccess) { cudaGetLastError(); ctx->err = "mat_get: allocation failed"; return IAK3D_ERR_CUDA; }
  cudaError_t e = cudaSuccess;
  for (int64_t j0 = 0; j0 < nc && e == cudaSuccess; j0 += 65535) {
    const int64_t ncj = std::min<int64_t>(65535, nc - j0);
    dim3 grid((unsigned)std::min<int64_t>((nr + 255) / 256, 1024), (unsigned)ncj);
    mat_get_kernel<<<grid, 256, 0, ctx->stream>>>(M->d, M->nRU(), r0, c0 + j0, nr, ncj, stage + (size_t)j0 * nr, nr);
    ctx->launches++;
    e = cudaGetLastError();
  }
  if (e == cudaSuccess)
    e = cudaMemcpy2DAsync(dst, (size_t)ld * 8, stage, (size_t)nr * 8, (size_t)nr * 8, (size_t)nc, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  iak_pool_free(ctx, stage);
  CUDA_TRY(ctx, e);
  return IAK3D_OK;
}

extern "C" int iak3d_mat_copy(iak3d_mat* D, int64_t dr0, int64_t dc0, const iak3d_mat* S, int64_t sr0, int64_t sc0, int64_t nr,
                              int64_t nc, int32_t transpose, double alpha, double beta) {
  if (!D || !S || D->ctx != S->ctx) return IAK3D_ERR_ARG;
  iak3d_ctx* ctx = D->ctx;
  const bool ok = in_range(D, dr0, dc0, nr, nc) && (transpose ? in_range(S, sr0, sc0, nc, nr) : in_range(S, sr0, sc0, nr, nc));
  if (!ok) { ctx->err = "mat_copy: sub-block out of range"; return IAK3D_ERR_ARG; }
  if (D == S) { ctx->err = "mat_copy: source and destination must be different matrices"; return IAK3D_ERR_ARG; }
  if (nr == 0 || nc == 0) return IAK3D_OK;
  cudaSetDevice(ctx->device);
  for (int64_t j0 = 0; j0 < nc; j0 += (int64_t)65535 * 32) {
    const int64_t ncj = std::min<int64_t>((int64_t)65535 * 32, nc - j0);
    dim3 grid((unsigned)((nr + 31) / 32), (unsigned)((ncj + 31) / 32));
    mat_copy_kernel<<<grid, 256, 0, ctx->stream>>>(D->d, D->nRU(), dr0, dc0 + j0, S->d, S->nRU(), transpose ? sr0 + j0 : sr0,
                                                   transpose ? sc0 : sc0 + j0, nr, ncj, transpose, alpha, beta);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
  }
  return IAK3D_OK;
}

// C = [C +] alpha A B'   (A: rows(C) x K, B: cols(C) x K)
extern "C" int iak3d_mat_gemm_nt(iak3d_mat* Cm, double alpha, const iak3d_mat* A, const iak3d_mat* B, int32_t accumulate) {
  if (!Cm || !A || !B || A->ctx != Cm->ctx || B->ctx != Cm->ctx) return IAK3D_ERR_ARG;
  iak3d_ctx* ctx = Cm->ctx;
  if (A->rows != Cm->rows || B->rows != Cm->cols || A->cols != B->cols) { ctx->err = "mat_gemm_nt: shapes do not conform"; return IAK3D_ERR_ARG; }
  if (Cm == A || Cm == B) { ctx->err = "mat_gemm_nt: the result must not alias an operand"; return IAK3D_ERR_ARG; }
  cudaSetDevice(ctx->device);
  int32_t* d_status = nullptr;
  if (iak_pool_alloc(ctx, (void**)&d_status, 2 * sizeof(int32_t)) != cudaSuccess) { cudaGetLastError(); ctx->err = "mat_gemm_nt: allocation failed"; return IAK3D_ERR_CUDA; }
  cudaMemsetAsync(d_status, 0, 2 * sizeof(int32_t), ctx->stream);
  int rc = iak_gemm_nt(ctx, A->d, A->prow, B->d, B->prow, Cm->d, Cm->prow, Cm->pcol, A->pcol, alpha, accumulate ? 1 : 0, d_status);
  int32_t hs[2] = {0, 0};
  if (rc == 0 && cudaMemcpy(hs, d_status, sizeof(hs), cudaMemcpyDeviceToHost) != cudaSuccess) { ctx->err = "mat_gemm_nt: status copy failed"; rc = IAK3D_ERR_CUDA; }
  iak_pool_free(ctx, d_status);
  if (rc == 0 && hs[1] != 0) { ctx->err = "mat_gemm_nt: wait timed out (internal error)"; rc = IAK3D_ERR_TIMEOUT; }
  return rc;
}

// M := inverse of the symmetric positive definite M (chol2inv(chol(M))); *status = IAK3D_NOT_SPD leaves M unchanged.
extern "C" int iak3d_mat_spd_inverse(iak3d_mat* M, int32_t* status) {
  if (!M || !status || M->rows != M->cols) return IAK3D_ERR_ARG;
  iak3d_ctx* ctx = M->ctx;
  cudaSetDevice(ctx->device);
  const int64_t n = M->rows;
  Slot* s = nullptr;
  int rc = iak_alloc_slot(ctx, n, 0, 0, 0, &s, false);
  if (rc != 0) return rc;
  do {
    if (s->ld != M->prow || s->Npad != M->pcol) { ctx->err = "mat_spd_inverse: geometry mismatch (internal error)"; rc = IAK3D_ERR_ARG; break; }
    if (M->pcol > 65535) { ctx->err = "mat_spd_inverse: more than 65535 columns"; rc = IAK3D_ERR_ARG; break; }
    dim3 grid((unsigned)std::min<int64_t>((M->prow + 255) / 256, 1024), (unsigned)M->pcol);
    mat_pack_factor_kernel<<<grid, 256, 0, ctx->stream>>>(M->d, s->d_L, M->nRU(), n, M->prow);
    ctx->launches++;
    double lndet = 0.0;
    rc = iak_factor_slot(ctx, s, status, &lndet, nullptr);
    if (rc != 0 || *status != 0) break;
    double* out = nullptr; int64_t rowsU = 0;
    rc = iak_inverse_blocked(ctx, s, &out, &rowsU, nullptr, M->d);
    if (rc != 0) break;
    if (M->pcol > n) {
      mat_zero_pad_diag_kernel<<<(unsigned)((M->pcol - n + 63) / 64), 64, 0, ctx->stream>>>(M->d, M->nRU(), n, M->pcol);
      ctx->launches++;
    }
    const cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { ctx->err = std::string("mat_spd_inverse: ") + cudaGetErrorString(e); rc = IAK3D_ERR_CUDA; }
  } while (0);
  cudaStreamSynchronize(ctx->stream);
  iak_free_slot(s);
  return rc;
}

// M (n x n) = setCIAK3D(ds): the square covariance matrix of the dataset's supports, built on the device.
extern "C" int iak3d_mat_cov_build(iak3d_ds* ds, const iak3d_pars* pars, iak3d_mat* M) {
  if (!ds || !pars || !M || M->ctx != ds->ctx) return IAK3D_ERR_ARG;
  iak3d_ctx* ctx = ds->ctx;
  if (M->rows != ds->n || M->cols != ds->n) { ctx->err = "mat_cov_build: the matrix must be n x n"; return IAK3D_ERR_ARG; }
  cudaSetDevice(ctx->device);
  EvalPlan pl;
  int st = iak_make_plan(ctx, pars, true, &pl);
  if (st == 0) st = iak_plan_check_avsd(ctx, ds, &pl);
  if (st != 0) return st;
  if (pl.status != 0) return pl.status;
  Slot* s = nullptr;
  st = iak_get_slot(ds, 0, &s);
  if (st != 0) return st;
  CovArgs a;
  st = iak_square_tables(ctx, ds, pl, s, &a);
  if (st != 0) return st;
  st = launch_cov_rect(ctx, a, M->d, M->prow, M->prow, M->pcol, 1, /*blocked=*/1);
  if (st != 0) return st;
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return IAK3D_OK;
}

// out[j] = sum_i A[i,j] B[i,j]   (colSums(A * B), compositeLikIAK3D.R:795)
extern "C" int iak3d_mat_coldot(const iak3d_mat* A, const iak3d_mat* B, double* out) {
  if (!A || !B || !out || A->ctx != B->ctx) return IAK3D_ERR_ARG;
  iak3d_ctx* ctx = A->ctx;
  if (A->rows != B->rows || A->cols != B->cols) { ctx->err = "mat_coldot: shapes differ"; return IAK3D_ERR_ARG; }
  cudaSetDevice(ctx->device);
  double* d_out = nullptr;
  if (iak_pool_alloc(ctx, (void**)&d_out, (size_t)A->cols * 8) != cudaSuccess) { cudaGetLastError(); ctx->err = "mat_coldot: allocation failed"; return IAK3D_ERR_CUDA; }
  mat_coldot_kernel<<<(unsigned)((A->cols * 32 + 255) / 256), 256, 0, ctx->stream>>>(A->d, A->nRU(), B->d, B->nRU(), A->rows, A->cols, d_out);
  ctx->launches++;
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out, (size_t)A->cols * 8, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  iak_pool_free(ctx, d_out);
  CUDA_TRY(ctx, e);
  return IAK3D_OK;
}
