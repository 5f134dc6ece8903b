// tables.cu -- unique-pair tables: the transcendental part of the covariance build.
//
// The reference evaluates the depth kernel on unique interval pairs (iaCovMatern.R:76-89) and the
// horizontal model on unique location pairs (fitIAK3D.R:979) and then gathers them to n x n through
// index vectors (fitIAK3D.R:1038-1074).  We keep that split: these kernels do the exp / Bessel work
// on the (small) unique sets; covbuild.cu expands by id.  This translation unit is compiled with
// -fmad=false so that the closed forms round like the reference's operation order.
#include <algorithm>

#include "common.cuh"

__global__ void iacov_table_kernel(IaPrm P, const double* __restrict__ dI1, int64_t n1,
                                   const double* __restrict__ dI2, int64_t n2, double* __restrict__ out,
                                   double* __restrict__ avvar1) {
  const int64_t total = n1 * n2;
  // one set against itself: the reference evaluates the pairs (i <= j) only, first interval = the one with the smaller index,
  // and fills both triangles with that value (iaCovMatern.R:60-70, 99-100).  Same here: half the exp() work, and the table is
  // symmetric to the bit also where two intervals share their upper limit (the only case in which the argument order matters).
  const bool sym = (dI1 == dI2 && n1 == n2);
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = idx % n1, j = idx / n1;
    if (sym) {
      if (i >= j) {
        const double v = iak_avcov(dI1[j], dI1[n1 + j], dI1[i], dI1[n1 + i], &P);
        out[idx] = v;
        out[j + i * n1] = v;
      }
    } else {
      out[idx] = iak_avcov(dI1[i], dI1[n1 + i], dI2[j], dI2[n2 + j], &P);
    }
    if (avvar1 != nullptr && j == 0) avvar1[i] = iak_avvar(dI1[i], dI1[n1 + i], &P);
  }
}

int launch_iacov_table(iak3d_ctx* ctx, const IaPrm& P, const double* d_dI1, int64_t n1, const double* d_dI2, int64_t n2,
                       double* d_out, double* d_avvar1) {
  const int64_t total = n1 * n2;
  if (total == 0) return 0;
  int blocks = (int)((total + 255) / 256);
  if (blocks > ctx->sm_count * 8) blocks = ctx->sm_count * 8;
  iacov_table_kernel<<<blocks, 256, 0, ctx->stream>>>(P, d_dI1, n1, d_dI2, n2, d_out, d_avvar1);
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  return 0;
}

// setCApproxIAK3D[2]: avCor * avsd(I) * avsd(J) on unique interval pairs (fitIAK3D.R:1960-1963, 2146).  The multiplication
// order is the reference's: rectangular (src * s1[i]) * s2[j]; symmetric (packed upper storage, element (i <= j) times
// s[j] then s[i])  (src * s[max]) * s[min].
__global__ void scale_table_kernel(const double* __restrict__ src, const double* __restrict__ s1, int64_t n1,
                                   const double* __restrict__ s2, int64_t n2, int symmetric, double* __restrict__ out,
                                   double* __restrict__ avvar1) {
  const int64_t total = n1 * n2;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = idx % n1, j = idx / n1;
    double v = src[idx];
    if (symmetric) { const int64_t hi = i > j ? i : j, lo = i > j ? j : i; v = v * s1[hi]; v = v * s1[lo]; }
    else { v = v * s1[i]; v = v * s2[j]; }
    out[idx] = v;
    if (avvar1 != nullptr && j == 0) avvar1[i] = s1[i] * s1[i];
  }
}

int launch_scale_table(iak3d_ctx* ctx, const double* d_src, const double* d_s1, int64_t n1, const double* d_s2, int64_t n2,
                       int symmetric, double* d_out, double* d_avvar1) {
  const int64_t total = n1 * n2;
  if (total == 0) return 0;
  int blocks = (int)((total + 255) / 256);
  if (blocks > ctx->sm_count * 8) blocks = ctx->sm_count * 8;
  scale_table_kernel<<<blocks, 256, 0, ctx->stream>>>(d_src, d_s1, n1, d_s2, n2, symmetric, d_out, d_avvar1);
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  return 0;
}

// phi_x on unique location pairs; D as in xyDist (iaCovMatern.R:888-920): squares added in dimension
// order, then sqrt.  Optionally records the exact-equality indicator 1[D == 0] (fitIAK3D.R:1138).
__global__ void phix_table_kernel(HxPrm H, const double* __restrict__ x1, int64_t n1, const double* __restrict__ x2,
                                  int64_t n2, double* __restrict__ out, uint8_t* __restrict__ zero, int symmetric) {
  __shared__ double s_rt[4 * IAK_BK_NRT];
  const bool bessel = (H.modelx == 2 && H.half == 0);
  if (bessel) {
    for (int i = 1 + threadIdx.x; i < IAK_BK_NRT; i += blockDim.x) iak_besselk_table(H.nu, s_rt, i);
    __syncthreads();
  }
  const double* rt = bessel ? s_rt : nullptr;
  const int64_t total = n1 * n2;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = idx % n1, j = idx / n1;
    // x1 == x2 and an expensive model (Bessel-K): half the entries are computed and mirrored -- a strided store.  The cheap
    // models compute both halves instead and keep every store coalesced: (x_i - x_j)^2 == (x_j - x_i)^2 exactly, so the two
    // halves carry the same bits either way.
    const bool mirror = symmetric && bessel;
    if (mirror && i < j) continue;
    const double dx = x1[i] - x2[j], dy = x1[n1 + i] - x2[n2 + j];
    double D = 0.0;
    D = D + dx * dx;
    D = D + dy * dy;
    D = sqrt(D);
    if (out != nullptr) {
      const double v = iak_phix(D, &H, rt);
      out[idx] = v;
      if (mirror) out[i * n1 + j] = v;
    }
    if (zero != nullptr) zero[idx] = (D == 0.0) ? 1 : 0;
  }
}

// rectangular case without the 64-bit index split: blockIdx.y = location of set 2, threads run down the rows of set 1
__global__ void __launch_bounds__(256) phix_rect_kernel(HxPrm H, const double* __restrict__ x1, int64_t n1, const double* __restrict__ x2,
                                                        int64_t n2, int64_t j0, double* __restrict__ out, uint8_t* __restrict__ zero) {
  __shared__ double s_rt[4 * IAK_BK_NRT];
  const bool bessel = (H.modelx == 2 && H.half == 0);
  if (bessel) {
    for (int i = 1 + threadIdx.x; i < IAK_BK_NRT; i += blockDim.x) iak_besselk_table(H.nu, s_rt, i);
    __syncthreads();
  }
  const double* rt = bessel ? s_rt : nullptr;
  const int64_t j = j0 + blockIdx.y;
  const double xj = x2[j], yj = x2[n2 + j];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n1; i += (int64_t)gridDim.x * blockDim.x) {
    const double dx = x1[i] - xj, dy = x1[n1 + i] - yj;
    double D = 0.0;
    D = D + dx * dx;
    D = D + dy * dy;
    D = sqrt(D);
    if (out != nullptr) out[j * n1 + i] = iak_phix(D, &H, rt);
    if (zero != nullptr) zero[j * n1 + i] = (D == 0.0) ? 1 : 0;
  }
}

int launch_phix_table(iak3d_ctx* ctx, const HxPrm& H, const double* d_x1, int64_t n1, const double* d_x2, int64_t n2,
                      double* d_out, uint8_t* d_zero) {
  const int64_t total = n1 * n2;
  if (total == 0) return 0;
  int blocks = (int)((total + 255) / 256);
  if (blocks > ctx->sm_count * 16) blocks = ctx->sm_count * 16;
  const int symmetric = (d_x1 == d_x2 && n1 == n2 && d_zero == nullptr) ? 1 : 0;
  if (!symmetric && n1 >= 1024) {
    const unsigned gx = (unsigned)std::min<int64_t>((n1 + 255) / 256, 64);
    for (int64_t j0 = 0; j0 < n2; j0 += 65535) {
      const unsigned gy = (unsigned)std::min<int64_t>(65535, n2 - j0);
      phix_rect_kernel<<<dim3(gx, gy), 256, 0, ctx->stream>>>(H, d_x1, n1, d_x2, n2, j0, d_out, d_zero);
      ctx->launches++;
      CUDA_TRY(ctx, cudaGetLastError());
    }
    return 0;
  }
  phix_table_kernel<<<blocks, 256, 0, ctx->stream>>>(H, d_x1, n1, d_x2, n2, d_out, d_zero, symmetric);
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  return 0;
}

// ---- the square tables of MANY problems in one launch each ------------------------------------------------------------------
// A composite-likelihood evaluation builds one depth table and one horizontal table per block pair: 286 pairs = 572 launches of
// 5-10 us kernels, each alone too small for the machine and separated by launch gaps (6.5 ms of a 229 ms evaluation).  The jobs
// are queued on the host (TableQueue, common.cuh) and run as two launches, blockIdx.y = job.  Same arithmetic per entry as the
// one-job kernels above (same device functions, same -fmad=false translation unit): same bits.
__global__ void iacov_table_batch_kernel(const IaTabJob* __restrict__ jobs) {
  const IaTabJob& jb = jobs[blockIdx.y];
  const int64_t n1 = jb.n, total = n1 * n1;
  const double* __restrict__ dI = jb.dI;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = idx % n1, j = idx / n1;
    if (i >= j) {                                  // pairs (j <= i), first interval = the smaller index; mirrored (see iacov_table_kernel)
      const double v = iak_avcov(dI[j], dI[n1 + j], dI[i], dI[n1 + i], &jb.P);
      jb.out[idx] = v;
      jb.out[j + i * n1] = v;
    }
    if (jb.av != nullptr && j == 0) jb.av[i] = iak_avvar(dI[i], dI[n1 + i], &jb.P);
  }
}

__global__ void phix_table_batch_kernel(const HxTabJob* __restrict__ jobs) {
  __shared__ double s_rt[4 * IAK_BK_NRT];
  const HxTabJob& jb = jobs[blockIdx.y];
  const HxPrm H = jb.H;
  const bool bessel = (H.modelx == 2 && H.half == 0);
  if (bessel) {
    for (int i = 1 + threadIdx.x; i < IAK_BK_NRT; i += blockDim.x) iak_besselk_table(H.nu, s_rt, i);
    __syncthreads();
  }
  const double* rt = bessel ? s_rt : nullptr;
  const int64_t n1 = jb.n, total = n1 * n1;
  const double* __restrict__ x = jb.x;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = idx % n1, j = idx / n1;
    if (bessel && i < j) continue;                // expensive model: mirrored half (see phix_table_kernel)
    const double dx = x[i] - x[j], dy = x[n1 + i] - x[n1 + j];
    double D = 0.0;
    D = D + dx * dx;
    D = D + dy * dy;
    D = sqrt(D);
    const double v = iak_phix(D, &H, rt);
    jb.out[idx] = v;
    if (bessel) jb.out[i * n1 + j] = v;
  }
}

int flush_table_queue(iak3d_ctx* ctx, TableQueue* tq) {
  if (!tq) return 0;
  const size_t na = tq->ia.size(), nh = tq->hx.size();
  if (na + nh == 0) return 0;
  const size_t bytes = na * sizeof(IaTabJob) + nh * sizeof(HxTabJob);
  if (bytes > ctx->tabjobs_cap) {
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->d_tabjobs); ctx->d_tabjobs = nullptr;
    if (ctx->h_tabjobs) { cudaFreeHost(ctx->h_tabjobs); ctx->h_tabjobs = nullptr; }
    CUDA_TRY(ctx, cudaMalloc(&ctx->d_tabjobs, bytes * 2));
    CUDA_TRY(ctx, cudaMallocHost(&ctx->h_tabjobs, bytes * 2));
    ctx->tabjobs_cap = bytes * 2;
  } else {
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));           // the pinned staging buffer of the previous flush
  }
  char* h = static_cast<char*>(ctx->h_tabjobs);
  if (na) memcpy(h, tq->ia.data(), na * sizeof(IaTabJob));
  if (nh) memcpy(h + na * sizeof(IaTabJob), tq->hx.data(), nh * sizeof(HxTabJob));
  CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_tabjobs, h, bytes, cudaMemcpyHostToDevice, ctx->stream));
  const char* d = static_cast<const char*>(ctx->d_tabjobs);
  for (size_t j0 = 0; j0 < na; j0 += 65535) {
    const size_t jc = std::min<size_t>(65535, na - j0);
    int64_t mx = 0;
    for (size_t j = j0; j < j0 + jc; j++) mx = std::max(mx, tq->ia[j].n * tq->ia[j].n);
    const unsigned bx = (unsigned)std::max<int64_t>(1, std::min<int64_t>((mx + 255) / 256, 64));
    iacov_table_batch_kernel<<<dim3(bx, (unsigned)jc), 256, 0, ctx->stream>>>(reinterpret_cast<const IaTabJob*>(d) + j0);
    ctx->launches++;
  }
  for (size_t j0 = 0; j0 < nh; j0 += 65535) {
    const size_t jc = std::min<size_t>(65535, nh - j0);
    int64_t mx = 0;
    for (size_t j = j0; j < j0 + jc; j++) mx = std::max(mx, tq->hx[j].n * tq->hx[j].n);
    const unsigned bx = (unsigned)std::max<int64_t>(1, std::min<int64_t>((mx + 255) / 256, 256));
    phix_table_batch_kernel<<<dim3(bx, (unsigned)jc), 256, 0, ctx->stream>>>(reinterpret_cast<const HxTabJob*>(d + na * sizeof(IaTabJob)) + j0);
    ctx->launches++;
  }
  CUDA_TRY(ctx, cudaGetLastError());
  tq->ia.clear(); tq->hx.clear();
  return 0;
}

__global__ void phix_vec_kernel(HxPrm H, const double* __restrict__ D, int64_t m, double* __restrict__ out) {
  __shared__ double s_rt[4 * IAK_BK_NRT];
  const bool bessel = (H.modelx == 2 && H.half == 0);
  if (bessel) {
    for (int i = 1 + threadIdx.x; i < IAK_BK_NRT; i += blockDim.x) iak_besselk_table(H.nu, s_rt, i);
    __syncthreads();
  }
  const double* rt = bessel ? s_rt : nullptr;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < m; idx += (int64_t)gridDim.x * blockDim.x)
    out[idx] = iak_phix(D[idx], &H, rt);
}

int launch_phix_vec(iak3d_ctx* ctx, const HxPrm& H, const double* d_D, int64_t m, double* d_out) {
  if (m == 0) return 0;
  int blocks = (int)((m + 255) / 256);
  if (blocks > ctx->sm_count * 16) blocks = ctx->sm_count * 16;
  phix_vec_kernel<<<blocks, 256, 0, ctx->stream>>>(H, d_D, m, d_out);
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  return 0;
}

// ---- host-side parameter preparation ------------------------------------------------------------
static double pochdown(double q, int k) {     // fields.pochdown incl. the 1:(k-1) quirk (iaCovMatern.R:1435-1452)
  if (k == 0) return 1.0;
  double n = q;
  if (k - 1 >= 1) for (int j = 1; j <= k - 1; j++) n = n * (q - j);
  return n;
}
static double pochup(double q, int k) {       // iaCovMatern.R:1454-1471
  if (k == 0) return 1.0;
  double n = q;
  if (k - 1 >= 1) for (int j = 1; j <= k - 1; j++) n = n * (q + j);
  return n;
}
// Wendland.beta(n = 2, k): column k of the (k+1)x(k+1) table (iaCovMatern.R:1405-1433)
static void wendland_beta_col(int k, double* out /* k+1 */) {
  const int l = 1 + k + 1;
  std::vector<double> M((size_t)(k + 1) * (k + 1), 0.0);
  auto at = [&](int r, int c) -> double& { return M[(size_t)c * (k + 1) + r]; };
  at(0, 0) = 1.0;
  for (int col = 0; col <= k - 1; col++) {
    int row = 0; double beta = 0.0;
    for (int m = 0; m <= col; m++) beta += at(m, col) * pochdown(m + 1, m - row + 1) / pochup(l + 2 * col - m + 1, m - row + 2);
    at(row, col + 1) = beta;
    for (row = 1; row <= col + 1; row++) {
      beta = 0.0;
      for (int m = row - 1; m <= col; m++) beta += at(m, col) * pochdown(m + 1, m - row + 1) / pochup(l + 2 * col - m + 1, m - row + 2);
      at(row, col + 1) = beta;
    }
  }
  for (int r = 0; r <= k; r++) out[r] = at(r, k);
}

int hx_init(HxPrm* H, int modelx, double a, double nu) {
  *H = HxPrm();
  H->modelx = modelx; H->a = a; H->nu = nu;
  if (modelx == IAK3D_MODELX_SPHERICAL) {
    if (!(a > 0.0)) return IAK3D_BAD_PARS;                               // iaCovMatern.R:335
  } else if (modelx == IAK3D_MODELX_MATERN) {
    if (!(a > 0.0 && nu >= 0.05 && nu <= 20.0)) return IAK3D_BAD_PARS;   // :354
    H->s = sqrt(2.0 * nu) / a;
    H->lconst = -(nu - 1.0) * log(2.0) - lgamma(nu);
    H->half = (nu == 0.5) ? 1 : (nu == 1.5) ? 3 : (nu == 2.5) ? 5 : 0;
    if (H->half == 0) iak_bk_fit_init(&H->bk, nu);                        // order-dependent work of K_nu, once per table
  } else if (modelx == IAK3D_MODELX_WENDLAND) {
    if (!(a > 0.0 && nu >= 1.0 && nu <= 20.0)) return IAK3D_BAD_PARS;    // :379
    H->kf = (int)floor(nu); H->kc = (int)ceil(nu);
    H->wf = (double)H->kc - nu; H->wc = nu - (double)H->kf;             // :385-386
    wendland_beta_col(H->kf, H->betaf); wendland_beta_col(H->kc, H->betac);
    H->scf = iak_wend_eval(0.0, H->kf, H->betaf); H->scc = iak_wend_eval(0.0, H->kc, H->betac);
  }
  return 0;
}
