// covbuild.cu -- covariance assembly (setCIAK3D fitIAK3D.R:955-1112, setCIAK3D2 :1117-1252).
//
//   A(i,j) = cme*[i==j] + [x_i==x_j]*(cx0 + cxd0*Phi_cxd0(I_i,I_j)) + kd1*Phi_cd1(I_i,I_j)
//            + phi_x(x_i,x_j)*(cx1 + cxd1*Phi_cxd1(I_i,I_j))
//
// The reference reaches this through two n(n+1)/2-long gather index vectors (iaCovMatern.R:473-481);
// here each entry is formed directly from the row's (location id, interval id) pair and the small
// unique-pair tables of tables.cu.  The factor build (cov_factor_kernel) is HBM-write-bound: 8 bytes per
// entry, 16-byte evict-first stores down the columns of the blocked layout, depth tables expanded along the
// rows so that their loads are coalesced.  The rectangular builds (cov_rect_kernel: matrices returned to the
// caller, kriging panels) keep the reference's term-by-term sums (cov_value).
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"

bool cov_one_table(const CovArgs& h) {
  for (int c = 0; c < 3; c++) if (h.tidx[c] > 0) return false;
  return h.tidx[1] == 0;
}

// L2 residency: the matrix streams through L2 once (evict-first stores) so that it does not push out the small tables
// every entry is made from (evict-last loads).  ncu on the version without hints: L2 hit rate 31 %, the kernel stalled on
// table loads that went to HBM behind 1.6 GB of its own writes.  The loads are NOT volatile: the compiler must be free
// to issue the loads of several columns before the first store (the kernel is bound by memory latency).
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t p; asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p;
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p; asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p;
}
__device__ __forceinline__ double ld_keep(const double* p, uint64_t pol) {
  double v; asm("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol)); return v;
}
__device__ __forceinline__ double2 ld_keep2(const double2* p, uint64_t pol) {
  double2 v; asm("ld.global.nc.L2::cache_hint.v2.f64 {%0,%1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(pol)); return v;
}
__device__ __forceinline__ void st_stream2(double* p, double a, double b, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1,%2}, %3;" ::"l"(p), "d"(a), "d"(b), "l"(pol) : "memory");
}

struct CovDev {
  const double* T[3];
  int tidx[3];
  int ntab;
  const double* comb;     // per-evaluation S(di,dj) = cx0 + cxd0*Phi_cxd0 table (nd^2), generic interior path
  const double* E;        // depth tables expanded along the rows, see expand_tables_kernel: E[dj*ldE + i] (double or double2)
  int64_t ldE;
  int one_table;          // every present component reads the same depth table
  const double* phix;
  const uint8_t* same;
  double cx0, cx1, kd1, cxd0, cxd1, cme;
  const int32_t *xid_r, *did_r, *xid_c, *did_c;
  int64_t nr, nc, nxU_r, ndIU_r;
};

// Terms are accumulated in the reference's order (fitIAK3D.R:973,976,1038,1048/1052,1067,1071) and this file is
// compiled with -fmad=false, so an entry rounds like the chain of R matrix additions.
__device__ __forceinline__ double cov_value(const CovDev& a, int xi, int di, int xj, int dj, bool diag) {
  double t[3];
#pragma unroll
  for (int k = 0; k < 3; k++) t[k] = (k < a.ntab) ? __ldg(a.T[k] + (int64_t)dj * a.ndIU_r + di) : 0.0;
  auto pick = [&](int k) -> double { return k == 0 ? t[0] : (k == 1 ? t[1] : (k == 2 ? t[2] : 0.0)); };
  const double td = pick(a.tidx[0]), t0 = pick(a.tidx[1]), t1 = pick(a.tidx[2]);
  bool same;
  if (a.same != nullptr) same = __ldg(a.same + (int64_t)xj * a.nxU_r + xi) != 0;
  else same = (xi == xj);
  double v = diag ? a.cme : 0.0;
  if (same) { v = v + a.cx0; v = v + a.cxd0 * t0; }
  if (a.tidx[0] >= 0) v = v + a.kd1 * td;
  if (a.phix != nullptr) {
    const double px = __ldg(a.phix + (int64_t)xj * a.nxU_r + xi);
    v = v + a.cx1 * px;
    v = v + (a.cxd1 * px) * t1;
  }
  return v;
}

// ---- factor-layout build: lower-triangular TM x TN tiles of the padded, BLOCKED factor buffer (common.cuh) ----
// rows [0,n) data, [n,Npad) identity padding, [Npad,Npad+q) the W' rows (bordered system), rest zero.
// One job per evaluation; a batch (the npar+1 evaluations of a forward-difference gradient, the block pairs of a
// composite likelihood) is ONE launch with blockIdx.z = job, so that the tails of the small per-evaluation grids overlap.
struct FbJob {
  CovDev a;
  const double* W; double* L; const double* lncol_v;
  int64_t n, Npad, ld;
  int q, nCB, lncol;
  double* Ebuf; double* Stab; int expand;        // expand_tables_kernel: this job's own expansion (0: shares another job's)
  int fold;                                      // grid layout, see cov_factor_kernel
};

__device__ __forceinline__ void cov_factor_body(const FbJob& J) {
  // tile (r, j): column block j has row blocks j/2 .. nRB-1.  The triangle is folded: grid row y serves column block y
  // and, with the CTAs that column does not need, its mirror nCB-1-y, so that (almost) no CTA is launched for nothing.
  const int64_t ld = J.ld, n = J.n, Npad = J.Npad;
  const int nRB = (int)(ld / TM);
  int j = blockIdx.y, r;
  {
    const int nCBj = J.nCB, half = (nCBj + 1) >> 1;
    if (!J.fold) {
      if (j >= nCBj) return;
      r = (j >> 1) + blockIdx.x;
      if (r >= nRB) return;
    } else if (j >= half) {
      return;
    } else if ((int)blockIdx.x < nRB - (j >> 1)) {
      r = (j >> 1) + blockIdx.x;
    } else {
      const int jm = nCBj - 1 - j;
      if (jm == j) return;                                             // the middle column of an odd count has no mirror
      j = jm;
      r = (j >> 1) + ((int)blockIdx.x - (nRB - ((nCBj - 1 - jm) >> 1)));
      if (r >= nRB) return;
    }
  }
  const CovDev& a = J.a;
  const double* __restrict__ W = J.W; double* __restrict__ L = J.L; const double* __restrict__ lncol_v = J.lncol_v;
  const int q = J.q, lncol = J.lncol;
  const int64_t R0 = (int64_t)r * TM, C0 = (int64_t)j * TN;
  const int tid = threadIdx.x;
  const int rp = tid & 63, cg = tid >> 6;           // row pair (2 rows), column group (16 columns)
  const int lr = rp * 2;
  const int64_t i0 = R0 + lr;
  // this thread's 16 columns lie in one unit: column half cg/2, columns (cg&1)*16.. of it; rows lr, lr+1 in row unit lr/64
  double* const obase = L + (((int64_t)(2 * j + (cg >> 1))) * (nRB * 2) + 2 * r + (lr >> 6)) * UNIT + ((cg & 1) * 16) * UL + (lr & 63);

  if (R0 >= C0 + TN && R0 + TM <= n) {
    // ---- interior tile (the bulk of the matrix): strictly below the diagonal, all rows and columns are data.
    // No shared memory, no barrier: every warp is independent.  A thread reads the ids of its 2 rows (one 8-byte load
    // each for location and interval) and of its 16 columns (warp-uniform 16-byte loads), then per 8 columns issues
    // all loads -- the row-expanded depth table E[dj][i] (expand_tables_kernel; coalesced 16 bytes per thread) and
    // phi_x(xi, xj) (1-2 cache lines per warp) -- before any arithmetic or store: the kernel is bound by memory latency.
    //     v = Phi * (kd1 + cxd1 phi_x) + cx1 phi_x   [+ cx0 + cxd0 Phi when the two rows share a profile: rare, warp-uniform test]
    // Stores are evict-first and table loads evict-last, so that the matrix streaming through L2 does not push out the
    // tables it is made from (ncu before the hints: L2 hit rate 31 %).  Rounding differs from the reference's term-by-term
    // sums by an ulp; the tolerance on covariance entries is 1e-10.
    // History (profiles/): v1 ~85 instructions per entry (selects, 64-bit index arithmetic, six dependent FP64 ops): issue-
    // bound; v2 arithmetic folded into tables: bound by the L1 wavefronts of the depth-table GATHER; v3 tables expanded along
    // the rows: L2 thrash + per-CTA id staging through TMA and a barrier (a third of all instructions was prologue).
    // (A warp-specialised persistent variant with TMA-staged table segments and TMA bulk stores of whole units was
    // measured 3x slower: 32 one-KB bulk copies per item cost ~2000 cycles of issue and one CTA per SM exposed latency.)
    const double* __restrict__ phix = a.phix;
    const unsigned nd = (unsigned)a.ndIU_r, nx = (unsigned)a.nxU_r;
    const bool px_on = (phix != nullptr);
    const int64_t ldE = a.ldE;
    const uint64_t keep = policy_evict_last(), strm = policy_evict_first();
    const int2 xx = __ldg(reinterpret_cast<const int2*>(a.xid_r + i0)), dd = __ldg(reinterpret_cast<const int2*>(a.did_r + i0));
    const int xi0 = xx.x, xi1 = xx.y, di0 = dd.x, di1 = dd.y;
    const int4* __restrict__ xc4 = reinterpret_cast<const int4*>(a.xid_c + C0 + cg * 16);
    const int4* __restrict__ dc4 = reinterpret_cast<const int4*>(a.did_c + C0 + cg * 16);
    const bool same_row_profile = (xi1 == xi0);
    if (a.one_table) {
      const double* __restrict__ Er = a.E + i0;
      const double cx0 = a.cx0, cx1 = px_on ? a.cx1 : 0.0, kd1 = a.tidx[0] >= 0 ? a.kd1 : 0.0, cxd0 = a.cxd0, cxd1 = px_on ? a.cxd1 : 0.0;
#pragma unroll
      for (int h = 0; h < 2; h++) {
        int xjs[8], djs[8];
        { const int4 u0 = __ldg(xc4 + 2 * h), u1 = __ldg(xc4 + 2 * h + 1), w0 = __ldg(dc4 + 2 * h), w1 = __ldg(dc4 + 2 * h + 1);
          xjs[0] = u0.x; xjs[1] = u0.y; xjs[2] = u0.z; xjs[3] = u0.w; xjs[4] = u1.x; xjs[5] = u1.y; xjs[6] = u1.z; xjs[7] = u1.w;
          djs[0] = w0.x; djs[1] = w0.y; djs[2] = w0.z; djs[3] = w0.w; djs[4] = w1.x; djs[5] = w1.y; djs[6] = w1.z; djs[7] = w1.w; }
        double2 t[8]; double pa[8], pb[8];
#pragma unroll
        for (int cc = 0; cc < 8; cc++) {
          t[cc] = ld_keep2(reinterpret_cast<const double2*>(Er + (int64_t)djs[cc] * ldE), keep);
          pa[cc] = 0.0; pb[cc] = 0.0;
          if (px_on) {      // (re-using phi_x across the consecutive columns of one profile was measured 20 % slower: it chains the loads)
            const unsigned pbase = (unsigned)xjs[cc] * nx;
            pa[cc] = ld_keep(phix + pbase + xi0, keep);
            if (!same_row_profile) pb[cc] = ld_keep(phix + pbase + xi1, keep);
          }
        }
#pragma unroll
        for (int cc = 0; cc < 8; cc++) {
          const double pbv = same_row_profile ? pa[cc] : pb[cc];
          double va = fma(t[cc].x, fma(cxd1, pa[cc], kd1), cx1 * pa[cc]);
          double vb = fma(t[cc].y, fma(cxd1, pbv, kd1), cx1 * pbv);
          if (__any_sync(0xffffffffu, (xi0 == xjs[cc]) | (xi1 == xjs[cc]))) {     // rare: rows of the column's own profile
            if (xi0 == xjs[cc]) va += fma(cxd0, t[cc].x, cx0);
            if (xi1 == xjs[cc]) vb += fma(cxd0, t[cc].y, cx0);
          }
          st_stream2(obase + (h * 8 + cc) * UL, va, vb, strm);
        }
      }
    } else {
      const double2* __restrict__ Er = reinterpret_cast<const double2*>(a.E) + i0;
      const double* __restrict__ S = a.comb;
#pragma unroll
      for (int h = 0; h < 4; h++) {
        int xjs[4], djs[4];
        { const int4 u0 = __ldg(xc4 + h), w0 = __ldg(dc4 + h);
          xjs[0] = u0.x; xjs[1] = u0.y; xjs[2] = u0.z; xjs[3] = u0.w; djs[0] = w0.x; djs[1] = w0.y; djs[2] = w0.z; djs[3] = w0.w; }
        double2 e0[4], e1[4]; double pa[4], pb[4];
#pragma unroll
        for (int cc = 0; cc < 4; cc++) {
          const double2* e = Er + (int64_t)djs[cc] * ldE;
          e0[cc] = ld_keep2(e, keep); e1[cc] = ld_keep2(e + 1, keep);
          pa[cc] = 0.0; pb[cc] = 0.0;
          if (px_on) {      // (re-using phi_x across the consecutive columns of one profile was measured 20 % slower: it chains the loads)
            const unsigned pbase = (unsigned)xjs[cc] * nx;
            pa[cc] = ld_keep(phix + pbase + xi0, keep);
            if (!same_row_profile) pb[cc] = ld_keep(phix + pbase + xi1, keep);
          }
        }
#pragma unroll
        for (int cc = 0; cc < 4; cc++) {
          const double pbv = same_row_profile ? pa[cc] : pb[cc];
          double va = fma(pa[cc], e0[cc].y, e0[cc].x), vb = fma(pbv, e1[cc].y, e1[cc].x);
          if (__any_sync(0xffffffffu, (xi0 == xjs[cc]) | (xi1 == xjs[cc]))) {
            if (xi0 == xjs[cc]) va += __ldg(S + (unsigned)djs[cc] * nd + di0);
            if (xi1 == xjs[cc]) vb += __ldg(S + (unsigned)djs[cc] * nd + di1);
          }
          st_stream2(obase + (h * 4 + cc) * UL, va, vb, strm);
        }
      }
    }
    return;
  }

  // ---- diagonal and border tiles: generic path with all the case distinctions ----
#pragma unroll 4
  for (int cc = 0; cc < 16; cc++) {
    const int lc = cg * 16 + cc;
    const int64_t c = C0 + lc;
    const int xj = c < n ? a.xid_c[c] : -1, dj = c < n ? a.did_c[c] : 0;
    double v[2];
#pragma unroll
    for (int e = 0; e < 2; e++) {
      const int64_t i = i0 + e;
      double val;
      if (i < c) val = 0.0;                                             // strictly upper: never read
      else if (c >= n) val = (i == c && c < Npad) ? 1.0 : 0.0;          // identity padding columns
      else if (i < n) val = cov_value(a, a.xid_r[i], a.did_r[i], xj, dj, i == c);
      else if (i >= Npad && i < Npad + q) val = (i - Npad == lncol) ? lncol_v[c] : W[(int64_t)(i - Npad) * n + c];   // W' rows
      else val = 0.0;
      v[e] = val;
    }
    *reinterpret_cast<double2*>(obase + cc * UL) = make_double2(v[0], v[1]);   // blocked layout, 16-byte store
  }
}

// two entry points for one body: a batch reads its job from an array in global memory (blockIdx.z = job); a single
// evaluation (the large-n case) gets the job as a kernel parameter -- constant bank, measured 12 % faster at n = 20000.
// Both are held to 80 registers (3 CTAs per SM): more resident warps beat more loads in flight per thread.
__global__ void __launch_bounds__(256) cov_factor_kernel(const FbJob* __restrict__ jobs) { cov_factor_body(jobs[blockIdx.z]); }
__global__ void __launch_bounds__(256, 3) cov_factor_kernel_one(const __grid_constant__ FbJob job) { cov_factor_body(job); }

// Depth tables expanded along the rows for the interior tiles (coalesced instead of gathered loads):
//   one table : E[dj*ldE + i]            = Phi[dj*nd + did[i]]
//   generic   : E2[dj*ldE + i] (double2) = (kd1*Phi_cd1, cx1 + cxd1*Phi_cxd1)[dj*nd + did[i]],  S[dj*nd + di] = cx0 + cxd0*Phi_cxd0
__global__ void __launch_bounds__(256) expand_tables_kernel(const FbJob* __restrict__ jobs) {
  // blockIdx.y = dj (one table column, 8 nd bytes: L1-resident), blockIdx.x walks the rows: coalesced id reads and writes
  const FbJob& J = jobs[blockIdx.z];
  if (!J.expand) return;
  const CovDev& a = J.a;
  const int64_t n = J.n, ldE = a.ldE;
  double* __restrict__ E = J.Ebuf; double* __restrict__ Stab = J.Stab;
  const int64_t nd = a.ndIU_r;
  const int64_t dj = blockIdx.y;
  if (dj >= nd) return;
  const double* Td = a.tidx[0] >= 0 ? a.T[a.tidx[0]] + dj * nd : nullptr;
  const double* T0 = a.tidx[1] >= 0 ? a.T[a.tidx[1]] + dj * nd : nullptr;
  const double* T1 = (a.tidx[2] >= 0 && a.phix != nullptr) ? a.T[a.tidx[2]] + dj * nd : nullptr;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < ldE; i += (int64_t)gridDim.x * blockDim.x) {
    const int di = i < n ? a.did_r[i] : 0;
    if (a.one_table) {
      E[dj * ldE + i] = a.T[0][dj * nd + di];
    } else {
      const double A = Td ? a.kd1 * Td[di] : 0.0;
      const double B = a.phix != nullptr ? (T1 ? fma(a.cxd1, T1[di], a.cx1) : a.cx1) : 0.0;
      reinterpret_cast<double2*>(E)[dj * ldE + i] = make_double2(A, B);
    }
  }
  if (!a.one_table)
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nd; i += (int64_t)gridDim.x * blockDim.x)
      Stab[dj * nd + i] = T0 ? fma(a.cxd0, T0[i], a.cx0) : a.cx0;
}

static void covdev_from(const CovArgs& h, CovDev* a) {
  for (int k = 0; k < 3; k++) { a->T[k] = h.T[k]; a->tidx[k] = h.tidx[k]; }
  a->ntab = h.ntab; a->phix = h.phix; a->same = h.same;
  a->cx0 = h.cx0; a->cx1 = h.cx1; a->kd1 = h.kd1; a->cxd0 = h.cxd0; a->cxd1 = h.cxd1; a->cme = h.cme;
  a->xid_r = h.xid_r; a->did_r = h.did_r; a->xid_c = h.xid_c; a->did_c = h.did_c;
  a->nr = h.nr; a->nc = h.nc; a->nxU_r = h.nxU_r; a->ndIU_r = h.ndIU_r;
  a->comb = h.comb; a->E = nullptr; a->ldE = 0; a->one_table = 0;
}

// ---- site-specific random effects: post-pass over a finished build (CovArgs.nssre > 0) ----------------------------------
struct SsreDev {
  int nssre; double c[IAK3D_MAX_SSRE];
  const int32_t *sid_r, *sid_c; const double *Xr, *Xc; int64_t ldr, ldc;
};
static SsreDev ssre_from(const CovArgs& h) {
  SsreDev s;
  s.nssre = h.nssre;
  for (int k = 0; k < IAK3D_MAX_SSRE; k++) s.c[k] = h.cssre[k];
  s.sid_r = h.sid_r; s.sid_c = h.sid_c; s.Xr = h.Xs_r; s.Xc = h.Xs_c; s.ldr = h.ldXs_r; s.ldc = h.ldXs_c;
  return s;
}
// out(i, j) += [sid_r(i) == sid_c(j)] sum_k c_k Xr(i,k) Xc(j,k) for i < nr, j < nc (lower != 0: i >= j only); blocked (uidx with nRU)
// or column-major with leading dimension ld.  The terms are added one by one like the reference's loop over ip (:1096-1099).
__global__ void __launch_bounds__(256) ssre_add_kernel(SsreDev s, double* __restrict__ out, int64_t nr, int64_t nc, int lower, int blocked,
                                                       int64_t nRU, int64_t ld) {
  const int64_t j = blockIdx.y;
  if (j >= nc) return;
  const int32_t sj = s.sid_c[j];
  double xc[IAK3D_MAX_SSRE];
  for (int k = 0; k < s.nssre; k++) xc[k] = s.Xc[(int64_t)k * s.ldc + j];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nr; i += (int64_t)gridDim.x * blockDim.x) {
    if (lower && i < j) continue;
    if (s.sid_r[i] != sj) continue;
    double* o = blocked ? out + uidx(i, j, nRU) : out + j * ld + i;
    double v = *o;
    for (int k = 0; k < s.nssre; k++) v = v + (s.Xr[(int64_t)k * s.ldr + i] * xc[k]) * s.c[k];
    *o = v;
  }
}
static int launch_ssre_add(iak3d_ctx* ctx, const CovArgs& h, double* d_out, int64_t nr, int64_t nc, int lower, int blocked, int64_t nRU, int64_t ld) {
  if (h.nssre <= 0 || nr <= 0 || nc <= 0) return 0;
  if (!h.sid_r || !h.sid_c || !h.Xs_r || !h.Xs_c) { ctx->err = "site-specific random effects: site ids / columns missing for one of the two sets"; return IAK3D_ERR_ARG; }
  if (nc > 65535) { ctx->err = "ssre: more than 65535 columns"; return IAK3D_ERR_ARG; }
  unsigned bx = (unsigned)((nr + 255) / 256);
  if (bx > 64) bx = 64;
  ssre_add_kernel<<<dim3(bx, (unsigned)nc), 256, 0, ctx->stream>>>(ssre_from(h), d_out, nr, nc, lower, blocked, nRU, ld);
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  return 0;
}
__global__ void ssre_diag_kernel(int nssre, SsreDev s, const double* __restrict__ Xs, int64_t ldXs, int64_t m, double* __restrict__ diag) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  double v = diag[i];
  for (int k = 0; k < nssre; k++) { const double x = Xs[(int64_t)k * ldXs + i]; v = v + (x * x) * s.c[k]; }
  diag[i] = v;
}
int launch_ssre_diag(iak3d_ctx* ctx, int nssre, const double* cssre, const double* d_Xs, int64_t ldXs, int64_t m, double* d_diag) {
  if (nssre <= 0 || m <= 0) return 0;
  SsreDev s; s.nssre = nssre;
  for (int k = 0; k < IAK3D_MAX_SSRE; k++) s.c[k] = k < nssre ? cssre[k] : 0.0;
  s.sid_r = s.sid_c = nullptr; s.Xr = s.Xc = nullptr; s.ldr = s.ldc = 0;
  ssre_diag_kernel<<<(unsigned)((m + 255) / 256), 256, 0, ctx->stream>>>(nssre, s, d_Xs, ldXs, m, d_diag);
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  return 0;
}

int launch_cov_factor_jobs(iak3d_ctx* ctx, const FactorBuildJob* jobs, int count) {
  if (count <= 0) return 0;
  std::vector<FbJob> hj((size_t)count);
  int64_t maxRB = 0, maxCB = 0, maxND = 0, maxLdE = 0;
  int nexpand = 0;
  const int fold = 1;
  for (int i = 0; i < count; i++) {
    const FactorBuildJob& b = jobs[i];
    FbJob& J = hj[(size_t)i];
    covdev_from(b.a, &J.a);
    J.a.one_table = cov_one_table(b.a) ? 1 : 0;      // every present component points at table 0 (cxd0 is always present)
    J.a.ldE = round_up(b.n, 16);        // every table row starts on a 128-byte line (ncu: half-empty load wavefronts otherwise)
    J.W = b.d_W; J.L = b.d_L; J.lncol_v = b.d_lncol; J.n = b.n; J.Npad = b.Npad; J.ld = b.ld; J.q = b.q; J.nCB = b.nCB;
    J.lncol = b.d_lncol ? b.lncol : -1;
    J.Ebuf = b.a.Ebuf; J.Stab = b.a.comb;
    if (b.a.E_shared != nullptr && J.a.one_table) {
      J.a.E = b.a.E_shared; J.expand = 0;            // another evaluation of this batch expands the same table
    } else {
      if (b.a.Ebuf == nullptr || b.a.comb == nullptr) { ctx->err = "cov_factor: expanded-table workspace missing"; return IAK3D_ERR_ARG; }
      J.a.E = b.a.Ebuf; J.expand = 1; nexpand++;
      maxND = std::max<int64_t>(maxND, b.a.ndIU_r); maxLdE = std::max<int64_t>(maxLdE, J.a.ldE);
    }
    {
      const int64_t nRBj = b.ld / TM, nCBj = b.nCB;
      J.fold = fold;
      if (fold) {                                                    // folded triangle: rows(y) + rows(nCB-1-y)
        int64_t gx = 0;
        for (int64_t y = 0; y < (nCBj + 1) / 2; y++) {
          const int64_t jm = nCBj - 1 - y;
          gx = std::max<int64_t>(gx, (nRBj - y / 2) + (jm != y ? nRBj - jm / 2 : 0));
        }
        maxRB = std::max<int64_t>(maxRB, gx); maxCB = std::max<int64_t>(maxCB, (nCBj + 1) / 2);
      } else {
        maxRB = std::max<int64_t>(maxRB, nRBj); maxCB = std::max<int64_t>(maxCB, nCBj);
      }
    }
  }
  if (maxCB > 65535 || maxND > 65535) { ctx->err = "cov_factor: more than 65535 column blocks / unique depth intervals"; return IAK3D_ERR_ARG; }
  // the interior path indexes the nxU x nxU horizontal table (and the nd x nd S table) with 32-bit offsets
  for (int i = 0; i < count; i++)
    if (jobs[i].a.nxU_r > 65535 || jobs[i].a.ndIU_r > 65535) {
      ctx->err = "cov_factor: more than 65535 unique locations or depth intervals in one block (32-bit table offsets)";
      return IAK3D_ERR_ARG;
    }
  if ((size_t)count > ctx->fbjobs_cap) {
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->d_fbjobs); ctx->d_fbjobs = nullptr;
    CUDA_TRY(ctx, cudaMalloc(&ctx->d_fbjobs, (size_t)count * sizeof(FbJob)));
    ctx->fbjobs_cap = (size_t)count;
  }
  const FbJob* dj = static_cast<const FbJob*>(ctx->d_fbjobs);
  for (int z0 = 0; z0 < count; z0 += 65535) {        // gridDim.z limit
    const int zc = std::min(65535, count - z0);
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_fbjobs, hj.data() + z0, (size_t)zc * sizeof(FbJob), cudaMemcpyHostToDevice, ctx->stream));
    if (nexpand > 0) {
      unsigned bx = (unsigned)((maxLdE + 255) / 256);
      if (bx > 64) bx = 64;
      expand_tables_kernel<<<dim3(bx, (unsigned)maxND, (unsigned)zc), 256, 0, ctx->stream>>>(dj);
      ctx->launches++;
    }
    // CTAs past the last row block of their column block, or outside a smaller job of the batch, leave at once
    if (count == 1) cov_factor_kernel_one<<<dim3((unsigned)maxRB, (unsigned)maxCB, 1), 256, 0, ctx->stream>>>(hj[0]);
    else cov_factor_kernel<<<dim3((unsigned)maxRB, (unsigned)maxCB, (unsigned)zc), 256, 0, ctx->stream>>>(dj);
    ctx->launches++;
    if (z0 + 65535 < count) CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));   // the job buffer is reused
  }
  CUDA_TRY(ctx, cudaGetLastError());
  for (int i = 0; i < count; i++) {
    const FactorBuildJob& b = jobs[i];
    const int rc = launch_ssre_add(ctx, b.a, b.d_L, b.n, b.n, 1, 1, b.ld / UR, 0);
    if (rc != 0) return rc;
  }
  return 0;
}

int launch_cov_factor_layout(iak3d_ctx* ctx, const CovArgs& h, const double* d_W, int32_t q, int64_t n, int64_t Npad, int64_t ld,
                             int32_t nCB_, double* d_L, const double* d_lncol, int lncol) {
  FactorBuildJob b;
  b.a = h; b.d_W = d_W; b.q = q; b.n = n; b.Npad = Npad; b.ld = ld; b.nCB = nCB_; b.d_L = d_L; b.d_lncol = d_lncol; b.lncol = lncol;
  return launch_cov_factor_jobs(ctx, &b, 1);
}

// ---- rectangular / full build (setCIAK3D returning C, setCIAK3D2, prediction row panels) -------------
// out is column-major with leading dimension ld; rows >= nr and cols >= nc are zero-filled up to
// rows_total x cols_total.  symmetric_square adds cme on the diagonal (row index == col index).
__global__ void __launch_bounds__(256) cov_rect_kernel(CovDev a, double* __restrict__ out, int64_t ld, int64_t rows_total,
                                                       int64_t cols_total, int sym, int blocked) {
  __shared__ int32_t s_xc[TN], s_dc[TN];
  const int64_t R0 = (int64_t)blockIdx.x * TM, C0 = (int64_t)blockIdx.y * TN;
  const int tid = threadIdx.x;
  if (tid < TN) { const int64_t c = C0 + tid; s_xc[tid] = c < a.nc ? a.xid_c[c] : -1; s_dc[tid] = c < a.nc ? a.did_c[c] : 0; }
  __syncthreads();
  const int rp = tid & 63, cg = tid >> 6;
  const int64_t i0 = R0 + rp * 2;
  int xi[2], di[2];
#pragma unroll
  for (int e = 0; e < 2; e++) { const int64_t i = i0 + e; xi[e] = i < a.nr ? a.xid_r[i] : -1; di[e] = i < a.nr ? a.did_r[i] : 0; }
  for (int cc = 0; cc < 16; cc++) {
    const int lc = cg * 16 + cc;
    const int64_t c = C0 + lc;
    if (c >= cols_total) break;
    const int xj = s_xc[lc], dj = s_dc[lc];
#pragma unroll
    for (int e = 0; e < 2; e++) {
      const int64_t i = i0 + e;
      if (i >= rows_total) continue;
      double val = 0.0;
      if (i < a.nr && c < a.nc) val = cov_value(a, xi[e], di[e], xj, dj, sym && i == c);
      if (blocked) out[uidx(i, c, rows_total >> 6)] = val; else out[c * ld + i] = val;
    }
  }
}

// Kriging panel (blocked layout): every row is a prediction support at its own location (xid_r[i] == i), so the two
// rows a thread owns are adjacent in the phi_x / same-location tables as well as in the output unit: 16-byte loads and one
// 16-byte evict-first store per column.  Same term order as cov_value.  nr and a.nxU_r even (the caller checks).
__global__ void __launch_bounds__(256) cov_panel_kernel(CovDev a, double* __restrict__ out, int64_t rows_total, int64_t cols_total) {
  __shared__ int32_t s_xc[TN], s_dc[TN];
  const int64_t R0 = (int64_t)blockIdx.x * TM, C0 = (int64_t)blockIdx.y * TN;
  const int tid = threadIdx.x;
  if (tid < TN) { const int64_t c = C0 + tid; s_xc[tid] = c < a.nc ? a.xid_c[c] : -1; s_dc[tid] = c < a.nc ? a.did_c[c] : 0; }
  __syncthreads();
  const uint64_t strm = policy_evict_first();
  const int rp = tid & 63, cg = tid >> 6;
  const int64_t i0 = R0 + rp * 2;
  if (i0 >= rows_total) return;
  const bool live = i0 < a.nr;                              // nr even: both rows or neither
  const int d0 = live ? a.did_r[i0] : 0, d1 = live ? a.did_r[i0 + 1] : 0;
  const int64_t nRU = rows_total >> 6;
  auto pick = [&](const double (&t)[3], int k) -> double { return k == 0 ? t[0] : (k == 1 ? t[1] : (k == 2 ? t[2] : 0.0)); };
#pragma unroll 4
  for (int cc = 0; cc < 16; cc++) {
    const int lc = cg * 16 + cc;
    const int64_t c = C0 + lc;
    if (c >= cols_total) break;
    double v0 = 0.0, v1 = 0.0;
    const int xj = s_xc[lc], dj = s_dc[lc];
    if (live && xj >= 0) {
      double ta[3], tb[3];
#pragma unroll
      for (int k = 0; k < 3; k++) {
        ta[k] = (k < a.ntab) ? __ldg(a.T[k] + (int64_t)dj * a.ndIU_r + d0) : 0.0;
        tb[k] = (k < a.ntab) ? ((d1 == d0) ? ta[k] : __ldg(a.T[k] + (int64_t)dj * a.ndIU_r + d1)) : 0.0;
      }
      const uchar2 sm = *reinterpret_cast<const uchar2*>(a.same + (int64_t)xj * a.nxU_r + i0);
      if (sm.x) { v0 = v0 + a.cx0; v0 = v0 + a.cxd0 * pick(ta, a.tidx[1]); }
      if (sm.y) { v1 = v1 + a.cx0; v1 = v1 + a.cxd0 * pick(tb, a.tidx[1]); }
      if (a.tidx[0] >= 0) { v0 = v0 + a.kd1 * pick(ta, a.tidx[0]); v1 = v1 + a.kd1 * pick(tb, a.tidx[0]); }
      if (a.phix != nullptr) {
        const double2 px = __ldg(reinterpret_cast<const double2*>(a.phix + (int64_t)xj * a.nxU_r + i0));
        v0 = v0 + a.cx1 * px.x; v0 = v0 + (a.cxd1 * px.x) * pick(ta, a.tidx[2]);
        v1 = v1 + a.cx1 * px.y; v1 = v1 + (a.cxd1 * px.y) * pick(tb, a.tidx[2]);
      }
    }
    st_stream2(out + uidx(i0, c, nRU), v0, v1, strm);
  }
}

int launch_cov_panel(iak3d_ctx* ctx, const CovArgs& h, double* d_out, int64_t rows_total, int64_t cols_total) {
  if ((h.nr & 1) || (h.nxU_r & 1) || h.same == nullptr) return launch_cov_rect(ctx, h, d_out, rows_total, rows_total, cols_total, 0, 1);
  CovDev a;
  for (int k = 0; k < 3; k++) { a.T[k] = h.T[k]; a.tidx[k] = h.tidx[k]; }
  a.ntab = h.ntab; a.phix = h.phix; a.same = h.same;
  a.cx0 = h.cx0; a.cx1 = h.cx1; a.kd1 = h.kd1; a.cxd0 = h.cxd0; a.cxd1 = h.cxd1; a.cme = h.cme;
  a.xid_r = h.xid_r; a.did_r = h.did_r; a.xid_c = h.xid_c; a.did_c = h.did_c;
  a.nr = h.nr; a.nc = h.nc; a.nxU_r = h.nxU_r; a.ndIU_r = h.ndIU_r;
  a.comb = nullptr; a.one_table = 0; a.E = nullptr; a.ldE = 0;
  if (rows_total == 0 || cols_total == 0) return 0;
  dim3 grid((unsigned)((rows_total + TM - 1) / TM), (unsigned)((cols_total + TN - 1) / TN));
  if (grid.y > 65535) { ctx->err = "cov_panel: too many column tiles"; return IAK3D_ERR_ARG; }
  cov_panel_kernel<<<grid, 256, 0, ctx->stream>>>(a, d_out, rows_total, cols_total);
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  return launch_ssre_add(ctx, h, d_out, h.nr, h.nc, 0, 1, rows_total / UR, 0);
}

int launch_cov_rect(iak3d_ctx* ctx, const CovArgs& h, double* d_out, int64_t ld, int64_t rows_total, int64_t cols_total,
                    int symmetric_square, int blocked) {
  CovDev a;
  for (int k = 0; k < 3; k++) { a.T[k] = h.T[k]; a.tidx[k] = h.tidx[k]; }
  a.ntab = h.ntab; a.phix = h.phix; a.same = h.same;
  a.cx0 = h.cx0; a.cx1 = h.cx1; a.kd1 = h.kd1; a.cxd0 = h.cxd0; a.cxd1 = h.cxd1; a.cme = h.cme;
  a.xid_r = h.xid_r; a.did_r = h.did_r; a.xid_c = h.xid_c; a.did_c = h.did_c;
  a.nr = h.nr; a.nc = h.nc; a.nxU_r = h.nxU_r; a.ndIU_r = h.ndIU_r;
  a.comb = nullptr; a.one_table = 0; a.E = nullptr; a.ldE = 0;
  if (rows_total == 0 || cols_total == 0) return 0;
  dim3 grid((unsigned)((rows_total + TM - 1) / TM), (unsigned)((cols_total + TN - 1) / TN));
  if (grid.y > 65535) { ctx->err = "cov_rect: too many column tiles"; return IAK3D_ERR_ARG; }
  cov_rect_kernel<<<grid, 256, 0, ctx->stream>>>(a, d_out, ld, rows_total, cols_total, symmetric_square, blocked);
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  return launch_ssre_add(ctx, h, d_out, h.nr, h.nc, 0, blocked, rows_total / UR, ld);
}

// ---- generic symmetric input (lndetANDinvCb called with a caller-built C, optim_functions.R:268) ------
__global__ void __launch_bounds__(256) pack_factor_kernel(const double* __restrict__ C, int64_t n, const double* __restrict__ b, int q,
                                                          int64_t Npad, int64_t ld, double* __restrict__ L) {
  const int64_t c = blockIdx.y;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < ld; i += (int64_t)gridDim.x * blockDim.x) {
    double v = 0.0;
    if (c < n) {
      if (i < n) v = (i >= c) ? C[c * n + i] : 0.0;                 // lower triangle of the caller's matrix
      else if (i >= Npad && i < Npad + q) v = b[(int64_t)(i - Npad) * n + c];
    } else {
      v = (i == c) ? 1.0 : 0.0;
    }
    L[uidx(i, c, ld >> 6)] = v;
  }
}
int launch_pack_factor_layout(iak3d_ctx* ctx, const double* d_C, int64_t n, const double* d_b, int32_t q, int64_t Npad,
                              int64_t ld, double* d_L) {
  if (Npad > 65535 * 64) { ctx->err = "pack_factor: matrix too large"; return IAK3D_ERR_ARG; }
  dim3 grid((unsigned)((ld + 255) / 256), (unsigned)Npad);
  if (Npad > 65535) { ctx->err = "pack_factor: more than 65535 columns"; return IAK3D_ERR_ARG; }
  pack_factor_kernel<<<grid, 256, 0, ctx->stream>>>(d_C, n, d_b, q, Npad, ld, d_L);
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  return 0;
}

// ---- blocked buffer -> column-major (parity tests of the factor build and of the kriging panel) ----------------------
// out[c*ld + i] = B(r0 + i, c) for i < nr, c < nc; mirror != 0: the entry of the lower triangle, B(max, min)
__global__ void unblock_rect_kernel(const double* __restrict__ B, int64_t nRU, int64_t r0, int64_t nr, int64_t nc, int mirror,
                                    double* __restrict__ out, int64_t ld) {
  const int64_t c = blockIdx.y;
  if (c >= nc) return;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nr; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t gi = r0 + i;
    out[c * ld + i] = (mirror && gi < c) ? B[uidx(c, gi, nRU)] : B[uidx(gi, c, nRU)];
  }
}
int launch_unblock_rect(iak3d_ctx* ctx, const double* d_B, int64_t nRU, int64_t r0, int64_t nr, int64_t nc, int mirror, double* d_out,
                        int64_t ld) {
  if (nr <= 0 || nc <= 0) return 0;
  if (nc > 65535) { ctx->err = "unblock_rect: more than 65535 columns"; return IAK3D_ERR_ARG; }
  unsigned bx = (unsigned)((nr + 255) / 256);
  if (bx > 256) bx = 256;
  unblock_rect_kernel<<<dim3(bx, (unsigned)nc), 256, 0, ctx->stream>>>(d_B, nRU, r0, nr, nc, mirror, d_out, ld);
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  return 0;
}

// ---- sigma2Vec (fitIAK3D.R:1083-1086) ------------------------------------------------------------------
struct S2Dev { double cxd0, cxd1, cme, cx0, cx1, cd1; const double *av0, *av1, *avd; int approx; };
__global__ void sigma2_kernel(S2Dev a, const int32_t* __restrict__ did, int64_t n, double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int d = did[i];
  const double a0 = a.av0 ? a.av0[d] : 0.0;
  const double a1 = a.av1 ? a.av1[d] : 0.0;
  const double ad = a.avd ? a.avd[d] : 0.0;
  double v;
  if (a.approx) {                    // setCApproxIAK3D's order (fitIAK3D.R:2056)
    v = a.cx0 + a.cx1;
    v = v + a.cd1 * ad; v = v + a.cxd0 * a0; v = v + a.cxd1 * a1;
    v = v + a.cme;
  } else {                           // setCIAK3D's order (fitIAK3D.R:1083)
    v = a.cxd0 * a0;
    v = v + a.cxd1 * a1;
    v = v + a.cme; v = v + a.cx0; v = v + a.cx1;
    v = v + a.cd1 * ad;
  }
  out[i] = v;
}
int launch_sigma2(iak3d_ctx* ctx, const EvalPlan& pl, const double* const* d_av, const int32_t* d_did, int64_t n, double* d_out) {
  if (n == 0) return 0;
  auto pick = [&](int t) -> const double* { return t >= 0 ? d_av[t] : nullptr; };
  S2Dev a{pl.cxd0, pl.cxd1, pl.cme, pl.cx0, pl.cx1, pl.cd1, pick(pl.tidx[1]), pick(pl.tidx[2]), pick(pl.tidx[0]), pl.approx ? 1 : 0};
  sigma2_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(a, d_did, n, d_out);
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  return 0;
}

__global__ void vec_sub_kernel(const double* __restrict__ a, const double* __restrict__ b, int64_t n, double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] - b[i];
}
int launch_vec_sub(iak3d_ctx* ctx, const double* d_a, const double* d_b, int64_t n, double* d_out) {
  if (n == 0) return 0;
  vec_sub_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(d_a, d_b, n, d_out);
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  return 0;
}

// ---- Ckk = diag(setCIAK3D(map supports)) (predictIAK3D.R:179-189): every term at zero separation --------
struct CkkDev { const double* T[3]; int tidx[3]; double cx0, cx1, kd1, cxd0, cxd1, cme; int has_phix; HxPrm H; };
__global__ void ckk_kernel(CkkDev a, int64_t ndIU1, const int32_t* __restrict__ did1, int64_t m, double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const int64_t d = did1[i];
  const double td = a.tidx[0] >= 0 ? a.T[a.tidx[0]][d * ndIU1 + d] : 0.0;
  const double t0 = a.tidx[1] >= 0 ? a.T[a.tidx[1]][d * ndIU1 + d] : 0.0;
  const double t1 = a.tidx[2] >= 0 ? a.T[a.tidx[2]][d * ndIU1 + d] : 0.0;
  double v = a.cme;
  v = v + a.cx0; v = v + a.cxd0 * t0;
  if (a.tidx[0] >= 0) v = v + a.kd1 * td;
  if (a.has_phix) {
    const double px = iak_phix(0.0, &a.H);
    v = v + a.cx1 * px;
    v = v + (a.cxd1 * px) * t1;
  }
  out[i] = v;
}
int launch_ckk(iak3d_ctx* ctx, const EvalPlan& pl, const double* const* d_Tself, int64_t ndIU1, const int32_t* d_did1, int64_t m,
               double* d_out) {
  if (m == 0) return 0;
  CkkDev a;
  for (int k = 0; k < 3; k++) { a.T[k] = d_Tself[k]; a.tidx[k] = pl.tidx[k]; }
  a.cx0 = pl.cx0; a.cx1 = pl.cx1; a.kd1 = pl.kd1; a.cxd0 = pl.cxd0; a.cxd1 = pl.cxd1; a.cme = pl.cme;
  a.has_phix = pl.has_phix ? 1 : 0; a.H = pl.H;
  ckk_kernel<<<(unsigned)((m + 255) / 256), 256, 0, ctx->stream>>>(a, ndIU1, d_did1, m, d_out);
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  return 0;
}

// ---- kriging epilogue (predictIAK3D.R:229-255) ---------------------------------------------------------
//   G(:,0..p-1) = Ckh iC X, G(:,p) = Ckh iC (z - X betahat), G(:,p+1) = rowSums((Ckh iC) * Ckh)
//   z = XMap betahat + G(:,p);  v = Ckk - G(:,p+1) + (XMap - CkhiCX) vbetahat (XMap - CkhiCX)'
__global__ void krige_finish_kernel(const double* __restrict__ G, int64_t ldG, int q, const double* __restrict__ XMap, int64_t ldX,
                                    int64_t m, const double* __restrict__ Ckk, const double* __restrict__ beta,
                                    const double* __restrict__ vbeta, double* __restrict__ z, double* __restrict__ v,
                                    double* __restrict__ cz, double* __restrict__ qf, double* __restrict__ cx, int64_t ldcx) {
  extern __shared__ double sh[];
  const int p = q - 1;
  double* sb = sh; double* sv = sh + p;
  for (int i = threadIdx.x; i < p; i += blockDim.x) sb[i] = beta[i];
  for (int i = threadIdx.x; i < p * p; i += blockDim.x) sv[i] = vbeta[i];
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  double mu = 0.0;
  for (int a = 0; a < p; a++) mu = mu + XMap[(int64_t)a * ldX + i] * sb[a];
  z[i] = mu + G[(int64_t)p * ldG + i];
  double quad = 0.0;
  for (int a = 0; a < p; a++) {
    double s = 0.0;
    for (int b = 0; b < p; b++) s = s + sv[(int64_t)b * p + a] * (XMap[(int64_t)b * ldX + i] - G[(int64_t)b * ldG + i]);
    quad = quad + (XMap[(int64_t)a * ldX + i] - G[(int64_t)a * ldG + i]) * s;
  }
  v[i] = Ckk[i] - G[(int64_t)q * ldG + i] + quad;
  if (cz != nullptr) cz[i] = G[(int64_t)p * ldG + i];
  if (qf != nullptr) qf[i] = G[(int64_t)q * ldG + i];
  if (cx != nullptr) for (int a = 0; a < p; a++) cx[(int64_t)a * ldcx + i] = G[(int64_t)a * ldG + i];
}
int launch_krige_finish(iak3d_ctx* ctx, const double* d_G, int64_t ldG, int32_t q, const double* d_XMap, int64_t ldX, int64_t m,
                        const double* d_Ckk, const double* d_beta, const double* d_vbeta, double* d_z, double* d_v, double* d_cz,
                        double* d_qf, double* d_cx, int64_t ldcx) {
  if (m == 0) return 0;
  const int p = q - 1;
  krige_finish_kernel<<<(unsigned)((m + 127) / 128), 128, (size_t)(p + p * p) * sizeof(double), ctx->stream>>>(
      d_G, ldG, q, d_XMap, ldX, m, d_Ckk, d_beta, d_vbeta, d_z, d_v, d_cz, d_qf, d_cx, ldcx);
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  return 0;
}
