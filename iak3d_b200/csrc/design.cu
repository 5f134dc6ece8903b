// design.cu -- the fixed-effect design matrix of depth-interval supports on the device: makeXvX (makeXvX.R:1-708) for
// lnTfmdData = FALSE, the call predictIAK3D makes once per (depth, batch) (predictIAK3D.R:155, :405) and fitIAK3D once per
// fit (fitIAK3D.R:212).  One thread per support; the design of a 10^6-cell x 6-interval grid is made from the per-cell
// covariates (48 MB) instead of being uploaded row by row (384 MB).
//
//   mlr branch    (makeXvX.R:212-353, "quicker version" :294-353): base(covariates) x depth moment of the interval, B-spline-of-depth
//                 columns averaged over nDiscPts points of seq(dU, dL, by)
//   cubist branch (makeXvX.R:360-434 with cubist2XGivenSetup, cubist2XIAK3D.R:46-176): the rule design is piecewise linear in
//                 depth, averaged by trapezoids between the rules' depth breaks (evaluated at break -/+ 1e-4); spline columns
//                 (allKnotsd2X, cubist2XIAK3D.R:187-208: bs() or the clamped natural spline) averaged over nDiscPts points
// Reference quirks reproduced: the d2 / d3 moments lack their 1/3 and 1/4 (:303, :308); the mlr limits are recycled from the
// first column of XLims (:327-328; clampcol[] is made on the host); limits never move exact zeros.
// Compiled with -fmad=false: products and sums round like the reference's operation order.
#include <math.h>

#include <algorithm>

#include "common.cuh"

constexpr int DES_MAXP = 64, DES_MAXRULES = 64, DES_MAXCONDS = 256, DES_MAXCOEF = 256, DES_MAXSET = 256, DES_MAXBRK = 64;
constexpr int DES_MAXKNOT = 14, DES_MAXSPL = 16, DES_MAXCOMM = 32;

struct DesignDev {
  int32_t type;            // 0 mlr, 1 cubist
  int32_t p, ncov;
  int32_t nspl, spl_kind;  // spline columns; 0 = bs (cubic B-spline, no intercept), 1 = makeXnsClampedUG0 without its constant
  int32_t nint;            // internal knots
  int32_t nDiscPts;
  int32_t has_lims, infBnds;
  double b1, b2;
  double tvec[DES_MAXKNOT + 8];     // bs knot vector: b1 x4, internal, b2 x4
  double ik[DES_MAXKNOT];
  double lims[2][DES_MAXP];
  // mlr
  int32_t kind[DES_MAXP];           // 0 no depth dependence, 1 d, 2 d2, 3 d3, 4 dSpline
  int32_t cov[DES_MAXP];            // covariate column (-1 = const); kind 4: spline column index
  double level[DES_MAXP];           // factor dummy: the code this column indicates; NaN = numeric covariate
  int32_t clampcol[DES_MAXP];       // position k of the column among the depth columns (id, idInt, id2, id2Int, id3, id3Int)
  int32_t ndepth;                   // their number K: entry (row i, k) is clamped by limit column (i K + k) mod p (makeXvX.R:327-328)
  // cubist
  int32_t pc;                       // rule columns (the spline columns follow)
  int32_t nRules, dcol, ncomm_unique, nbreaks;
  int32_t cond_off[DES_MAXRULES + 1], col0[DES_MAXRULES], ncoef[DES_MAXRULES], has_const[DES_MAXRULES], coef_off[DES_MAXRULES + 1];
  int32_t cond_var[DES_MAXCONDS], cond_dir[DES_MAXCONDS], cond_set_off[DES_MAXCONDS], cond_set_len[DES_MAXCONDS];
  double cond_val[DES_MAXCONDS];
  double setvals[DES_MAXSET];
  int32_t coef_iv[DES_MAXCOEF];
  int32_t col_blk[DES_MAXP];        // committee block of each rule column (-1: not normalised)
  int32_t blk_r0[DES_MAXCOMM], blk_r1[DES_MAXCOMM];   // rules [r0, r1) of a committee block
  double breaks[DES_MAXBRK];
};

struct iak3d_design {
  iak3d_ctx* ctx = nullptr;
  DesignDev h;
  DesignDev* d = nullptr;
};

namespace {

__device__ __forceinline__ double nanv() { return __longlong_as_double(0x7ff8000000000000LL); }

// splines::bs(x, knots, degree = 3, intercept = F, Boundary.knots): the nint + 3 columns at one depth
__device__ void bs_point(const DesignDev& D, double x, double* __restrict__ out) {
  const int nb = D.nint + 4;
  if (!(x >= D.b1 && x <= D.b2)) { for (int j = 0; j < nb - 1; j++) out[j] = nanv(); return; }
  for (int j = 0; j < nb - 1; j++) out[j] = 0.0;
  if (x == D.b2) { out[nb - 2] = 1.0; return; }
  int mu = 3;
  while (mu < nb - 1 && !(x < D.tvec[mu + 1])) mu++;
  double N[4], left[4], right[4];
  N[0] = 1.0;
#pragma unroll
  for (int jd = 1; jd <= 3; jd++) {
    left[jd] = x - D.tvec[mu + 1 - jd];
    right[jd] = D.tvec[mu + jd] - x;
    double saved = 0.0;
#pragma unroll
    for (int r = 0; r < jd; r++) {
      const double temp = N[r] / (right[r + 1] + left[jd - r]);
      N[r] = saved + right[r + 1] * temp;
      saved = left[jd - r] * temp;
    }
    N[jd] = saved;
  }
#pragma unroll
  for (int r = 0; r < 4; r++) { const int c = mu - 3 + r - 1; if (c >= 0) out[c] = N[r]; }
}

// makeXnsClampedUG0(x, bdryKnots, intKnots)[, -1] at one depth (splineBasisFns.R:204-217, paramVs = 0)
__device__ void ns_point(const DesignDev& D, double x, double* __restrict__ out) {
  const double b1 = D.b1, b2 = D.b2;
  auto cube = [&](double k) { const double v = x - k; return v < 0.0 ? 0.0 : pow(v, 3.0); };
  const double cL = cube(b1), cU = cube(b2);
  for (int j = 0; j < D.nint; j++) {
    const double k = D.ik[j];
    out[j] = cube(k) - ((b2 - k) / (b2 - b1)) * cL + ((b1 - k) / (b2 - b1)) * cU + 3 * (b2 - k) * (k - b1) * x;
  }
}

__device__ __forceinline__ void spline_point(const DesignDev& D, double x, double* __restrict__ out) {
  if (D.spl_kind == 0) bs_point(D, x, out); else ns_point(D, x, out);
}

__device__ __forceinline__ double clamp_nz(double v, double lo, double hi) {
  if (v < lo && v != 0.0) v = lo;            // replaceLT(replaceZeros = FALSE), then replaceGT on the result
  if (v > hi && v != 0.0) v = hi;
  return v;
}

// one row of cubist2XGivenSetup at depth d: x[0..pc) rule columns (normalised), x[pc..p) spline columns
__device__ void cubist_row(const DesignDev& D, const double* __restrict__ covs, int64_t ldc, int64_t ci, double d, double* __restrict__ x) {
  unsigned long long inrule = 0ull;
  for (int r = 0; r < D.nRules; r++) {
    bool in = true;
    if (D.nRules > 1) {
      for (int c = D.cond_off[r]; c < D.cond_off[r + 1]; c++) {
        const int v = D.cond_var[c];
        const double val = (v == D.dcol) ? d : covs[(int64_t)v * ldc + ci];
        if (D.cond_dir[c] == 0) in = in && (val <= D.cond_val[c]);
        else if (D.cond_dir[c] == 1) in = in && (val > D.cond_val[c]);
        else {
          bool any = false;
          for (int s = 0; s < D.cond_set_len[c]; s++) any = any || (val == D.setvals[D.cond_set_off[c] + s]);
          in = in && any;
        }
      }
    }
    if (in) inrule |= (1ull << r);
  }
  for (int j = 0; j < D.pc; j++) x[j] = 0.0;
  for (int r = 0; r < D.nRules; r++) {
    if (!((inrule >> r) & 1ull)) continue;
    int j = D.col0[r];
    if (D.has_const[r]) x[j++] = 1.0;
    for (int k = D.coef_off[r]; k < D.coef_off[r + 1]; k++) {
      const int v = D.coef_iv[k];
      x[j++] = (v == D.dcol) ? d : covs[(int64_t)v * ldc + ci];
    }
  }
  if (D.nRules > 1) {
    for (int j = 0; j < D.pc; j++) {
      const int b = D.col_blk[j];
      if (b >= 0) {
        const int r0 = D.blk_r0[b], r1 = D.blk_r1[b];
        const unsigned long long mask = ((r1 >= 64) ? ~0ull : ((1ull << r1) - 1ull)) & ~((1ull << r0) - 1ull);
        const int cnt = __popcll(inrule & mask);
        if (cnt > 0) x[j] = x[j] / (double)cnt;
      }
      x[j] = x[j] / (double)D.ncomm_unique;
    }
  }
  if (D.nspl > 0) spline_point(D, d, x + D.pc);
  bool bad = false;
  for (int j = 0; j < D.p; j++) bad = bad || isnan(x[j]);
  if (bad) for (int j = 0; j < D.p; j++) x[j] = nanv();
}

// Row i of the output is support (cell ci, interval ii): grid mode (ncell > 0) ii = i / ncell, ci = i % ncell with the
// interval taken from dIint; list mode ci = i with the row's own interval.
__global__ void __launch_bounds__(128) design_rows_kernel(const DesignDev* __restrict__ Dp, int64_t m, int64_t ncell, int64_t batch,
                                                          const double* __restrict__ covs, int64_t ldc,
                                                          const double* __restrict__ dI, int64_t lddI,
                                                          double* __restrict__ X, int64_t ldx,
                                                          double* __restrict__ splmin, double* __restrict__ splmax) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const DesignDev& D = *Dp;
  const int64_t ci = ncell > 0 ? i % ncell : i;
  const int64_t ii = ncell > 0 ? i / ncell : i;
  const double a = dI[ii], b = dI[lddI + ii];
  const bool clamp = D.has_lims && !D.infBnds;
  const int64_t irow = (batch > 0) ? ci % batch : ci;      // row inside the makeXvX call the reference would have made
  double sacc[DES_MAXSPL], sp[DES_MAXSPL], smn[DES_MAXSPL], smx[DES_MAXSPL];
  if (D.type == 0) {
    // ---- mlr ----
    if (D.nspl > 0) {
      for (int k = 0; k < D.nspl; k++) { sacc[k] = 0.0; smn[k] = INFINITY; smx[k] = -INFINITY; }
      int npts = 1;
      double by = 0.0;
      if (D.nDiscPts > 1) { by = (b - a) / (double)(D.nDiscPts - 1); npts = (int)floor((b - a) / by + 1e-10) + 1; }
      for (int t = 0; t < npts; t++) {
        double x = (D.nDiscPts > 1) ? a + (double)t * by : 0.5 * (a + b);
        if (D.nDiscPts > 1 && x > b) x = b;
        bs_point(D, x, sp);                                    // the mlr branch always uses bs() (makeXvX.R:334)
        for (int k = 0; k < D.nspl; k++) {
          double v = sp[k];
          if (v != 0.0) { smn[k] = fmin(smn[k], v); smx[k] = fmax(smx[k], v); }
          sacc[k] += v;
        }
      }
      if (clamp) {
        // limits apply to the point values (:341-346): redo the sum with clamped values
        for (int k = 0; k < D.nspl; k++) sacc[k] = 0.0;
        for (int t = 0; t < npts; t++) {
          double x = (D.nDiscPts > 1) ? a + (double)t * by : 0.5 * (a + b);
          if (D.nDiscPts > 1 && x > b) x = b;
          bs_point(D, x, sp);
          for (int j = 0; j < D.p; j++)
            if (D.kind[j] == 4) { const int k = D.cov[j]; sacc[k] += clamp_nz(sp[k], D.lims[0][j], D.lims[1][j]); }
        }
      }
      for (int k = 0; k < D.nspl; k++) sacc[k] = sacc[k] / (double)npts;
    }
    for (int j = 0; j < D.p; j++) {
      const int kd = D.kind[j];
      double v;
      if (kd == 4) {
        const int k = D.cov[j];
        v = sacc[k];
        if (splmin) { splmin[(int64_t)k * m + i] = smn[k]; splmax[(int64_t)k * m + i] = smx[k]; }
      } else {
        double base = 1.0;
        if (D.cov[j] >= 0) {
          const double c = covs[(int64_t)D.cov[j] * ldc + ci];
          base = isnan(D.level[j]) ? c : ((c == D.level[j]) ? 1.0 : 0.0);
        }
        if (kd == 0) v = base;
        else {
          double mom;
          if (kd == 1) mom = (a + b) / 2.0;
          else if (kd == 2) mom = a * a + a * b + b * b;
          else mom = pow(a, 3.0) + (a * a) * b + a * (b * b) + pow(b, 3.0);
          v = base * mom;
          if (clamp) {
            const int lc = (int)((irow * (int64_t)D.ndepth + D.clampcol[j]) % D.p);
            v = clamp_nz(v, D.lims[0][lc], D.lims[1][lc]);
          }
        }
      }
      X[(int64_t)j * ldx + i] = v;
    }
    return;
  }
  // ---- cubist ----
  double xt[DES_MAXP], xb[DES_MAXP], mw[DES_MAXP];
  const int p = D.p;
  cubist_row(D, covs, ldc, ci, a, xt);
  for (int j = 0; j < p; j++) mw[j] = 0.0;
  double lb = a;
  for (int k = 0; k < D.nbreaks; k++) {
    const double brk = D.breaks[k];
    if (a < brk && b > brk) {
      const double w = (brk - lb) / (b - a);
      cubist_row(D, covs, ldc, ci, brk - 0.0001, xb);
      for (int j = 0; j < p; j++) mw[j] = mw[j] + 0.5 * (xb[j] + xt[j]) * w;
      lb = brk;
      cubist_row(D, covs, ldc, ci, brk + 0.0001, xt);
    }
  }
  {
    const double w = (b - lb) / (b - a);
    cubist_row(D, covs, ldc, ci, b, xb);
    for (int j = 0; j < p; j++) mw[j] = mw[j] + 0.5 * (xb[j] + xt[j]) * w;
  }
  if (clamp) for (int j = 0; j < p; j++) mw[j] = clamp_nz(mw[j], D.lims[0][j], D.lims[1][j]);
  if (D.nspl > 0) {
    // set-limits callers want the trapezoid values of the spline columns: XLims is taken before they are replaced (:406-409)
    if (splmin) for (int k = 0; k < D.nspl; k++) splmin[(int64_t)k * m + i] = mw[D.pc + k];
    for (int k = 0; k < D.nspl; k++) sacc[k] = 0.0;
    for (int t = 1; t <= D.nDiscPts; t++) {
      const double x = a + ((double)(t - 1) / (double)(D.nDiscPts - 1)) * (b - a);
      spline_point(D, x, sp);
      for (int k = 0; k < D.nspl; k++) sacc[k] = sacc[k] + sp[k];
    }
    for (int k = 0; k < D.nspl; k++) mw[D.pc + k] = sacc[k] / (double)D.nDiscPts;
  }
  for (int j = 0; j < p; j++) X[(int64_t)j * ldx + i] = mw[j];
}

// ---- lnTfmdData = TRUE (makeXvX.R:246-290 'mlr', :437-557 'cubist'): a support is discretised into nDiscPts point depths; X =
// column means of the point designs, vXU = their sample covariance (R's var(): two passes, n - 1) on the columns iU that vary
// inside a support.  Limits act on the point values.  One thread per support; the covariance block is accumulated in the
// output itself (vXU is (m pU) x pU column-major: element (i pU + a, b)).
__device__ void point_row(const DesignDev& D, const double* __restrict__ covs, int64_t ldc, int64_t ci, double d, bool clamp, double* __restrict__ x,
                          double* __restrict__ sp) {
  if (D.type == 1) {
    cubist_row(D, covs, ldc, ci, d, x);
  } else {
    if (D.nspl > 0) bs_point(D, d, sp);
    for (int j = 0; j < D.p; j++) {
      const int kd = D.kind[j];
      if (kd == 4) { x[j] = sp[D.cov[j]]; continue; }
      double base = 1.0;
      if (D.cov[j] >= 0) {
        const double c = covs[(int64_t)D.cov[j] * ldc + ci];
        base = isnan(D.level[j]) ? c : ((c == D.level[j]) ? 1.0 : 0.0);
      }
      x[j] = kd == 0 ? base : (kd == 1 ? base * d : (kd == 2 ? base * (d * d) : base * pow(d, 3.0)));
    }
  }
  if (clamp) for (int j = 0; j < D.p; j++) x[j] = clamp_nz(x[j], D.lims[0][j], D.lims[1][j]);
}
__device__ __forceinline__ double point_depth(const DesignDev& D, double a, double b, int t, double by) {
  if (D.type == 1) return a + ((double)t * by) * (b - a);           // dIData[,1] + seq(0, 1, 1 / (nDiscPts - 1)) * (dL - dU), :497
  if (D.nDiscPts <= 1) return 0.5 * (a + b);
  const double x = a + (double)t * by;                              // seq(dU, dL, by), :250
  return x > b ? b : x;
}
__global__ void __launch_bounds__(128) design_lnN_kernel(const DesignDev* __restrict__ Dp, int64_t m, const double* __restrict__ covs, int64_t ldc,
                                                         const double* __restrict__ dI, const int32_t* __restrict__ iU, int pU,
                                                         double* __restrict__ X, double* __restrict__ vXU, double* __restrict__ pmin,
                                                         double* __restrict__ pmax) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const DesignDev& D = *Dp;
  const double a = dI[i], b = dI[m + i];
  const bool clamp = D.has_lims && !D.infBnds;
  int npts = D.nDiscPts; double by;
  if (D.type == 1) by = 1.0 / (double)(D.nDiscPts - 1);
  else if (D.nDiscPts > 1) { by = (b - a) / (double)(D.nDiscPts - 1); npts = (int)floor((b - a) / by + 1e-10) + 1; }
  else by = 0.0;
  double x[DES_MAXP], mean[DES_MAXP], sp[DES_MAXSPL];
  const int p = D.p;
  for (int j = 0; j < p; j++) mean[j] = 0.0;
  if (pmin) for (int j = 0; j < p; j++) { pmin[(int64_t)j * m + i] = nanv(); pmax[(int64_t)j * m + i] = nanv(); }
  for (int t = 0; t < npts; t++) {
    point_row(D, covs, ldc, i, point_depth(D, a, b, t, by), clamp, x, sp);
    for (int j = 0; j < p; j++) {
      mean[j] += x[j];
      if (pmin && x[j] != 0.0) {          // minNonZero / maxNonZero over the point values (:274-276)
        double* mn = pmin + (int64_t)j * m + i; double* mx = pmax + (int64_t)j * m + i;
        if (isnan(*mn) || x[j] < *mn) *mn = x[j];
        if (isnan(*mx) || x[j] > *mx) *mx = x[j];
      }
    }
  }
  for (int j = 0; j < p; j++) { mean[j] = mean[j] / (double)npts; X[(int64_t)j * m + i] = mean[j]; }
  const int64_t ldv = m * pU;
  for (int aa = 0; aa < pU; aa++) for (int bb = 0; bb < pU; bb++) vXU[(int64_t)bb * ldv + i * pU + aa] = 0.0;
  for (int t = 0; t < npts; t++) {
    point_row(D, covs, ldc, i, point_depth(D, a, b, t, by), clamp, x, sp);
    for (int aa = 0; aa < pU; aa++) {
      const double da = x[iU[aa]] - mean[iU[aa]];
      for (int bb = 0; bb < pU; bb++) vXU[(int64_t)bb * ldv + i * pU + aa] += da * (x[iU[bb]] - mean[iU[bb]]);
    }
  }
  for (int aa = 0; aa < pU; aa++) for (int bb = 0; bb < pU; bb++) vXU[(int64_t)bb * ldv + i * pU + aa] /= (double)(npts - 1);
}

// per-chunk (rows x 2) coordinate blocks and interval ids of a grid job (api_fit.cu's staging layout)
__global__ void grid_supports_kernel(const double* __restrict__ xyCell, int64_t ncell, int64_t m, int64_t chunk,
                                     double* __restrict__ xy, int32_t* __restrict__ did) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const int64_t ci = i % ncell, ii = i / ncell;
  const int64_t r0 = (i / chunk) * chunk;
  const int64_t rows = (m - r0 < chunk) ? (m - r0) : chunk;
  xy[2 * r0 + (i - r0)] = xyCell[ci];
  xy[2 * r0 + rows + (i - r0)] = xyCell[ncell + ci];
  did[i] = (int32_t)ii;
}

}  // namespace

// ---- host side: the flat specification --------------------------------------------------------------------------
// ispec / dspec are the integer and double parts of a DesignDev in declaration order (include/iak3d.h documents the layout);
// they are checked against the capacities here so that the kernel never indexes outside the structure.
extern "C" int iak3d_design_create(iak3d_ctx* ctx, const int32_t* is, int64_t nis, const double* dsp, int64_t nds, iak3d_design** out) {
  if (!ctx || !is || !dsp || !out) return IAK3D_ERR_ARG;
  *out = nullptr;
  cudaSetDevice(ctx->device);
  iak3d_design* g = new iak3d_design();
  g->ctx = ctx;
  DesignDev& D = g->h;
  memset(&D, 0, sizeof(D));
  int64_t ip = 0, dp = 0;
  bool ok = true;
  auto geti = [&]() -> int32_t { if (ip >= nis) { ok = false; return 0; } return is[ip++]; };
  auto getd = [&]() -> double { if (dp >= nds) { ok = false; return 0.0; } return dsp[dp++]; };
  auto fail = [&](const char* msg) { ctx->err = std::string("design_create: ") + msg; delete g; return IAK3D_ERR_ARG; };
  D.type = geti(); D.p = geti(); D.ncov = geti(); D.nspl = geti(); D.spl_kind = geti(); D.nint = geti(); D.nDiscPts = geti();
  D.has_lims = geti(); D.infBnds = geti();
  if (!ok || D.type < 0 || D.type > 1 || D.p < 1 || D.p > DES_MAXP || D.ncov < 0 || D.nspl < 0 || D.nspl > DES_MAXSPL || D.nint < 0 ||
      D.nint > DES_MAXKNOT || D.nDiscPts < 1 || (D.spl_kind != 0 && D.spl_kind != 1))
    return fail("sizes out of range (p <= 64, <= 14 internal knots)");
  if (D.nspl > 0 && D.nspl != (D.spl_kind == 0 ? D.nint + 3 : D.nint)) return fail("spline column count does not match the knots");
  if (D.type == 1 && D.nspl > 0 && D.nDiscPts < 2) return fail("nDiscPts >= 2 with spline columns");
  D.b1 = getd(); D.b2 = getd();
  for (int k = 0; k < D.nint; k++) D.ik[k] = getd();
  for (int k = 0; k < 4; k++) { D.tvec[k] = D.b1; D.tvec[4 + D.nint + k] = D.b2; }
  for (int k = 0; k < D.nint; k++) D.tvec[4 + k] = D.ik[k];
  for (int j = 0; j < D.p; j++) D.lims[0][j] = getd();
  for (int j = 0; j < D.p; j++) D.lims[1][j] = getd();
  if (D.type == 0) {
    for (int j = 0; j < D.p; j++) {
      D.kind[j] = geti(); D.cov[j] = geti(); D.clampcol[j] = geti(); D.level[j] = getd();
      if (D.kind[j] < 0 || D.kind[j] > 4) return fail("column kind");
      if (D.kind[j] == 4 ? (D.cov[j] < 0 || D.cov[j] >= D.nspl) : (D.cov[j] < -1 || D.cov[j] >= D.ncov)) return fail("column covariate index");
      if (D.clampcol[j] < -1 || D.clampcol[j] >= D.p) return fail("clamp position");
      if (D.clampcol[j] >= 0) D.ndepth++;
      if ((D.kind[j] >= 1 && D.kind[j] <= 3) != (D.clampcol[j] >= 0)) return fail("clamp position of a column without / with depth moment");
    }
  } else {
    D.pc = geti(); D.nRules = geti(); D.dcol = geti(); D.ncomm_unique = geti(); D.nbreaks = geti();
    const int nblk = geti();
    if (!ok || D.pc < 0 || D.pc + D.nspl != D.p || D.nRules < 1 || D.nRules > DES_MAXRULES || D.dcol < 0 || D.dcol >= D.ncov ||
        D.ncomm_unique < 1 || D.nbreaks < 0 || D.nbreaks > DES_MAXBRK || nblk < 0 || nblk > DES_MAXCOMM)
      return fail("cubist sizes out of range (<= 64 rules, <= 64 depth breaks)");
    int nc = 0, nk = 0, ns = 0, col = 0;
    for (int r = 0; r < D.nRules; r++) {
      D.has_const[r] = geti();
      const int ncond = geti(), ncoef = geti();
      if (!ok || ncond < 0 || ncoef < 0 || nc + ncond > DES_MAXCONDS || nk + ncoef > DES_MAXCOEF) return fail("too many rule conditions / coefficients");
      D.cond_off[r] = nc; D.coef_off[r] = nk; D.col0[r] = col; D.ncoef[r] = ncoef;
      for (int c = 0; c < ncond; c++, nc++) {
        D.cond_var[nc] = geti(); D.cond_dir[nc] = geti(); D.cond_set_len[nc] = geti(); D.cond_val[nc] = getd();
        D.cond_set_off[nc] = ns;
        if (!ok || D.cond_var[nc] < 0 || D.cond_var[nc] >= D.ncov || D.cond_dir[nc] < 0 || D.cond_dir[nc] > 2 || D.cond_set_len[nc] < 0 ||
            ns + D.cond_set_len[nc] > DES_MAXSET)
          return fail("rule condition");
        for (int s = 0; s < D.cond_set_len[nc]; s++) D.setvals[ns++] = getd();
      }
      for (int k = 0; k < ncoef; k++, nk++) {
        D.coef_iv[nk] = geti();
        if (!ok || D.coef_iv[nk] < 0 || D.coef_iv[nk] >= D.ncov) return fail("rule coefficient variable");
      }
      col += (D.has_const[r] ? 1 : 0) + ncoef;
    }
    D.cond_off[D.nRules] = nc; D.coef_off[D.nRules] = nk;
    if (col != D.pc) return fail("rule columns do not add up to pc");
    for (int j = 0; j < D.pc; j++) { D.col_blk[j] = geti(); if (D.col_blk[j] < -1 || D.col_blk[j] >= nblk) return fail("column block"); }
    for (int b = 0; b < nblk; b++) {
      D.blk_r0[b] = geti(); D.blk_r1[b] = geti();
      if (!ok || D.blk_r0[b] < 0 || D.blk_r1[b] < D.blk_r0[b] || D.blk_r1[b] > D.nRules) return fail("committee block");
    }
    for (int k = 0; k < D.nbreaks; k++) D.breaks[k] = getd();
  }
  if (!ok || ip != nis || dp != nds) return fail("specification length mismatch");
  if (cudaMalloc(&g->d, sizeof(DesignDev)) != cudaSuccess) { cudaGetLastError(); ctx->err = "design_create: allocation failed"; delete g; return IAK3D_ERR_CUDA; }
  cudaMemcpyAsync(g->d, &g->h, sizeof(DesignDev), cudaMemcpyHostToDevice, ctx->stream);
  if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { cudaFree(g->d); ctx->err = "design_create: copy failed"; delete g; return IAK3D_ERR_CUDA; }
  *out = g;
  return IAK3D_OK;
}

extern "C" int iak3d_design_destroy(iak3d_design* g) {
  if (!g) return IAK3D_OK;
  cudaSetDevice(g->ctx->device);
  cudaStreamSynchronize(g->ctx->stream);
  cudaFree(g->d);
  delete g;
  return IAK3D_OK;
}

extern "C" int iak3d_design_info(iak3d_design* g, int32_t* p, int32_t* ncov) {
  if (!g) return IAK3D_ERR_ARG;
  if (p) *p = g->h.p;
  if (ncov) *ncov = g->h.ncov;
  return IAK3D_OK;
}

// device-side entry used by the kriging stage (api_fit.cu)
int launch_design_rows(iak3d_ctx* ctx, const iak3d_design* g, int64_t m, int64_t ncell, int64_t batch, const double* d_covs, int64_t ldc,
                       const double* d_dI, int64_t lddI, double* d_X, int64_t ldx, double* d_splmin, double* d_splmax) {
  if (m <= 0) return 0;
  design_rows_kernel<<<(unsigned)((m + 127) / 128), 128, 0, ctx->stream>>>(g->d, m, ncell, batch, d_covs, ldc, d_dI, lddI, d_X, ldx, d_splmin, d_splmax);
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  return 0;
}
int launch_grid_supports(iak3d_ctx* ctx, const double* d_xyCell, int64_t ncell, int64_t m, int64_t chunk, double* d_xy, int32_t* d_did) {
  if (m <= 0) return 0;
  grid_supports_kernel<<<(unsigned)((m + 255) / 256), 256, 0, ctx->stream>>>(d_xyCell, ncell, m, chunk, d_xy, d_did);
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  return 0;
}
int design_p(const iak3d_design* g) { return g->h.p; }
int design_ncov(const iak3d_design* g) { return g->h.ncov; }
iak3d_ctx* design_ctx(const iak3d_design* g) { return g->ctx; }

extern "C" int iak3d_design_eval(iak3d_design* g, int64_t m, const double* covs, const double* dI, double* X, double* splmin, double* splmax) {
  if (!g || m < 0 || (m > 0 && (!dI || !X)) || (g->h.ncov > 0 && m > 0 && !covs)) return IAK3D_ERR_ARG;
  iak3d_ctx* ctx = g->ctx;
  cudaSetDevice(ctx->device);
  if (m == 0) return IAK3D_OK;
  const int p = g->h.p, ncov = g->h.ncov, nspl = g->h.nspl;
  double *d_c = nullptr, *d_dI = nullptr, *d_X = nullptr, *d_mn = nullptr, *d_mx = nullptr;
  int rc = IAK3D_OK;
  do {
    cudaError_t e = cudaSuccess;
    if (ncov > 0) e = iak_pool_alloc(ctx, (void**)&d_c, (size_t)m * ncov * 8);
    if (e == cudaSuccess) e = iak_pool_alloc(ctx, (void**)&d_dI, (size_t)m * 2 * 8);
    if (e == cudaSuccess) e = iak_pool_alloc(ctx, (void**)&d_X, (size_t)m * p * 8);
    const bool want_lim = (splmin && splmax && nspl > 0);
    if (e == cudaSuccess && want_lim) e = iak_pool_alloc(ctx, (void**)&d_mn, (size_t)m * nspl * 8);
    if (e == cudaSuccess && want_lim) e = iak_pool_alloc(ctx, (void**)&d_mx, (size_t)m * nspl * 8);
    if (e == cudaSuccess && ncov > 0) e = cudaMemcpyAsync(d_c, covs, (size_t)m * ncov * 8, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_dI, dI, (size_t)m * 2 * 8, cudaMemcpyHostToDevice, ctx->stream);
    if (e != cudaSuccess) { ctx->err = std::string("design_eval: ") + cudaGetErrorString(e); cudaGetLastError(); rc = IAK3D_ERR_CUDA; break; }
    rc = launch_design_rows(ctx, g, m, 0, 0, d_c, m, d_dI, m, d_X, m, d_mn, d_mx);
    if (rc != 0) break;
    e = cudaMemcpyAsync(X, d_X, (size_t)m * p * 8, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess && want_lim) e = cudaMemcpyAsync(splmin, d_mn, (size_t)m * nspl * 8, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess && want_lim) e = cudaMemcpyAsync(splmax, d_mx, (size_t)m * nspl * 8, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { ctx->err = std::string("design_eval: ") + cudaGetErrorString(e); cudaGetLastError(); rc = IAK3D_ERR_CUDA; }
  } while (0);
  iak_pool_free(ctx, d_c); iak_pool_free(ctx, d_dI); iak_pool_free(ctx, d_X); iak_pool_free(ctx, d_mn); iak_pool_free(ctx, d_mx);
  return rc;
}

extern "C" int iak3d_design_eval_lnN(iak3d_design* g, int64_t m, const double* covs, const double* dI, int32_t pU, const int32_t* iU, double* X,
                                     double* vXU, double* pmin, double* pmax) {
  if (!g || m < 0 || pU < 0 || (m > 0 && (!dI || !X || (pU > 0 && (!iU || !vXU)))) || (g->h.ncov > 0 && m > 0 && !covs)) return IAK3D_ERR_ARG;
  iak3d_ctx* ctx = g->ctx;
  cudaSetDevice(ctx->device);
  if (m == 0) return IAK3D_OK;
  const int p = g->h.p, ncov = g->h.ncov;
  for (int k = 0; k < pU; k++) if (iU[k] < 0 || iU[k] >= p) { ctx->err = "design_eval_lnN: iU out of range"; return IAK3D_ERR_ARG; }
  if (g->h.nDiscPts < 2) { ctx->err = "design_eval_lnN: nDiscPts >= 2"; return IAK3D_ERR_ARG; }
  double *d_c = nullptr, *d_dI = nullptr, *d_X = nullptr, *d_V = nullptr, *d_mn = nullptr, *d_mx = nullptr; int32_t* d_iU = nullptr;
  int rc = IAK3D_OK;
  const bool want_lim = pmin && pmax;
  do {
    cudaError_t e = cudaSuccess;
    if (ncov > 0) e = iak_pool_alloc(ctx, (void**)&d_c, (size_t)m * ncov * 8);
    if (e == cudaSuccess) e = iak_pool_alloc(ctx, (void**)&d_dI, (size_t)m * 16);
    if (e == cudaSuccess) e = iak_pool_alloc(ctx, (void**)&d_X, (size_t)m * p * 8);
    if (e == cudaSuccess) e = iak_pool_alloc(ctx, (void**)&d_V, (size_t)std::max<int64_t>(m * pU * pU, 1) * 8);
    if (e == cudaSuccess) e = iak_pool_alloc(ctx, (void**)&d_iU, (size_t)std::max(pU, 1) * 4);
    if (e == cudaSuccess && want_lim) e = iak_pool_alloc(ctx, (void**)&d_mn, (size_t)m * p * 8);
    if (e == cudaSuccess && want_lim) e = iak_pool_alloc(ctx, (void**)&d_mx, (size_t)m * p * 8);
    if (e == cudaSuccess && ncov > 0) e = cudaMemcpyAsync(d_c, covs, (size_t)m * ncov * 8, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_dI, dI, (size_t)m * 16, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess && pU > 0) e = cudaMemcpyAsync(d_iU, iU, (size_t)pU * 4, cudaMemcpyHostToDevice, ctx->stream);
    if (e != cudaSuccess) { ctx->err = std::string("design_eval_lnN: ") + cudaGetErrorString(e); cudaGetLastError(); rc = IAK3D_ERR_CUDA; break; }
    design_lnN_kernel<<<(unsigned)((m + 127) / 128), 128, 0, ctx->stream>>>(g->d, m, d_c, m, d_dI, d_iU, pU, d_X, d_V, d_mn, d_mx);
    ctx->launches++;
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(X, d_X, (size_t)m * p * 8, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess && pU > 0) e = cudaMemcpyAsync(vXU, d_V, (size_t)m * pU * pU * 8, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess && want_lim) e = cudaMemcpyAsync(pmin, d_mn, (size_t)m * p * 8, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess && want_lim) e = cudaMemcpyAsync(pmax, d_mx, (size_t)m * p * 8, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { ctx->err = std::string("design_eval_lnN: ") + cudaGetErrorString(e); cudaGetLastError(); rc = IAK3D_ERR_CUDA; }
  } while (0);
  iak_pool_free(ctx, d_c); iak_pool_free(ctx, d_dI); iak_pool_free(ctx, d_X); iak_pool_free(ctx, d_V); iak_pool_free(ctx, d_iU);
  iak_pool_free(ctx, d_mn); iak_pool_free(ctx, d_mx);
  return rc;
}
