// common.cuh -- internal structures of libiak3d.so (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/iak3d.h"
#include "iak3d_math.h"

// ---- tiling constants shared by the covariance-build and factorisation kernels ---------------
constexpr int TM = 128;      // tile rows  (row-block height)
constexpr int TN = 64;       // tile cols  (column-block width == panel width of the factorisation)
constexpr int INVD_LD = TN + 4;              // inverse diagonal blocks are kept with the padded shared-memory stride
constexpr int INVD_BLOCK = TN * INVD_LD;     // doubles per stored inverse block

// ---- blocked storage of factor buffers and kriging panels ---------------------------------------
// A matrix with R rows (multiple of 128) is stored as UNITS of 64 rows x 32 columns.  Inside a unit the 32
// columns are column-major with leading dimension UL = 68 -- exactly the bank-conflict-free layout the DMMA
// fragment loads want in shared memory -- so that one cp.async.bulk (TMA) of a whole unit (17,408 B), or of the two
// vertically adjacent units of a 128-row block (34,816 B), lands a pipeline stage without any reshaping.
// Units are ordered column-panel by column-panel: unit (ru, cu) sits at (cu * nRU + ru) * UNIT, nRU = R / 64.
constexpr int UR = 64, UC = 32, UL = 68;
constexpr int UNIT = UC * UL;                // 2176 doubles
#if defined(__CUDACC__)
__host__ __device__
#endif
static inline int64_t uidx(int64_t i, int64_t c, int64_t nRU) {
  return ((c >> 5) * nRU + (i >> 6)) * UNIT + (c & 31) * UL + (i & 63);
}
static inline int64_t ubuf_doubles(int64_t rows, int64_t cols) { return (cols / UC) * (rows / UR) * (int64_t)UNIT; }

static inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

struct iak3d_ctx {
  int device = 0;
  int sm_count = 0;
  int cc_major = 0, cc_minor = 0;
  size_t total_mem = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t evs[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // stage boundaries of the last batch
  std::string err;
  int64_t launches = 0;
  int pair_warps = 4;        // MMA warps per CTA of the pair engine
  int pair_min_width = 148;  // tasks per dependency step from which a throughput launch uses the CTA-pair engine
  bool no_table_queue = false;  // IAK3D_TABLE_QUEUE=0: one launch per table, as before (comparison)
  bool diag_chain = true;    // full-size CTAs: warp-specialised diagonal blocks in every factorisation launch (IAK3D_DIAG=0: only single-problem ones)
  int engine = 1;            // tile engine of throughput launches: 1 = one CTA per SM (chol.cu), 2 = CTA pairs (chol_pair.cuh)
  // scratch
  void* flush_buf = nullptr; size_t flush_bytes = 0;
  // batch state (nll_terms_batch_enqueue -> fetch)
  struct Batch* batch = nullptr;
  bool batch_pending() const;
  double last_chol_flops = 0, last_cov_bytes = 0;
  bool stage_events_valid = false;
  void* d_fbjobs = nullptr; size_t fbjobs_cap = 0;     // device copy of the covariance-build jobs of one launch
  void* d_tabjobs = nullptr; void* h_tabjobs = nullptr; size_t tabjobs_cap = 0;   // queued table jobs of one batch (device / pinned)
  long long* d_prof = nullptr; int64_t prof_cap = 0; int32_t prof_ntasks = 0; const int4* d_prof_tasks = nullptr;   // IAK3D_PROF=1 diagnostics
};

constexpr int NTAB_BUF = 4;   // depth-table buffers of a slot: <= 3 tables in use + the stationary source of the approximation

// Workspace of ONE likelihood evaluation of a dataset: the factor buffer, the unique-pair tables and the
// dependency flags.  A dataset owns as many slots as the largest number of simultaneous evaluations asked of
// it (1 for a composite-likelihood block, npar+1 for a forward-difference gradient).
struct Slot {
  int64_t n = 0, Npad = 0, ld = 0, nxU = 0, ndIU = 0;
  int32_t q = 0, nCB = 0, nRB = 0;
  double* d_L = nullptr;          // ld x Npad factor buffer, BLOCKED (uidx): A in, L out; W' rows at Npad..Npad+q-1
  double* d_phix = nullptr;       // nxU x nxU horizontal correlation table
  double* d_phid = nullptr;       // NTAB_BUF tables ndIU x ndIU + NTAB_BUF avVar vectors + 3 avsd vectors (spline sd)
  double* d_invd = nullptr;       // nCB x INVD_BLOCK inverted diagonal blocks
  int32_t* d_flags = nullptr;     // done[nRB*nCB] + invdone[nCB]
  double* d_logsum = nullptr;     // nCB
  double* d_G = nullptr;          // 64*64 Gram accumulator (WiAW)
  double* d_lncol = nullptr;      // 2n: sigma2Vec - diag(C) (lognormal bias column of W) | scratch
  double* d_E = nullptr;          // 2 * ndIU * round_up(n,16): depth tables expanded along the rows (factor build)
  int32_t epoch = 0;              // flags are valid when == epoch
};

// Device-resident dataset == setupMats (+ W) of the reference, expressed as id vectors.
struct iak3d_ds {
  iak3d_ctx* ctx = nullptr;
  int64_t n = 0, nxU = 0, ndIU = 0;
  int32_t q = 0;
  std::vector<int32_t> h_xid, h_did;
  std::vector<double> h_xU, h_dIU;       // nxU x 2, ndIU x 2 (column-major)
  std::vector<double> h_spl[3];          // XsdfdSplineU_{cd1,cxd0,cxd1}: ndIU x spl_ncol[c] column-major (iaCovMatern.R:520-539)
  int32_t spl_ncol[3] = {0, 0, 0};
  int32_t lncol = -1;                    // column of W replaced by sigma2Vec - diag(C) at every evaluation (lognormal data)
  int32_t nssre = 0;                     // site-specific random effects: columns of d_Xs, site id of every row
  int32_t* d_sid = nullptr; double* d_Xs = nullptr;
  double* h_W = nullptr;                 // n x q column-major copy in PINNED host memory: staging of the H2D copy
  bool w_set = false;                    // h_W holds what the device holds
  int32_t *d_xid = nullptr, *d_did = nullptr;
  double *d_xU = nullptr, *d_dIU = nullptr;   // column-major nxU x 2 / ndIU x 2
  double* d_W = nullptr;                  // n x q column-major
  // factorisation geometry (depends on n and q only)
  int64_t Npad = 0;     // round_up(n, TN): columns factorised (identity padded)
  int64_t ld = 0;       // rows of the factor buffer = round_up(Npad + q, TM)
  int32_t nCB = 0, nRB = 0;
  std::vector<Slot*> slots;
  std::vector<Slot*> tslots;            // table-only slots (perturbed parameter sets of the analytic gradient)
  double* d_gradU = nullptr; double* d_gradiC = nullptr; size_t grad_bytes = 0;   // analytic gradient: L^-T and the explicit inverse (blocked), kept between calls
};

// One covariance evaluation planned on the host from parsBTfmd (no device work).
struct EvalPlan {
  int status = 0;           // 0, IAK3D_BAD_PARS
  int ntab = 0;             // distinct depth tables
  IaPrm prm[3];
  int tidx[3] = {-1, -1, -1};   // table of cd1, cxd0, cxd1 (-1 absent)
  bool has_phix = false;
  HxPrm H;
  // setCApproxIAK3D[2] (any sdfdType > 0, fitIAK3D.R:959-962): every table is the stationary table scaled by the
  // interval-averaged sd of its component, avCor * avsd(I) * avsd(J) (:1960-1963).  Per TABLE k:
  bool approx = false;
  int sd_type[3] = {0, 0, 0};            // 0: the stationary table itself; -1 / k > 0: scaled by iasdfd (fitIAK3D.R:1448-1468)
  int sd_comp[3] = {0, 0, 0};            // component (0 cd1, 1 cxd0, 2 cxd1) whose spline basis applies
  double sd_coef[3][IAK3D_MAX_SPL] = {}; // (tau1, tau2) or the spline coefficients
  double cx0 = 0, cx1 = 0, cd1 = 0, kd1 = 0, cxd0 = 0, cxd1 = 0, cme = 0;
  int nssre = 0; double cssre[IAK3D_MAX_SSRE] = {};
};

#define CUDA_TRY(ctx, expr)                                                                   \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      char _b[512];                                                                           \
      snprintf(_b, sizeof(_b), "%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      (ctx)->err = _b;                                                                        \
      return IAK3D_ERR_CUDA;                                                                  \
    }                                                                                         \
  } while (0)

// ---- stream-ordered scratch memory (api.cu) ----
// Temporaries that come and go with every call (task lists, partial sums, device matrices): cudaMallocAsync on the
// context's stream from the device's default pool, kept cached, so that neither allocation nor release synchronises.
cudaError_t iak_pool_alloc(iak3d_ctx* ctx, void** p, size_t bytes);
void iak_pool_free(iak3d_ctx* ctx, void* p);
// cudaMalloc for the long-lived buffers; when it fails the cached scratch pool is handed back to the driver and the
// allocation retried once
cudaError_t iak_malloc(iak3d_ctx* ctx, void** p, size_t bytes);

// ---- table kernels (tables.cu) ----------------------------------------------------------------
// square tables of many problems, queued on the host and run as one launch per kind (tables.cu: flush_table_queue)
struct IaTabJob { IaPrm P; const double* dI; int64_t n; double* out; double* av; };
struct HxTabJob { HxPrm H; const double* x; int64_t n; double* out; };
struct TableQueue { std::vector<IaTabJob> ia; std::vector<HxTabJob> hx; };
int flush_table_queue(iak3d_ctx* ctx, TableQueue* tq);
int launch_iacov_table(iak3d_ctx* ctx, const IaPrm& P, const double* d_dI1, int64_t n1, const double* d_dI2,
                       int64_t n2, double* d_out /* n1 x n2 col-major: out[j*n1+i] */, double* d_avvar1);
int launch_phix_table(iak3d_ctx* ctx, const HxPrm& H, const double* d_x1, int64_t n1, const double* d_x2, int64_t n2,
                      double* d_out /* out[j*n1+i] */, uint8_t* d_zero /* optional: D==0 indicator */);
int launch_phix_vec(iak3d_ctx* ctx, const HxPrm& H, const double* d_D, int64_t m, double* d_out);
int hx_init(HxPrm* H, int modelx, double a, double nu);   // returns 2 when the reference returns NA

// ---- covariance assembly (covbuild.cu) --------------------------------------------------------
struct CovArgs {
  // value(i,j) = [same loc](cx0 + cxd0*T0) + kd1*Td + phix*(cx1 + cxd1*T1) + (i==j)*cme
  const double* T[3];      // distinct depth tables (may alias); T[tidx[c]] for c = cd1, cxd0, cxd1; tidx<0 = absent
  const double* av[3] = {nullptr, nullptr, nullptr};   // avVarVec of each table on the row set (square builds only)
  double* comb = nullptr;  // nd^2 doubles of per-evaluation workspace: S table of the factor build's generic path
  double* Ebuf = nullptr;  // 2 * nd * round_up(n,16) doubles of per-evaluation workspace: depth tables expanded along the rows
  const double* E_shared = nullptr;   // one-table case: the expansion another evaluation of the batch already made
  int tidx[3];
  int ntab;
  const double* phix;      // horizontal table or nullptr (nugget)
  const uint8_t* same;     // cross only: D==0 indicator table; square: nullptr (xid equality)
  double cx0, cx1, kd1, cxd0, cxd1, cme;
  const int32_t *xid_r, *did_r, *xid_c, *did_c;   // row ids / col ids
  int64_t nr, nc;          // live rows / cols
  int64_t nxU_r, ndIU_r;   // leading dims of tables (row-id extent)
  // site-specific random effects (listStructs4ssre, iaCovMatern.R:440-460, 645-663; added at fitIAK3D.R:1091-1101, 1233-1240):
  //   value(i, j) += [site_r(i) == site_c(j)] * sum_k cssre[k] Xs_r(i, k) Xs_c(j, k)
  int nssre = 0;
  double cssre[IAK3D_MAX_SSRE] = {};
  const int32_t *sid_r = nullptr, *sid_c = nullptr;
  const double *Xs_r = nullptr, *Xs_c = nullptr;   // column-major, leading dimensions ldXs_r / ldXs_c
  int64_t ldXs_r = 0, ldXs_c = 0;
};
// one covariance build into a factor buffer; a batch of them is ONE launch (launch_cov_factor_jobs)
struct FactorBuildJob {
  CovArgs a;
  const double* d_W = nullptr; int32_t q = 0; int64_t n = 0, Npad = 0, ld = 0; int32_t nCB = 0; double* d_L = nullptr;
  const double* d_lncol = nullptr; int lncol = -1;
};
int launch_cov_factor_jobs(iak3d_ctx* ctx, const FactorBuildJob* jobs, int count);
// d_lncol / lncol: column `lncol` of W is taken from the vector d_lncol instead (lognormal bias column, see api.cu)
int launch_cov_factor_layout(iak3d_ctx* ctx, const CovArgs& a, const double* d_W, int32_t q, int64_t n, int64_t Npad, int64_t ld,
                             int32_t nCB, double* d_L, const double* d_lncol = nullptr, int lncol = -1);
// out = a - b (n doubles)
bool cov_one_table(const CovArgs& a);   // every present component reads depth table 0
int launch_vec_sub(iak3d_ctx* ctx, const double* d_a, const double* d_b, int64_t n, double* d_out);
// generic symmetric matrix (host-supplied C, optim_functions.R:268) + b' rows into the factor layout
int launch_pack_factor_layout(iak3d_ctx* ctx, const double* d_C, int64_t n, const double* d_b, int32_t q, int64_t Npad,
                              int64_t ld, double* d_L);
int launch_sigma2(iak3d_ctx* ctx, const EvalPlan& pl, const double* const* d_av /* per table: avVarVec */, const int32_t* d_did,
                  int64_t n, double* d_out);
// setCApproxIAK3D[2]: out(i,j) = (src(i,j) * s1(i)) * s2(j) (rectangular, fitIAK3D.R:2146) or, symmetric, the packed-upper
// order of :1963  (src * s[max(i,j)]) * s[min(i,j)];  avvar1 = s1^2 when not null
int launch_scale_table(iak3d_ctx* ctx, const double* d_src, const double* d_s1, int64_t n1, const double* d_s2, int64_t n2,
                       int symmetric, double* d_out, double* d_avvar1);
int launch_ckk(iak3d_ctx* ctx, const EvalPlan& pl, const double* const* d_Tself /* per table: ndIU1 x ndIU1 */, int64_t ndIU1,
               const int32_t* d_did1, int64_t m, double* d_out);
// blocked != 0: d_out is a blocked buffer (uidx) with rows_total rows; otherwise plain column-major with leading dim ld
int launch_cov_rect(iak3d_ctx* ctx, const CovArgs& a, double* d_out, int64_t ld, int64_t rows_total, int64_t cols_total,
                    int symmetric_square, int blocked = 0);
// kriging panel: blocked output, rows = prediction supports at their own locations (xid_r[i] == i); falls back to
// launch_cov_rect when the row count is odd
int launch_cov_panel(iak3d_ctx* ctx, const CovArgs& a, double* d_out, int64_t rows_total, int64_t cols_total);

// diag[i] += sum_k cssre[k] Xs(i, k)^2  (the ssre part of a diagonal: Ckk of map supports, diag(C))
int launch_ssre_diag(iak3d_ctx* ctx, int nssre, const double* cssre, const double* d_Xs, int64_t ldXs, int64_t m, double* d_diag);
int launch_unblock_rect(iak3d_ctx* ctx, const double* d_B, int64_t nRU, int64_t r0, int64_t nr, int64_t nc, int mirror, double* d_out,
                        int64_t ld);

// ---- fixed-effect design matrix (design.cu) ------------------------------------------------------
struct iak3d_design;
int launch_design_rows(iak3d_ctx* ctx, const iak3d_design* g, int64_t m, int64_t ncell, int64_t batch, const double* d_covs, int64_t ldc,
                       const double* d_dI, int64_t lddI, double* d_X, int64_t ldx, double* d_splmin, double* d_splmax);
int launch_grid_supports(iak3d_ctx* ctx, const double* d_xyCell, int64_t ncell, int64_t m, int64_t chunk, double* d_xy, int32_t* d_did);
int design_p(const iak3d_design* g);
int design_ncov(const iak3d_design* g);
iak3d_ctx* design_ctx(const iak3d_design* g);

// ---- factorisation (chol.cu) -------------------------------------------------------------------
struct ProblemDesc {
  double* L; int64_t nRUL;       // factor storage (blocked, nRUL row units); B operand + diagonal blocks
  double* P; int64_t nRUP;       // row-panel storage the task owns (== L when factorising), blocked
  double* invd;                  // nCB x INVD_BLOCK inverse diagonal blocks
  int32_t* doneL; int32_t* doneP; int32_t* invdone;
  int32_t nCB, nRB;              // col-blocks of L / row-blocks of P
  int32_t mode;                  // 0 = factorise (P == L), 1 = bordered rows against a finished factor,
                                 // 2 = plain product  P(r,j) = use_cin * P(r,j) + alpha * sum_k A(r,k) L(j,k)'  (no flags)
  const double* A; int64_t nRUA; // mode 2: left operand (blocked); modes 0/1 read their own panel P
  int32_t kchunks;               // mode 2: K / 32
  int32_t use_cin;               // mode 2: accumulate onto the existing tile
  double alpha;                  // mode 2: scale of the product
  int32_t epoch;
  int32_t w0, q;                 // W rows start (global row index in L) and count
  int32_t nlive;                 // mode 0: rows [nlive, w0) are identity padding and rows >= w0 + q are empty (0: unknown, all rows live)
  double* G;                     // mode 0: q*q Gram; mode 1: rowsP x (q+1) accumulators
  int64_t ldG;
  double* logsum;                // nCB
  int32_t* status;               // [0] not-SPD, [1] timeout
  long long* prof;               // optional per-task timestamps (8 per task, globaltimer ns), nullptr = off
};
int chol_configure(iak3d_ctx* ctx);
int launch_tile_tasks(iak3d_ctx* ctx, const ProblemDesc* d_problems, const int4* d_tasks, int32_t ntasks, int32_t* d_ticket /* 2 ints */,
                      int mode /* the mode of every problem of the launch */,
                      int latency_bound = 0 /* one factorisation alone: never the CTA-pair engine (the column-to-column path is what counts) */,
                      int width = 0 /* independent tasks per dependency step (column of a factorisation, row chains of a panel,
                                       all tiles of a product): selects the CTA-pair engine for wide launches */);
int launch_finish(iak3d_ctx* ctx, const ProblemDesc* d_problems, int32_t count, double* d_results, int32_t stride);
// X = L^-T Y for the q bordered rows of a finished blocked factor; d_X: Npad x q column-major out
int launch_backsolve(iak3d_ctx* ctx, const double* d_L, int64_t nRU, int64_t Npad, const double* d_invd, int32_t q, double* d_X);
int launch_peak(iak3d_ctx* ctx, int kind, double* tflops);
// kriging epilogue (predictIAK3D.R:229-255): z, v from the bordered-row accumulators G (ldG x (q+1))
int launch_krige_finish(iak3d_ctx* ctx, const double* d_G, int64_t ldG, int32_t q, const double* d_XMap, int64_t ldX, int64_t m,
                        const double* d_Ckk, const double* d_beta, const double* d_vbeta, double* d_z, double* d_v,
                        double* d_cz = nullptr, double* d_qf = nullptr, double* d_cx = nullptr, int64_t ldcx = 0);
