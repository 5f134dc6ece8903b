// chol.cu -- the factorisation engine: a persistent, ticket-scheduled, left-looking tile Cholesky
// whose update GEMMs run on the FP64 tensor pipe (DMMA, mma.sync.m8n8k4.f64) and whose operands are
// staged by the TMA engine (cp.async.bulk + mbarrier), one CTA per SM.
//
// It replaces  chol() + forwardsolve()  of lndetANDinvCb (optim_functions.R:292-302) and the
// t(W) %*% iAW product of nllIAK3D (fitIAK3D.R:810), and -- run on extra "bordered" row panels --
// the  Ckh %*% iC  products of predictIAK3D (predictIAK3D.R:202-208).
//
// Data layout.  The factor buffer is BLOCKED (common.cuh: 64 x 32 units with the padded stride 68, ordered column
// panel by column panel) and has ld rows (multiple of TM):
//   rows [0,n) x cols [0,n)            lower triangle of A, overwritten by L
//   rows/cols [n,Npad)                 identity padding (Npad = n rounded up to TN)
//   rows [Npad,Npad+q) x cols [0,Npad) W' (q = p+1 rows), overwritten by Y' = W' L^-T
// so that sum_j Y_j' Y_j = W' A^-1 W comes out of the same sweep ("bordered" Cholesky).
//
// Tiles are TM x TN = 128 x 64.  Task (r, j) owns tile rows [128r,128r+128) x cols [64j,64j+64):
//   acc = A_tile - sum_{k<j} L(r,k) L(jrows,k)'          K-loop, DMMA; a 32-deep stage is TWO bulk copies
//                                                         (two stacked A units, one B unit)
//   diagonal tile:  potf2 of the 64x64 block fused with its inverse (registers + shared memory)
//   all other rows: X = acc * inv(L_jj)'                  DMMA
// Dependencies are per-tile epoch flags in global memory (release/acquire); tasks are claimed from
// an atomic ticket in a topological order, so a waiting CTA only ever waits for a CTA
// that already holds an earlier ticket: no deadlock, no co-residency requirement.
#include <algorithm>

#include "common.cuh"

namespace {

constexpr int KC = UC;                 // k-depth of one pipeline stage == unit width (32)
constexpr int STAGES = 3;
constexpr int LDB_S = UL;              // 68: conflict-free DMMA fragment loads (stride = 4 mod 16)
constexpr int A_STAGE = 2 * UNIT;      // two vertically stacked units: 128 rows x 32 k
constexpr int B_STAGE = UNIT;          // one unit: 64 rows x 32 k
constexpr int STAGE_D = A_STAGE + B_STAGE;          // doubles per stage (52,224 B)
constexpr int AT_D = 4 * UNIT;                      // the task's own tile: [col half][row half] units (69,632 B)
constexpr int SMEM_BYTES = (STAGES * STAGE_D + AT_D) * 8;   // 226,304 B
// tile element (row 0..127, col 0..63) inside At
#define AT_IDX(row, col) ((((col) >> 5) * 2 + ((row) >> 6)) * UNIT + ((col) & 31) * UL + ((row) & 63))
constexpr int NCW = 8;                 // consumer (MMA) warps: 4 (rows) x 2 (cols), 32x32 each
constexpr int NTHREADS = (NCW + 1) * 32;
constexpr unsigned long long WAIT_TIMEOUT_NS = 4000000000ull;

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t;
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
               : "=r"(ok) : "r"(s_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// bounded wait: a protocol error turns into status[1] = 1 instead of a hung GPU
__device__ __forceinline__ int ld_volatile(const int32_t* p) {
  int v; asm volatile("ld.volatile.global.b32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int32_t* status, int32_t* abort_flag) {
  if (mbar_try(bar, parity)) return;
  const unsigned long long t0 = gtime();
  uint32_t spins = 0;
  while (!mbar_try(bar, parity)) {
    if ((++spins & 1023u) == 0) {
      if (ld_volatile(abort_flag) != 0) return;
      if (gtime() - t0 > WAIT_TIMEOUT_NS) { atomicExch(status + 1, 1); atomicExch(abort_flag, 1); return; }
    }
  }
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s_u32(dst)),
               "l"(src), "r"(bytes), "r"(s_u32(bar)) : "memory");
}
__device__ __forceinline__ int ld_acquire(const int32_t* p) {
  int v; asm volatile("ld.acquire.gpu.global.b32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void st_release(int32_t* p, int v) {
  asm volatile("st.release.gpu.global.b32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void wait_flag(const int32_t* f, int epoch, int32_t* status, int32_t* abort_flag) {
  if (ld_acquire(f) == epoch) return;
  const unsigned long long t0 = gtime();
  uint32_t spins = 0;
  while (ld_acquire(f) != epoch) {
    __nanosleep(40);
    if ((++spins & 63u) == 0) {
      if (ld_volatile(abort_flag) != 0) return;
      if (gtime() - t0 > WAIT_TIMEOUT_NS) { atomicExch(status + 1, 1); atomicExch(abort_flag, 1); return; }
    }
  }
}
__device__ __forceinline__ void consumer_bar() { asm volatile("bar.sync 1, %0;" ::"n"(NCW * 32) : "memory"); }

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// one k4 step of a warp's 32x32 sub-tile.  A element (row, k) at Ab[k*UL + row-part] with the row-part folded into
// aoff (unit of the row half + row within the unit); B element (col, k) at Bb[k*UL + col].
__device__ __forceinline__ void warp_mma_k4(double (&acc)[4][4][2], const double* __restrict__ Ab, int aoff,
                                            const double* __restrict__ Bb, int boff, int k0, int lane) {
  const int kq = lane & 3;
  double a[4], b[4];
#pragma unroll
  for (int mi = 0; mi < 4; mi++) a[mi] = Ab[(k0 + kq) * UL + aoff + mi * 8];
#pragma unroll
  for (int ni = 0; ni < 4; ni++) b[ni] = Bb[(k0 + kq) * UL + boff + ni * 8];
#pragma unroll
  for (int mi = 0; mi < 4; mi++)
#pragma unroll
    for (int ni = 0; ni < 4; ni++) dmma(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
}

// E3 helper: k-steps [kbeg, kend) of X = tile * inv(L_jj)' for the column tiles NLO..3 of this warp (tcol: their 8-column
// tile indices).  A element (row, k) at Ab0[(k>>5)*2*UNIT + (k&31)*UL + mi*8], B element (col, k) at Bb0[k*LDB_S + 8*tile].
template <int NLO>
__device__ __forceinline__ void solve_segment(double (&x)[4][4][2], const double* __restrict__ Ab0, const double* __restrict__ Bb0,
                                              const int (&tcol)[4], int kbeg, int kend) {
#pragma unroll 2
  for (int k0 = kbeg; k0 < kend; k0 += 4) {
    const double* Ab = Ab0 + (k0 >> 5) * 2 * UNIT + (k0 & 31) * UL;
    const double* Bb = Bb0 + k0 * LDB_S;
    double a[4], b[4];
#pragma unroll
    for (int mi = 0; mi < 4; mi++) a[mi] = Ab[mi * 8];
#pragma unroll
    for (int ni = NLO; ni < 4; ni++) b[ni] = Bb[tcol[ni] * 8];
#pragma unroll
    for (int ni = NLO; ni < 4; ni++)
#pragma unroll
      for (int mi = 0; mi < 4; mi++) dmma(x[mi][ni][0], x[mi][ni][1], a[mi], b[ni]);
  }
}

// 1/sqrt(x) for a positive finite pivot: MUFU seed (relative error 9e-7) + ONE third-order step,
//   y (1 + e/2 + 3 e^2/8),  e = 1 - x y^2   (remainder 5/16 e^3 ~ 2e-18),
// max relative error 1.23 ulp over 2^20 inputs (scripts/micro/rsqrt_acc.cu; two Newton steps: 1.24 ulp) in four dependent
// operations instead of six: the result sits on the column-to-column dependency chain of every diagonal block.  rsqrt() from
// the math library carries special-case handling on top.
__device__ __forceinline__ double fast_rsqrt(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double t = x * y;
  const double e = fma(-t, y, 1.0);
  const double w = fma(e, 0.375, 0.5);
  const double ye = y * e;
  return fma(ye, w, y);
}

// ---- diagonal block, 4 columns per step ------------------------------------------------------------------------------
// potf2 fused with the inverse of the factor.  Thread (br, bc) of a 16 x 16 grid owns the 4 x 4 blocks d of A/L and y of
// the inverse; a step is four columns: the owner of the diagonal 4 x 4 block factorises it in registers (the only serial rsqrt chain), the
// panel below it and the block row of the inverse left of it are finished by substitution against that 4 x 4 factor,
// and every other block takes one rank-4 update.  Two block barriers per four columns (the column-by-column version of
// the first half of the round needed four and an exchange per column: 53.8k cycles per block against 40.0k).  scratch: 2 x BLK4_BUF doubles (double-buffered by step parity).
constexpr int BLK4_L = 0;                 // l10 l20 l30 l21 l31 l32 r0 r1 r2 r3 (16 reserved)
constexpr int BLK4_P = 16;                // panel: [64 rows][4]
constexpr int BLK4_Y = 16 + 256;          // inverse block row: [4][64]
constexpr int BLK4_BUF = 16 + 256 + 256;

__device__ __forceinline__ void blk4_subst_row(const double (&l)[6], const double (&r)[4], double& v0, double& v1, double& v2, double& v3) {
  // (v0 v1 v2 v3) := (v0 v1 v2 v3) inv(L44)'  <=>  forward substitution along the four entries
  v0 = v0 * r[0];
  v1 = (v1 - v0 * l[0]) * r[1];
  v2 = (v2 - v0 * l[1] - v1 * l[3]) * r[2];
  v3 = (v3 - v0 * l[2] - v1 * l[4] - v2 * l[5]) * r[3];
}

// ALIAS: scratch overlaps Ys (the pair engine has no spare stage): one more barrier before the inverse is written there
template <bool ALIAS = false>
__device__ void potf2_inv_blk4(double* __restrict__ At, int off, double* __restrict__ Ys, double* __restrict__ scratch,
                               double* __restrict__ pivs /* 64 */, int tid, double* logsum_out, int32_t* status) {
  const int br = tid >> 4, bc = tid & 15;
  const bool live = (bc <= br);
  double d[4][4], y[4][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int jj = 0; jj < 4; jj++) {
      d[i][jj] = live ? At[AT_IDX(off + 4 * br + i, 4 * bc + jj)] : 0.0;
      y[i][jj] = (br == bc && i == jj) ? 1.0 : 0.0;
    }
  bool bad = false;
#pragma unroll 1
  for (int cb = 0; cb < TN / 4; cb++) {
    double* buf = scratch + (cb & 1) * BLK4_BUF;
    if (br == cb && bc == cb) {
      // ---- 4 x 4 factor in registers: l = (l10 l20 l30 l21 l31 l32), r = 1 / diag ----
      double l[6], r[4], pv[4];
      pv[0] = d[0][0]; r[0] = fast_rsqrt(pv[0]);
      l[0] = d[1][0] * r[0]; l[1] = d[2][0] * r[0]; l[2] = d[3][0] * r[0];
      pv[1] = d[1][1] - l[0] * l[0]; r[1] = fast_rsqrt(pv[1]);
      l[3] = (d[2][1] - l[1] * l[0]) * r[1]; l[4] = (d[3][1] - l[2] * l[0]) * r[1];
      pv[2] = d[2][2] - l[1] * l[1] - l[3] * l[3]; r[2] = fast_rsqrt(pv[2]);
      l[5] = (d[3][2] - l[2] * l[1] - l[4] * l[3]) * r[2];
      pv[3] = d[3][3] - l[2] * l[2] - l[4] * l[4] - l[5] * l[5]; r[3] = fast_rsqrt(pv[3]);
#pragma unroll
      for (int k = 0; k < 4; k++) { if (!(pv[k] > 0.0) || isinf(pv[k])) bad = true; pivs[4 * cb + k] = pv[k]; }
#pragma unroll
      for (int k = 0; k < 6; k++) buf[BLK4_L + k] = l[k];
#pragma unroll
      for (int k = 0; k < 4; k++) buf[BLK4_L + 6 + k] = r[k];
      // L44 replaces d; the diagonal block of the inverse is inv(L44): substitution applied to the identity's columns
      d[0][0] = pv[0] * r[0]; d[0][1] = 0.0; d[0][2] = 0.0; d[0][3] = 0.0;
      d[1][0] = l[0]; d[1][1] = pv[1] * r[1]; d[1][2] = 0.0; d[1][3] = 0.0;
      d[2][0] = l[1]; d[2][1] = l[3]; d[2][2] = pv[2] * r[2]; d[2][3] = 0.0;
      d[3][0] = l[2]; d[3][1] = l[4]; d[3][2] = l[5]; d[3][3] = pv[3] * r[3];
#pragma unroll
      for (int jj = 0; jj < 4; jj++) {
        blk4_subst_row(l, r, y[0][jj], y[1][jj], y[2][jj], y[3][jj]);
#pragma unroll
        for (int k = 0; k < 4; k++) buf[BLK4_Y + k * 64 + 4 * bc + jj] = y[k][jj];
      }
    }
    consumer_bar();
    if (live && (bc == cb) != (br == cb)) {
      double l[6], r[4];
#pragma unroll
      for (int k = 0; k < 6; k++) l[k] = buf[BLK4_L + k];
#pragma unroll
      for (int k = 0; k < 4; k++) r[k] = buf[BLK4_L + 6 + k];
      if (bc == cb) {       // panel block below the diagonal block: rows of d times inv(L44)'
#pragma unroll
        for (int i = 0; i < 4; i++) {
          blk4_subst_row(l, r, d[i][0], d[i][1], d[i][2], d[i][3]);
#pragma unroll
          for (int k = 0; k < 4; k++) buf[BLK4_P + (4 * br + i) * 4 + k] = d[i][k];
        }
      } else {              // block row cb of the inverse, left of the diagonal: inv(L44) times the columns of y
#pragma unroll
        for (int jj = 0; jj < 4; jj++) {
          blk4_subst_row(l, r, y[0][jj], y[1][jj], y[2][jj], y[3][jj]);
#pragma unroll
          for (int k = 0; k < 4; k++) buf[BLK4_Y + k * 64 + 4 * bc + jj] = y[k][jj];
        }
      }
    }
    consumer_bar();
    if (live && br > cb) {
      double pr[4][4];
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int k = 0; k < 4; k++) pr[i][k] = buf[BLK4_P + (4 * br + i) * 4 + k];
      if (bc > cb) {        // trailing block: d -= P_br P_bc'
        double pc[4][4];
#pragma unroll
        for (int jj = 0; jj < 4; jj++)
#pragma unroll
          for (int k = 0; k < 4; k++) pc[jj][k] = buf[BLK4_P + (4 * bc + jj) * 4 + k];
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
          for (int jj = 0; jj < 4; jj++)
#pragma unroll
            for (int k = 0; k < 4; k++) d[i][jj] -= pr[i][k] * pc[jj][k];
      } else {              // inverse block (br, bc <= cb): y -= P_br Y(cb, bc)
        double yb[4][4];
#pragma unroll
        for (int k = 0; k < 4; k++)
#pragma unroll
          for (int jj = 0; jj < 4; jj++) yb[k][jj] = buf[BLK4_Y + k * 64 + 4 * bc + jj];
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
          for (int jj = 0; jj < 4; jj++)
#pragma unroll
            for (int k = 0; k < 4; k++) y[i][jj] -= pr[i][k] * yb[k][jj];
      }
    }
  }
  // ---- write L (zeros above the diagonal) and the inverse back to shared memory ----
  if (ALIAS) consumer_bar();
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int jj = 0; jj < 4; jj++) {
      const int r = 4 * br + i, cc = 4 * bc + jj;
      At[AT_IDX(off + r, cc)] = (live && r >= cc) ? d[i][jj] : 0.0;
      Ys[cc * LDB_S + r] = (live && r >= cc) ? y[i][jj] : 0.0;
    }
  if (bad) atomicExch(status, 1);
  consumer_bar();
  if (tid < 32) {   // log-determinant of the block: fixed-order tree, deterministic
    double v = log(pivs[tid]) + log(pivs[tid + 32]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (tid == 0) *logsum_out = v;
  }
}

// ---- diagonal block, warp-specialised (every factorisation launch of the full-size CTA) ---------------------------------
// The per-column critical path of a single evaluation is this routine (23 of the 35 us between two diagonal blocks at
// n = 1200 ... 5000 with round 1's look-ahead variant), and there most cycles were not the pivot chain but what surrounded it:
// two block-wide barriers and three shared-memory round trips per four columns, and -- in every warp -- the panel,
// inverse-row and trailing-update branches executed one after the other (divergence).  Here:
//   two CHAIN warps (0, 1): thread t owns row t of the current four-column panel.  Per step every lane factorises the 4 x 4
//     diagonal block for itself (no broadcast), substitutes its own row (rows of the diagonal block take the same path: the
//     entries right of the diagonal are replaced by zeros with selects, never with branches), publishes it transposed for
//     the helpers, then applies this step's rank-4 update to the NEXT panel column itself -- whose earlier updates the helpers
//     deposited a step ago -- and to the next diagonal block (again in every lane).  The first warp's rows are finished after
//     step 7: it retires and only keeps the barrier counts.
//   HELPER warps (2, 3, 6, 7 and 4): thread <-> one 4 x 4 block (br, bc <= br) held in registers: the trailing block T(br, bc)
//     until its column is handed to the chain warps, the block Y(br, bc) of the inverse afterwards.  One update per step and
//     thread, one instruction stream:  u -= P_br * Q,  Q = P_bc' of this step (trailing block) or the finished row block of
//     the inverse of the PREVIOUS step (the lag removes a helper-only barrier).  The long-lived blocks sit on the two
//     schedulers the chain warps do not use.  Warp 5 copies each finished panel into the tile.
// Hand-overs are bar.arrive / bar.sync pairs on named barriers (ids 3..6 alternating by step parity, 8 between the two chain
// warps); the per-step buffers cycle over three steps.  Measured alternatives and what each part costs: DESIGN.md section 4,
// profiles/potf2_variants_r02.txt (scripts/micro/potf2_bench.cu).
constexpr int CH_LR = 0;                  // l10 l20 l30 l21 l31 l32 | r0..r3   (16 reserved)
constexpr int CH_PT = 16;                 // this step's panel, transposed: [4][64] (conflict-free for the helpers' block loads)
constexpr int CH_YR = 16 + 256;           // finished row block of the inverse: [4][64]
constexpr int CH_BUF = 16 + 512;          // one of THREE such buffers per step (s mod 3): the helpers still read the previous
                                          // step's while the chain warps write the next step's
constexpr int CH_NB = 3 * CH_BUF;         // + parity * 256: next panel column as deposited by the helpers: [64 rows][4]
constexpr int CH_CHAIN = 2;               // chain warps
constexpr int CH_HELPERS = (NCW - CH_CHAIN) * 32;

template <int ID, int COUNT> __device__ __forceinline__ void nbar_sync() { asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(COUNT) : "memory"); }
template <int ID, int COUNT> __device__ __forceinline__ void nbar_arrive() { asm volatile("bar.arrive %0, %1;" ::"n"(ID), "n"(COUNT) : "memory"); }

__device__ __forceinline__ void ld4(const double* __restrict__ p, double (&v)[4]) {       // 32-byte aligned
  const double2 a = reinterpret_cast<const double2*>(p)[0], b = reinterpret_cast<const double2*>(p)[1];
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
__device__ __forceinline__ void st4(double* __restrict__ p, const double (&v)[4]) {
  reinterpret_cast<double2*>(p)[0] = make_double2(v[0], v[1]);
  reinterpret_cast<double2*>(p)[1] = make_double2(v[2], v[3]);
}

#ifdef POTF2_PROF
__device__ long long g_potf2_prof[16];
// BAR.SYNC is "defer blocking": the warp runs on until its next memory instruction.  A probe after a barrier first reads shared
// memory and branches on the value, so that the wait is charged to the barrier and not to whatever follows.
#define PROF_T(k) do { const long long t_ = clock64(); pacc_[k] += t_ - tp_; tp_ = t_; } while (0)
#define PROF_TB(k) do { unsigned x_; asm volatile("ld.volatile.shared.b32 %0, [%1];" : "=r"(x_) : "r"(s_u32(scratch)) : "memory"); \
                        if (x_ == 0x7fc12345u) pacc_[k] += 1; PROF_T(k); } while (0)
#else
#define PROF_T(k) do { } while (0)
#define PROF_TB(k) do { } while (0)
#endif

__device__ __forceinline__ double zero_if_less(double v, int a, int b) {     // a < b ? 0 : v, without a branch
  double o;
  asm volatile("{\n .reg .pred p;\n setp.lt.s32 p, %2, %3;\n selp.f64 %0, 0d0000000000000000, %1, p;\n}\n" : "=d"(o) : "d"(v), "r"(a), "r"(b));
  return o;
}

// DBG (micro-benchmark only, wrong results): 1 no panel stores, 2 no l / r / pivot publication, 4 no update of the next diagonal
// block, 8 no update of the next panel row, 16 helpers skip their updates, 32 helpers skip the inverse rows
template <int DBG = 0>
__device__ void potf2_inv_chain(double* __restrict__ At, int off, double* __restrict__ Ys, double* __restrict__ scratch,
                                double* __restrict__ pivs /* 64 */, int tid, double* logsum_out, int32_t* status) {
  const int warp = tid >> 5, lane = tid & 31;
  if (warp < CH_CHAIN) {
    // ================= chain warps =================
    // a lone warp issues an instruction every ~3.5 cycles whatever it is: what counts below is the number of instructions on
    // the path.  So the 64 panel rows are split over two warps (one row per lane, each warp factorising the diagonal block
    // for itself), addresses are running pointers, and the rows of the diagonal block take the same path as every other row.
    const int row = tid;                              // 0..63
    double c[4], dg[10];
    double* ptp = scratch + CH_PT + row;              // own row in the transposed panel buffer (parity 0)
#pragma unroll
    for (int k = 0; k < 4; k++) c[k] = At[AT_IDX(off + row, k)];
    dg[0] = At[AT_IDX(off + 0, 0)]; dg[1] = At[AT_IDX(off + 1, 0)]; dg[2] = At[AT_IDX(off + 2, 0)]; dg[3] = At[AT_IDX(off + 3, 0)];
    dg[4] = At[AT_IDX(off + 1, 1)]; dg[5] = At[AT_IDX(off + 2, 1)]; dg[6] = At[AT_IDX(off + 3, 1)];
    dg[7] = At[AT_IDX(off + 2, 2)]; dg[8] = At[AT_IDX(off + 3, 2)]; dg[9] = At[AT_IDX(off + 3, 3)];
    int rel = row;                                    // row - 4 cb
    int pofs = 0;                                     // this step's buffer (offset from scratch), cycling over three
#ifdef POTF2_PROF
    long long pacc_[5] = {0, 0, 0, 0, 0};
    long long tp_ = clock64();
#endif
#pragma unroll 1
    for (int cb = 0; cb < TN / 4; cb++) {
      // ---- 4 x 4 factor, the same in every lane (a non-positive pivot shows as NaN in pivs: checked at the end) ----
      double l[6], r[4], pv[4];
      pv[0] = dg[0]; r[0] = fast_rsqrt(pv[0]);
      l[0] = dg[1] * r[0]; l[1] = dg[2] * r[0]; l[2] = dg[3] * r[0];
      pv[1] = dg[4] - l[0] * l[0]; r[1] = fast_rsqrt(pv[1]);
      l[3] = (dg[5] - l[1] * l[0]) * r[1]; l[4] = (dg[6] - l[2] * l[0]) * r[1];
      pv[2] = dg[7] - l[1] * l[1] - l[3] * l[3]; r[2] = fast_rsqrt(pv[2]);
      l[5] = (dg[8] - l[2] * l[1] - l[4] * l[3]) * r[2];
      pv[3] = dg[9] - l[2] * l[2] - l[4] * l[4] - l[5] * l[5]; r[3] = fast_rsqrt(pv[3]);
      PROF_T(0);
      // ---- the next panel as the helpers left it (updates of all earlier steps applied) ----
      const bool more = cb + 1 < TN / 4;
      double nx[4], dn[4][4];
      if (more) {
        if (cb & 1) nbar_sync<6, NCW * 32>(); else nbar_sync<5, NCW * 32>();
        PROF_TB(1);
        const double* nb = scratch + CH_NB + (cb & 1) * 256;
        ld4(nb + row * 4, nx);
#pragma unroll
        for (int i = 0; i < 4; i++) ld4(nb + (4 * cb + 4 + i) * 4, dn[i]);
      }
      // ---- own panel row: (c0 c1 c2 c3) inv(L44)'.  A row of the diagonal block comes out as the row of L44 (the entries
      //      right of the diagonal were computed from what lies above the diagonal of A: replaced by zeros) ----
      blk4_subst_row(l, r, c[0], c[1], c[2], c[3]);
      c[1] = zero_if_less(c[1], rel, 1);              // selects, not branches: the lanes must not split into five paths
      c[2] = zero_if_less(c[2], rel, 2);
      c[3] = zero_if_less(c[3], rel, 3);
      if (rel >= 0 && !(DBG & 1)) { double* t = ptp + pofs; t[0] = c[0]; t[64] = c[1]; t[128] = c[2]; t[192] = c[3]; }   // (a helper warp copies it into the tile)
      rel -= 4;
      PROF_T(2);
      if (tid == 32 && !(DBG & 2)) {                  // the second warp's first lane: its rows die last
        double2* lr = reinterpret_cast<double2*>(scratch + pofs + CH_LR);
        lr[0] = make_double2(l[0], l[1]); lr[1] = make_double2(l[2], l[3]); lr[2] = make_double2(l[4], l[5]);
        lr[3] = make_double2(r[0], r[1]); lr[4] = make_double2(r[2], r[3]);
        st4(pivs + 4 * cb, pv);
      }
      // the rows of the next diagonal block belong to the first warp up to step 6, to the second from step 7 on: only the
      // second warp ever waits for the other one's stores
      if (cb < 7) { if (warp == 0) nbar_arrive<8, CH_CHAIN * 32>(); else nbar_sync<8, CH_CHAIN * 32>(); }
      else __syncwarp();
      if (cb & 1) nbar_arrive<4, NCW * 32>(); else nbar_arrive<3, NCW * 32>();
      PROF_T(3);
      if (warp == 0 && cb >= 7) {                     // the first warp's rows are finished: it only keeps the barrier counts
#pragma unroll 1
        for (int c2 = cb + 1; c2 < TN / 4; c2++) {
          if (c2 + 1 < TN / 4) { if (c2 & 1) nbar_sync<6, NCW * 32>(); else nbar_sync<5, NCW * 32>(); }
          if (c2 & 1) nbar_arrive<4, NCW * 32>(); else nbar_arrive<3, NCW * 32>();
        }
        break;
      }
      if (more) {
        // ---- this step's rank-4 update of the next panel (own row) and of the next diagonal block (every lane) ----
        double pd[4][4];                              // pd[k][i] = P[4 cb + 4 + i][k]
#pragma unroll
        for (int k = 0; k < 4; k++) ld4(scratch + pofs + CH_PT + k * 64 + 4 * cb + 4, pd[k]);
        if (!(DBG & 4)) {
#pragma unroll
          for (int i = 0; i < 4; i++)
#pragma unroll
            for (int jj = 0; jj <= i; jj++)
#pragma unroll
              for (int k = 0; k < 4; k++) dn[i][jj] -= pd[k][i] * pd[k][jj];
        }
        dg[0] = dn[0][0]; dg[1] = dn[1][0]; dg[2] = dn[2][0]; dg[3] = dn[3][0]; dg[4] = dn[1][1]; dg[5] = dn[2][1]; dg[6] = dn[3][1];
        dg[7] = dn[2][2]; dg[8] = dn[3][2]; dg[9] = dn[3][3];
        if (!(DBG & 8)) {
#pragma unroll
          for (int jj = 0; jj < 4; jj++)
#pragma unroll
            for (int k = 0; k < 4; k++) nx[jj] -= c[k] * pd[k][jj];
        }
#pragma unroll
        for (int jj = 0; jj < 4; jj++) c[jj] = nx[jj];
      }
      pofs = (pofs == 2 * CH_BUF) ? 0 : pofs + CH_BUF;
      PROF_T(4);
    }
#ifdef POTF2_PROF
    if (tid == 32) for (int k = 0; k < 5; k++) g_potf2_prof[k] += pacc_[k];
#endif
  } else {
    // ================= helper warps =================
    // the chain warps sit on the first two of the SM's four schedulers (warp id mod 4) and share their FP64 pipes with warps 4
    // and 5: the long-lived blocks (late rows) go to the helper warps of the other two schedulers (2, 3, 6, 7), the first
    // rows -- finished after three steps -- to warp 4; warp 5 holds no block (it copies the finished panels into the tile)
    const int hw = warp == 2 ? 0 : (warp == 3 ? 1 : (warp == 6 ? 2 : (warp == 7 ? 3 : (warp == 4 ? 4 : 5))));
    const int b = 135 - (hw * 32 + lane);             // block number, row-major over the lower triangle; < 0: no block
    const bool live = b >= 0;
    int br = live ? (int)((sqrtf(8.0f * (float)b + 1.0f) - 1.0f) * 0.5f) : TN / 4;
    if (live) {
      while ((br + 1) * (br + 2) / 2 <= b) br++;
      while (br * (br + 1) / 2 > b) br--;
    }
    const int bc = live ? b - br * (br + 1) / 2 : 0;
    // zeros above the block diagonal of L and of the inverse: 120 blocks, spread over the helper threads
    for (int z = tid - CH_CHAIN * 32; z < 120; z += CH_HELPERS) {
      int zc = (int)((sqrtf(8.0f * (float)z + 1.0f) - 1.0f) * 0.5f);
      while ((zc + 1) * (zc + 2) / 2 <= z) zc++;
      while (zc * (zc + 1) / 2 > z) zc--;
      const int zr = z - zc * (zc + 1) / 2;          // block (zr, zc + 1), zr <= zc: strictly above the diagonal
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int jj = 0; jj < 4; jj++) { At[AT_IDX(off + 4 * zr + i, 4 * (zc + 1) + jj)] = 0.0; Ys[(4 * (zc + 1) + jj) * LDB_S + 4 * zr + i] = 0.0; }
    }
    double u[4][4];
    const bool tmode = live && bc >= 1;               // u holds the trailing block first, the block of the inverse later
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int jj = 0; jj < 4; jj++) u[i][jj] = tmode ? At[AT_IDX(off + 4 * br + i, 4 * bc + jj)] : ((br == bc && i == jj) ? 1.0 : 0.0);
    if (live && bc == 1) {                            // column block 1 goes to the chain warps untouched
#pragma unroll
      for (int i = 0; i < 4; i++) st4(scratch + CH_NB + (4 * br + i) * 4, u[i]);
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int jj = 0; jj < 4; jj++) u[i][jj] = (br == bc && i == jj) ? 1.0 : 0.0;
    }
    nbar_arrive<5, NCW * 32>();                       // column 1 is there
    // One update per thread and step, one instruction stream for all of them:  u -= P_br * Q.  A trailing block takes the
    // update of THIS step (Q = P_bc', from the panel just published); a block of the inverse takes the update of the
    // PREVIOUS step (Q = row block s - 1 of the inverse, finished at the end of the previous step) -- the lag removes the
    // helpers' own barrier between "finish the row block" and "use it": the next step's barrier with the chain warps orders
    // them.  The column the chain takes over next (bc == s + 2) is handed over right after the update.
    const int po_br = CH_PT + 4 * br, po_bc = CH_PT + 4 * bc, yo_bc = CH_YR + 4 * bc;
#ifdef POTF2_PROF
    long long pacc_[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long tp_ = clock64();
#endif
    const int crow = lane;                            // warp 5: rows crow and crow + 32 of every finished panel go into the tile
#pragma unroll 1
    for (int s = 0; s < TN / 4; s++) {
      double* buf = scratch + (s % 3) * CH_BUF;
      double* bufp = scratch + ((s + 2) % 3) * CH_BUF;
      if (s & 1) nbar_sync<4, NCW * 32>(); else nbar_sync<3, NCW * 32>();
      PROF_TB(5);
      const bool tupd = bc >= s + 2;                  // live implies br >= bc; column s + 1 is with the chain warps
      const bool yupd = br >= s && bc <= s - 1;
      if (live && (tupd || yupd) && !(DBG & 16)) {
        const double* pb_ = (tupd ? buf : bufp) + po_br;
        const double* qb_ = tupd ? buf + po_bc : bufp + yo_bc;
        double pr[4][4], qq[4][4];                    // pr[k][i] = P[4 br + i][k];  qq[k][jj] = P[4 bc + jj][k] or Y[4 s - 4 + k][4 bc + jj]
#pragma unroll
        for (int k = 0; k < 4; k++) { ld4(pb_ + k * 64, pr[k]); ld4(qb_ + k * 64, qq[k]); }
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
          for (int jj = 0; jj < 4; jj++)
#pragma unroll
            for (int k = 0; k < 4; k++) u[i][jj] -= pr[k][i] * qq[k][jj];
        if (bc == s + 2) {                            // all earlier updates applied: the chain warps take the column over
          double* nb = scratch + CH_NB + ((s + 1) & 1) * 256;
#pragma unroll
          for (int i = 0; i < 4; i++) st4(nb + (4 * br + i) * 4, u[i]);
#pragma unroll
          for (int i = 0; i < 4; i++)
#pragma unroll
            for (int jj = 0; jj < 4; jj++) u[i][jj] = (br == bc && i == jj) ? 1.0 : 0.0;      // from here on: the block of the inverse
        }
      }
      PROF_T(6);
      if (s + 2 < TN / 4) { if ((s + 1) & 1) nbar_arrive<6, NCW * 32>(); else nbar_arrive<5, NCW * 32>(); }
      PROF_T(7);
      if (warp == 5 && !(DBG & 1)) {                  // the finished panel (rows of the diagonal block included) into the tile
#pragma unroll
        for (int q = 0; q < 2; q++) {
          const int row = crow + 32 * q;
          if (row >= 4 * s) {
#pragma unroll
            for (int k = 0; k < 4; k++) At[AT_IDX(off + row, 4 * s + k)] = buf[CH_PT + k * 64 + row];
          }
        }
      }
      // ---- row block s of the inverse: inv(L44) times the columns of u; final ----
      if (live && br == s && !(DBG & 32)) {
        double l[6], r[4];
        const double2* lr = reinterpret_cast<const double2*>(buf + CH_LR);
        const double2 t0 = lr[0], t1 = lr[1], t2 = lr[2], t3 = lr[3], t4 = lr[4];
        l[0] = t0.x; l[1] = t0.y; l[2] = t1.x; l[3] = t1.y; l[4] = t2.x; l[5] = t2.y; r[0] = t3.x; r[1] = t3.y; r[2] = t4.x; r[3] = t4.y;
#pragma unroll
        for (int jj = 0; jj < 4; jj++) blk4_subst_row(l, r, u[0][jj], u[1][jj], u[2][jj], u[3][jj]);
#pragma unroll
        for (int k = 0; k < 4; k++) st4(buf + yo_bc + k * 64, u[k]);
#pragma unroll
        for (int jj = 0; jj < 4; jj++) {
          const double col[4] = {u[0][jj], u[1][jj], u[2][jj], u[3][jj]};
          st4(Ys + (4 * bc + jj) * LDB_S + 4 * br, col);
        }
      }
      PROF_T(8);
    }
#ifdef POTF2_PROF
    if (tid == POTF2_PROF_HELPER) for (int k = 5; k < 9; k++) g_potf2_prof[k] += pacc_[k];
#endif
  }
  consumer_bar();
  if (tid < 64) { const double pv = pivs[tid]; if (!(pv > 0.0) || isinf(pv)) atomicExch(status, 1); }
  if (tid < 32) {   // log-determinant of the block: fixed-order tree, deterministic
    double v = log(pivs[tid]) + log(pivs[tid + 32]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (tid == 0) *logsum_out = v;
  }
}

// MODE: every problem of a launch has the same mode (0 factorisation, 1 panel solve against a finished factor, 2 plain
// product).  Modes 1 and 2 have their own instantiation (no registers spent on the other epilogues: +1.2 % on the kriging
// panels); the factorisation runs the generic body (MODE = -1, mode read from the problem) -- its specialised
// instantiation spills 32 bytes and loses 1.3 %.
template <int MODE, int LOOKAHEAD>
__global__ void __launch_bounds__(NTHREADS, 1)
tile_task_kernel(const ProblemDesc* __restrict__ probs, const int4* __restrict__ tasks, int ntasks, int32_t* ticket) {
  extern __shared__ __align__(128) double smem[];
  double* At = smem + STAGES * STAGE_D;
  double* Ys = smem;                           // stage 0 (A part, 2 units) is free once the K-loop has drained
  double* Ws = smem + STAGE_D;                 // stage 1: W rows of the factor for the kriging epilogue
  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ __align__(8) uint64_t empty_bar[STAGES];
  __shared__ __align__(8) uint64_t at_bar;
  __shared__ __align__(8) uint64_t inv_bar;
  __shared__ __align__(16) double s_comm[4 * 64 + 64];
  __shared__ int s_task;
  __shared__ long long s_tdep;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < STAGES; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], NCW); }
    mbar_init(&at_bar, 1);
    mbar_init(&inv_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  int32_t* const abort_flag = ticket + 1;   // set by the first CTA whose bounded wait expired
  uint32_t it = 0;          // pipeline chunk counter, identical in every thread
  uint32_t at_phase = 0, inv_phase = 0;

  while (true) {
    if (tid == 0) s_task = (ld_volatile(abort_flag) != 0) ? ntasks : atomicAdd(ticket, 1);
    __syncthreads();
    const int t = s_task;
    if (t >= ntasks) break;
    const int4 tk = tasks[t];
    const ProblemDesc* pb = probs + tk.x;
    const int r = tk.y, j = tk.z;
    double* __restrict__ L = pb->L; const int64_t nRUL = pb->nRUL;
    double* __restrict__ P = pb->P; const int64_t nRUP = pb->nRUP;
    const int nCB = pb->nCB, epoch = pb->epoch;
    const int mode = (MODE >= 0) ? MODE : pb->mode;
    int32_t* status = pb->status;
    const int64_t R0 = (int64_t)r * TM;
    const int nchunks = (mode == 2) ? pb->kchunks : j * (TN / KC);
    // tasks of structurally triangular operands (the explicit inverse, gemm.cu) start their K-loop at chunk tk.w: the
    // operand is zero before it.  0 for every factorisation / kriging task.
    const int kstart = min(tk.w, nchunks);
    const double* __restrict__ Asrc = (mode == 2) ? pb->A : P;
    const int64_t nRUA = (mode == 2) ? pb->nRUA : nRUP;
    const uint32_t it0 = it;
    long long* prof = pb->prof ? pb->prof + (int64_t)t * 8 : nullptr;
    long long tp0 = 0, tp1 = 0, tp2 = 0, tp3 = 0, tp4 = 0, kwait = 0;
    if (prof && tid == 0) tp0 = (long long)gtime();
    // the tile's four units in the owner's buffer: column halves 2j, 2j+1; row units 2r, 2r+1 (adjacent in memory)
    double* const Pt0 = P + ((int64_t)(2 * j) * nRUP + 2 * r) * UNIT;
    double* const Pt1 = P + ((int64_t)(2 * j + 1) * nRUP + 2 * r) * UNIT;

    if (warp == NCW) {
      // ===================== producer: TMA bulk copies, issued by one lane =====================
      if (lane == 0) {
        // generic-proxy accesses of the previous task (epilogue reads/writes of At, Ys) were ordered by the
        // end-of-task __syncthreads; make them visible to the async proxy before it overwrites smem.
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(&at_bar, AT_D * 8);
        bulk_g2s(At, Pt0, 2 * UNIT * 8, &at_bar);
        bulk_g2s(At + 2 * UNIT, Pt1, 2 * UNIT * 8, &at_bar);
        uint32_t itp = it0;
        for (int c = kstart; c < nchunks; c++, itp++) {
          const int k = c >> 1;
          if ((c & 1) == 0 && mode != 2) {
            wait_flag(pb->doneP + (int64_t)r * nCB + k, epoch, status, abort_flag);
            if (mode == 0) wait_flag(pb->doneL + (int64_t)(j >> 1) * nCB + k, epoch, status, abort_flag);
            asm volatile("fence.proxy.async;" ::: "memory");
            if (prof && k == j - 1) s_tdep = (long long)gtime();
          }
          const uint32_t s = itp % STAGES, use = itp / STAGES;
          if (use > 0) mbar_wait(&empty_bar[s], (use - 1) & 1, status, abort_flag);
          mbar_expect_tx(&full_bar[s], STAGE_D * 8);
          double* As = smem + s * STAGE_D;
          // unit column c (32 columns of L): A = row units 2r, 2r+1 of the owner's panel; B = row unit j of the factor
          bulk_g2s(As, Asrc + ((int64_t)c * nRUA + 2 * r) * UNIT, A_STAGE * 8, &full_bar[s]);
          bulk_g2s(As + A_STAGE, L + ((int64_t)c * nRUL + j) * UNIT, B_STAGE * 8, &full_bar[s]);
        }
      }
      __syncwarp();
    } else {
      // ===================== consumers: DMMA K-loop + epilogue =====================
      const int wm = warp & 3, wn = warp >> 2;
      const int kq = lane & 3, rq = lane >> 2;
      const int aoff = (wm >> 1) * UNIT + (wm & 1) * 32 + rq;       // A fragment: unit of the row half + row in unit
      const int boff = wn * 32 + rq;                                 // B fragment: row of the 64-row unit
      double acc[4][4][2];
#pragma unroll
      for (int mi = 0; mi < 4; mi++)
#pragma unroll
        for (int ni = 0; ni < 4; ni++) { acc[mi][ni][0] = 0.0; acc[mi][ni][1] = 0.0; }
      uint32_t itc = it0;
      // diagonal tile of an odd column block: its upper 64 rows lie above the diagonal block and are never stored; the
      // four warps that own them leave the tensor pipe to the other four
      const bool dead_rows = (mode == 0) && (r == (j >> 1)) && ((j & 1) != 0) && (wm < 2);
      for (int c = kstart; c < nchunks; c++, itc++) {
        const uint32_t s = itc % STAGES;
        if (prof && tid == 0) {
          const long long c0 = clock64();
          mbar_wait(&full_bar[s], (itc / STAGES) & 1, status, abort_flag);
          kwait += clock64() - c0;
        } else {
          mbar_wait(&full_bar[s], (itc / STAGES) & 1, status, abort_flag);
        }
        const double* As = smem + s * STAGE_D;
        const double* Bs = As + A_STAGE;
        if (!dead_rows) {
#pragma unroll
          for (int k4 = 0; k4 < KC / 4; k4++) warp_mma_k4(acc, As, aoff, Bs, boff, k4 * 4, lane);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[s]);
      }
      // ---- E1: tile = A - acc (in shared memory) ----
      if (prof && tid == 0) tp1 = (long long)gtime();
      const bool is_diag = (mode == 0) && (r == (j >> 1));
      if (mode != 2 && !is_diag) {
        // the inverted diagonal block goes to stage 0 (Ys), free once every consumer has left the K-loop; fetched now so
        // that the copy runs under E1.  A finished factor (mode 1) needs no wait: its blocks were published long ago.
        consumer_bar();
        if (tid == 0) {
          if (mode == 0) wait_flag(pb->invdone + j, epoch, status, abort_flag);
          asm volatile("fence.proxy.async;" ::: "memory");
          const bool wfetch = (mode == 1) && (pb->q > 0);
          mbar_expect_tx(&inv_bar, (INVD_BLOCK + (wfetch ? 2 * UNIT : 0)) * 8);
          bulk_g2s(Ys, pb->invd + (int64_t)j * INVD_BLOCK, INVD_BLOCK * 8, &inv_bar);
          if (wfetch) {
            // kriging epilogue: the two units of the factor buffer that hold the W rows (w0 is a multiple of 64) under
            // this column block go to stage 1 as they lie -- a unit is already the B-operand layout [k*UL + row]
            const int64_t wu = (int64_t)pb->w0 / UR;
            bulk_g2s(Ws, L + ((int64_t)(2 * j) * nRUL + wu) * UNIT, UNIT * 8, &inv_bar);
            bulk_g2s(Ws + UNIT, L + ((int64_t)(2 * j + 1) * nRUL + wu) * UNIT, UNIT * 8, &inv_bar);
          }
        }
      }
      mbar_wait(&at_bar, at_phase, status, abort_flag);
      if (mode == 2) {
        // ---- plain product: tile = use_cin * tile + alpha * acc, stored straight back; no dependencies to publish ----
        const double alpha = pb->alpha; const bool cin = pb->use_cin != 0;
#pragma unroll
        for (int mi = 0; mi < 4; mi++)
#pragma unroll
          for (int ni = 0; ni < 4; ni++) {
            const int row = wm * 32 + mi * 8 + rq, col = wn * 32 + ni * 8 + kq * 2;
            const int i0 = AT_IDX(row, col), i1 = AT_IDX(row, col + 1);
            At[i0] = (cin ? At[i0] : 0.0) + alpha * acc[mi][ni][0];
            At[i1] = (cin ? At[i1] : 0.0) + alpha * acc[mi][ni][1];
          }
        consumer_bar();
        for (int idx = tid; idx < 4 * (UNIT / 2); idx += NCW * 32) {
          const int u = idx / (UNIT / 2), w2 = idx % (UNIT / 2);
          reinterpret_cast<double2*>(((u >> 1) ? Pt1 : Pt0) + (u & 1) * UNIT)[w2] = reinterpret_cast<const double2*>(At + u * UNIT)[w2];
        }
        it = it0 + (nchunks - kstart);
        at_phase ^= 1;
        __syncthreads();
        continue;
      }
#pragma unroll
      for (int mi = 0; mi < 4; mi++)
#pragma unroll
        for (int ni = 0; ni < 4; ni++) {
          const int row = wm * 32 + mi * 8 + rq, col = wn * 32 + ni * 8 + kq * 2;
          At[AT_IDX(row, col)] -= acc[mi][ni][0];
          At[AT_IDX(row, col + 1)] -= acc[mi][ni][1];
        }
      consumer_bar();
      if (prof && tid == 0) tp2 = (long long)gtime();
      // ---- E2: diagonal block or fetch of the inverted diagonal block ----
      int row_lo = 0;
      bool defer_inv = false;
      if (is_diag) {
        const int off = (j & 1) * TN;
        if (LOOKAHEAD) potf2_inv_chain(At, off, Ys, smem + 2 * STAGE_D, s_comm + 256, tid, pb->logsum + j, status);
        else potf2_inv_blk4<false>(At, off, Ys, smem + 2 * STAGE_D, s_comm + 256, tid, pb->logsum + j, status);   // scratch: stage 2, drained
        // Even columns: the next diagonal block waits for THIS task's lower half (solved below from the inverse in shared
        // memory) and the column's other tiles have a whole diagonal block of slack: the inverse goes out after the tile has
        // been published (end of the task).  Odd columns: the next diagonal block waits for another CTA's tile, which waits
        // for the inverse: it goes out first.
        defer_inv = LOOKAHEAD && off == 0;
        if (!defer_inv) {
          double* invd = pb->invd + (int64_t)j * INVD_BLOCK;     // stored with the padded stride: one bulk copy reloads it
          for (int idx = tid; idx < INVD_BLOCK / 2; idx += NCW * 32)
            reinterpret_cast<double2*>(invd)[idx] = reinterpret_cast<const double2*>(Ys)[idx];
          __threadfence();
          asm volatile("fence.proxy.async;" ::: "memory");   // later readers fetch this block with the TMA engine
          consumer_bar();
          if (tid == 0) st_release(pb->invdone + j, epoch);
        }
        // L_jj itself: the two units (column halves) of row half off/64.  Nobody reads it before the tile's own flag (E6, which
        // fences): it is stored AFTER the inverse has been published, off the column-to-column critical path
        for (int idx = tid; idx < UNIT; idx += NCW * 32) {        // UNIT double2 = 2 units
          const int ch = idx / (UNIT / 2), w2 = idx % (UNIT / 2);
          const double2 v = reinterpret_cast<const double2*>(At + (ch * 2 + (off >> 6)) * UNIT)[w2];
          reinterpret_cast<double2*>((ch ? Pt1 : Pt0) + (off >> 6) * UNIT)[w2] = v;
        }
        row_lo = off + TN;
      } else {
        mbar_wait(&inv_bar, inv_phase, status, abort_flag);
        inv_phase ^= 1;
      }
      if (prof && tid == 0) tp3 = (long long)gtime();
      if (row_lo < TM) {
        // ---- E3: X = tile * inv(L_jj)' on the tensor pipe (k runs over the tile's 64 columns: two column halves) ----
        // inv(L_jj)' is upper triangular: the 8-column tile t of X needs k < 8(t+1) only.  The eight column tiles are
        // dealt to the two column warps as {0,3,4,7} and {1,2,5,6}, so both run 36 of the 64 (tile, k4) steps.
        double x[4][4][2];
#pragma unroll
        for (int mi = 0; mi < 4; mi++)
#pragma unroll
          for (int ni = 0; ni < 4; ni++) { x[mi][ni][0] = 0.0; x[mi][ni][1] = 0.0; }
        int tcol[4];
#pragma unroll
        for (int ni = 0; ni < 4; ni++) tcol[ni] = 4 * (ni >> 1) + ((ni & 1) ? (3 - wn) : wn);
        // four segments of k with a fixed set of live tiles each (tiles are sorted by extent): no branch inside a step
        const double* Ab0 = At + kq * UL + aoff;
        const double* Bb0 = Ys + kq * LDB_S + rq;
        solve_segment<0>(x, Ab0, Bb0, tcol, 0, 8 * (tcol[0] + 1));
        solve_segment<1>(x, Ab0, Bb0, tcol, 8 * (tcol[0] + 1), 8 * (tcol[1] + 1));
        solve_segment<2>(x, Ab0, Bb0, tcol, 8 * (tcol[1] + 1), 8 * (tcol[2] + 1));
        solve_segment<3>(x, Ab0, Bb0, tcol, 8 * (tcol[2] + 1), 8 * (tcol[3] + 1));
        consumer_bar();
#pragma unroll
        for (int mi = 0; mi < 4; mi++)
#pragma unroll
          for (int ni = 0; ni < 4; ni++) {
            const int row = wm * 32 + mi * 8 + rq, col = tcol[ni] * 8 + kq * 2;
            if (row >= row_lo) { At[AT_IDX(row, col)] = x[mi][ni][0]; At[AT_IDX(row, col + 1)] = x[mi][ni][1]; }
          }
        consumer_bar();
        if (prof && tid == 0 && mode == 1) s_tdep = (long long)gtime();     // mode 1 has no dependency stamp: end of E3 instead
        // ---- E4: copy-out of row halves [row_lo/64, 2): whole units, coalesced 16-byte stores ----
        {
          const int rh0 = row_lo >> 6, nun = (2 - rh0) * 2;       // units to store
          for (int idx = tid; idx < nun * (UNIT / 2); idx += NCW * 32) {
            const int u = idx / (UNIT / 2), w2 = idx % (UNIT / 2);
            const int ch = u / (2 - rh0), rh = rh0 + u % (2 - rh0);
            const double2 v = reinterpret_cast<const double2*>(At + (ch * 2 + rh) * UNIT)[w2];
            reinterpret_cast<double2*>((ch ? Pt1 : Pt0) + rh * UNIT)[w2] = v;
          }
        }
        // ---- E5: border reductions (ordered by the flag chain => deterministic, no atomics) ----
        const int q = pb->q;
        if (mode == 0) {
          const int wl0 = pb->w0 - (int)R0;
          if (q > 0 && wl0 >= row_lo && wl0 < TM) {
            for (int e = tid; e < q * q; e += NCW * 32) {
              const int a = e % q, b = e / q;
              double s = 0.0;
              for (int c = 0; c < TN; c++) s += At[AT_IDX(wl0 + a, c)] * At[AT_IDX(wl0 + b, c)];
              double* G = pb->G + e;            // __ldcg: the previous column's CTA wrote it through L2
              *G = (j == 0 ? 0.0 : __ldcg(G)) + s;
            }
          }
        } else {
          // kriging rows: G(:,0..q-1) += X Yw', G(:,q) += rowsum(X^2); Yw = rows w0.. of the factor buffer.
          // On the tensor pipe: each warp owns 16 rows of the tile (two m8 tiles), the W rows are the B operand
          // (fetched with the inverse diagonal block: element (w, c) at Ws[(c>>5)*UNIT + (c&31)*UL + w]), 16 W rows per pass.
          {
            const int row0 = warp * 16 + rq;                       // + 8*mi
            double* const Gr = pb->G + R0 + row0;
            const int64_t ldG = pb->ldG;
            double ssq[2] = {0.0, 0.0};
            for (int wb = 0; wb < q; wb += 16) {
              double g[2][2][2], gold[2][2][2];
#pragma unroll
              for (int mi = 0; mi < 2; mi++)
#pragma unroll
                for (int ni = 0; ni < 2; ni++)
#pragma unroll
                  for (int e = 0; e < 2; e++) {
                    g[mi][ni][e] = 0.0;
                    const int col = wb + ni * 8 + kq * 2 + e;
                    gold[mi][ni][e] = (j == 0 || col >= q) ? 0.0 : __ldcg(Gr + mi * 8 + (int64_t)col * ldG);
                  }
#pragma unroll 4
              for (int k4 = 0; k4 < TN / 4; k4++) {
                const int k = k4 * 4 + kq;
                const double a0 = At[AT_IDX(row0, k)], a1 = At[AT_IDX(row0 + 8, k)];
                const double* wk = Ws + (k >> 5) * UNIT + (k & 31) * UL + wb + rq;
                const double b0 = wk[0], b1 = wk[8];
                if (wb == 0) { ssq[0] = fma(a0, a0, ssq[0]); ssq[1] = fma(a1, a1, ssq[1]); }
                dmma(g[0][0][0], g[0][0][1], a0, b0);
                dmma(g[0][1][0], g[0][1][1], a0, b1);
                dmma(g[1][0][0], g[1][0][1], a1, b0);
                dmma(g[1][1][0], g[1][1][1], a1, b1);
              }
#pragma unroll
              for (int mi = 0; mi < 2; mi++)
#pragma unroll
                for (int ni = 0; ni < 2; ni++)
#pragma unroll
                  for (int e = 0; e < 2; e++) {
                    const int col = wb + ni * 8 + kq * 2 + e;
                    if (col < q) Gr[mi * 8 + (int64_t)col * ldG] = gold[mi][ni][e] + g[mi][ni][e];
                  }
            }
#pragma unroll
            for (int mi = 0; mi < 2; mi++) {                       // the four k-lanes of a row: fixed-order butterfly
              ssq[mi] += __shfl_xor_sync(0xffffffffu, ssq[mi], 1);
              ssq[mi] += __shfl_xor_sync(0xffffffffu, ssq[mi], 2);
            }
            if (kq == 0) {
#pragma unroll
              for (int mi = 0; mi < 2; mi++) {
                double* Gq = Gr + mi * 8 + (int64_t)q * ldG;
                *Gq = (j == 0 ? 0.0 : __ldcg(Gq)) + ssq[mi];
              }
            }
          }
        }
      }
      // ---- E6: publish ----
      if (prof && tid == 0) tp4 = (long long)gtime();
      __threadfence();
      asm volatile("fence.proxy.async;" ::: "memory");     // the tile is re-read by other CTAs' TMA bulk copies
      consumer_bar();
      if (tid == 0) st_release(pb->doneP + (int64_t)r * nCB + j, epoch);
      if (defer_inv) {                                           // (Ys is untouched until the next task's operands arrive)
        double* invd = pb->invd + (int64_t)j * INVD_BLOCK;
        for (int idx = tid; idx < INVD_BLOCK / 2; idx += NCW * 32)
          reinterpret_cast<double2*>(invd)[idx] = reinterpret_cast<const double2*>(Ys)[idx];
        __threadfence();
        asm volatile("fence.proxy.async;" ::: "memory");
        consumer_bar();
        if (tid == 0) st_release(pb->invdone + j, epoch);
      }
      if (prof && tid == 0) {
        prof[0] = tp0; prof[1] = tp1; prof[2] = tp2; prof[3] = tp3; prof[4] = tp4; prof[5] = (long long)gtime();
        prof[6] = (nchunks > kstart) ? s_tdep : tp0;
        prof[7] = ((long long)blockIdx.x << 48) | (kwait << 1) | (is_diag ? 1 : 0);   // CTA | cycles warp 0 waited for operands in the K-loop | diag
      }
    }
    it = it0 + (nchunks - kstart);
    at_phase ^= 1;
    __syncthreads();
  }
}

// lndet = sum_j logsum_j in fixed order; results[b] = {lndet, G(q*q)}; status copied alongside
__global__ void finish_kernel(const ProblemDesc* __restrict__ probs, int count, double* __restrict__ results, int stride) {
  const int b = blockIdx.x;
  if (b >= count) return;
  const ProblemDesc* pb = probs + b;
  double* out = results + (int64_t)b * stride;
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int j = 0; j < pb->nCB; j++) s += pb->logsum[j];
    out[0] = s;
    out[1] = (double)(pb->status[0] | (pb->status[1] << 1));
  }
  const int qq = pb->q * pb->q;
  for (int i = threadIdx.x; i < qq; i += blockDim.x) out[2 + i] = pb->G[i];
}

// ---- back-substitution  L' X = Y  for a few right-hand sides (iAW of fitIAK3D.R:782-808) ------------
// Y' lives in the bordered rows of the blocked factor buffer; it is first gathered to a compact q x Npad array.
// Block-column sweep from the right:  x_J = inv(L_JJ)' y_J ;  y_K -= L(J,K)' x_J for all K < J.
// Not on the optimiser's hot loop.
__global__ void gather_border_kernel(const double* __restrict__ L, int64_t nRU, int64_t Npad, int q, double* __restrict__ Yw) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= Npad) return;
  for (int w = 0; w < q; w++) Yw[c * q + w] = L[uidx(Npad + w, c, nRU)];
}
// One persistent kernel for the whole sweep.  Column block K is OWNED by CTA (K mod gridDim.x): the owner keeps y_K, applies
// the updates  y_K -= L(J,K)' x_J  for J = nCB-1 ... K+1 in that order as the x_J are published (per-block flags, release /
// acquire), then computes  x_K = inv(L_KK)' y_K  and publishes it.  At step J the owner of block J-1 serves that block
// first and publishes x_{J-1} before it turns to its other blocks, so the chain x_J -> y_{J-1} -> x_{J-1} never waits behind
// other work.  All CTAs are co-resident (the launcher caps the grid at the SM count); every wait is bounded.  Fixed summation
// order: deterministic.  Shared memory: three 64 x (q + 1) arrays (x_J, the look-ahead x_{J-1}, y of the block being solved).
__device__ __forceinline__ void bs_update_block(const double* __restrict__ L, int64_t nRU, int J, int K, const double* __restrict__ xs, int ldx,
                                                int q, double* __restrict__ Yw, int c, int g) {
  const double* col = L + uidx((int64_t)J * TN, (int64_t)K * TN + c, nRU);     // 64 consecutive rows of one unit
  for (int w0 = g; w0 < q; w0 += 16) {
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int i = 0; i < TN; i++) {
      const double l = col[i];
#pragma unroll
      for (int e = 0; e < 4; e++) { const int w = w0 + 4 * e; if (w < q) acc[e] += l * xs[i * ldx + w]; }
    }
#pragma unroll
    for (int e = 0; e < 4; e++) { const int w = w0 + 4 * e; if (w < q) Yw[(int64_t)(K * TN + c) * q + w] -= acc[e]; }
  }
}
// x_K = inv(L_KK)' y_K into xo (shared) and X (global), then the flag; called by the whole CTA
__device__ __forceinline__ void bs_solve_block(const double* __restrict__ invd, int K, const double* __restrict__ Yw, int q, double* __restrict__ ys,
                                               double* __restrict__ xo, int ldx, double* __restrict__ X, int64_t Npad, int32_t* flags, int tid) {
  const int c = tid & 63, g = tid >> 6;
  __syncthreads();                                  // the block's updates (global) and earlier readers of ys / xo
  for (int idx = tid; idx < TN * q; idx += 256) { const int k = idx / q, w = idx % q; ys[k * ldx + w] = Yw[(int64_t)(K * TN + k) * q + w]; }
  __syncthreads();
  const double* Li = invd + (int64_t)K * INVD_BLOCK;   // column-major 64 x 64 (stride LDB_S), lower: (invL')(c, k) = invL(k, c)
  for (int w = g; w < q; w += 4) {
    double s = 0.0;
    for (int k = c; k < TN; k++) s += Li[(int64_t)c * LDB_S + k] * ys[k * ldx + w];
    xo[c * ldx + w] = s;
    X[(int64_t)w * Npad + K * TN + c] = s;
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) st_release(flags + K, 1);
}
__global__ void __launch_bounds__(256) backsolve_sweep_kernel(const double* __restrict__ L, int64_t nRU, int64_t Npad, int nCB,
                                                              const double* __restrict__ invd, int q, double* __restrict__ Yw /* [col][q] */,
                                                              double* __restrict__ X /* Npad x q col-major */, int32_t* __restrict__ flags,
                                                              int32_t* __restrict__ status) {
  extern __shared__ double bsm[];
  const int ldx = q + 1;
  double* xs = bsm; double* xn = bsm + TN * ldx; double* ys = bsm + 2 * TN * ldx;
  const int tid = threadIdx.x, c = tid & 63, g = tid >> 6;
  const int b = (int)blockIdx.x, G = (int)gridDim.x;
  int32_t* abort_flag = status + 1;
  bool have_next = false;                           // x_J is already in xn (computed and published during step J + 1)
  for (int J = nCB - 1; J >= 0; J--) {
    const bool mine = (J % G == b);
    if (mine && !have_next) bs_solve_block(invd, J, Yw, q, ys, xn, ldx, X, Npad, flags, tid);
    if (b >= J) break;                              // no owned block below J: this CTA is done
    __syncthreads();
    if (mine) {
      for (int idx = tid; idx < TN * q; idx += 256) { const int i = idx / q, w = idx % q; xs[i * ldx + w] = xn[i * ldx + w]; }
    } else {
      if (tid == 0) wait_flag(flags + J, 1, status, abort_flag);
      __syncthreads();
      if (ld_volatile(abort_flag) != 0) return;
      for (int idx = tid; idx < TN * q; idx += 256) { const int i = idx % TN, w = idx / TN; xs[i * ldx + w] = __ldcg(X + (int64_t)w * Npad + J * TN + i); }
    }
    __syncthreads();
    have_next = false;
    if ((J - 1) % G == b) {                         // critical block first, and its x at once
      bs_update_block(L, nRU, J, J - 1, xs, ldx, q, Yw, c, g);
      bs_solve_block(invd, J - 1, Yw, q, ys, xn, ldx, X, Npad, flags, tid);
      have_next = true;
    }
    for (int K = b; K < J - 1; K += G) bs_update_block(L, nRU, J, K, xs, ldx, q, Yw, c, g);
  }
}

// ---- FP64 peak microbenchmarks -------------------------------------------------------------------------
__global__ void __launch_bounds__(256) peak_dfma_kernel(double* out, int iters) {
  double a[16];
#pragma unroll
  for (int i = 0; i < 16; i++) a[i] = 1.0 + 1e-9 * (threadIdx.x + i);
  const double b = 1.0000001, c = 1e-12;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = fma(a[i], b, c);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) peak_dmma_kernel(double* out, int iters) {
  double acc[8][2];
#pragma unroll
  for (int i = 0; i < 8; i++) { acc[i][0] = 0.0; acc[i][1] = 0.0; }
  const double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) dmma(acc[i][0], acc[i][1], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += acc[i][0] + acc[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// kind 2: are the two FP64 paths one pipe?  Even warps issue DMMA, odd warps DFMA, 8 independent chains each.  If the
// combined rate exceeds either peak alone the pipes are separate and a mixed kernel could use both.
__global__ void __launch_bounds__(256) peak_mixed_kernel(double* out, int iters) {
  double s = 0.0;
  if (((threadIdx.x >> 5) & 1) == 0) {
    double acc[8][2];
#pragma unroll
    for (int i = 0; i < 8; i++) { acc[i][0] = 0.0; acc[i][1] = 0.0; }
    const double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int i = 0; i < 8; i++) dmma(acc[i][0], acc[i][1], a, b);
    }
#pragma unroll
    for (int i = 0; i < 8; i++) s += acc[i][0] + acc[i][1];
  } else {
    double a[16];
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = 1.0 + 1e-9 * (threadIdx.x + i);
    const double b = 1.0000001, c = 1e-12;
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int i = 0; i < 16; i++) a[i] = fma(a[i], b, c);
    }
#pragma unroll
    for (int i = 0; i < 16; i++) s += a[i];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

constexpr int PAIR_WARPS_DEFAULT = 4;     // MMA warps per CTA of the pair engine (IAK3D_PAIR_WARPS = 4 | 8 overrides)
#include "chol_pair.cuh"

}  // namespace

int chol_configure(iak3d_ctx* ctx) {
  CUDA_TRY(ctx, cudaFuncSetAttribute(tile_task_kernel<-1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  CUDA_TRY(ctx, cudaFuncSetAttribute(tile_task_kernel<-1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  CUDA_TRY(ctx, cudaFuncSetAttribute(tile_task_kernel<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  CUDA_TRY(ctx, cudaFuncSetAttribute(tile_task_kernel<2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  CUDA_TRY(ctx, cudaFuncSetAttribute(backsolve_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * TN * 65 * (int)sizeof(double)));
  CUDA_TRY(ctx, cudaFuncSetAttribute(tile_pair_kernel<0, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2_BYTES));
  CUDA_TRY(ctx, cudaFuncSetAttribute(tile_pair_kernel<1, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2_BYTES));
  CUDA_TRY(ctx, cudaFuncSetAttribute(tile_pair_kernel<2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2_BYTES));
  CUDA_TRY(ctx, cudaFuncSetAttribute(tile_pair_kernel<0, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2_BYTES));
  CUDA_TRY(ctx, cudaFuncSetAttribute(tile_pair_kernel<1, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2_BYTES));
  CUDA_TRY(ctx, cudaFuncSetAttribute(tile_pair_kernel<2, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2_BYTES));
  // The pair engine's MMA warps grow to 216 registers with what the producer warpgroup gives back (setmaxnreg); that only
  // adds up when the kernel is launched with exactly 128 registers per thread and two CTAs fit an SM.  Anything else
  // (a different compiler, a smaller device) keeps the one-CTA engine rather than risk a warp waiting for registers.
  ctx->engine = 1;
  cudaFuncAttributes fa0, fa1, fa2;
  int per_sm = 0;
  const char* pw = getenv("IAK3D_PAIR_WARPS");
  ctx->pair_warps = (pw && pw[0] == '4') ? 4 : ((pw && pw[0] == '8') ? 8 : PAIR_WARPS_DEFAULT);
  for (int attempt = 0; attempt < 2 && ctx->engine == 1; attempt++) {
    if (ctx->pair_warps == 8) {       // 384-thread CTAs: 80 registers per thread at launch, 104 / 32 after the split
      CUDA_TRY(ctx, cudaFuncGetAttributes(&fa0, tile_pair_kernel<0, 8>));
      CUDA_TRY(ctx, cudaFuncGetAttributes(&fa1, tile_pair_kernel<1, 8>));
      CUDA_TRY(ctx, cudaFuncGetAttributes(&fa2, tile_pair_kernel<2, 8>));
      CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tile_pair_kernel<0, 8>, 384, SMEM2_BYTES));
      if (fa0.numRegs == 80 && fa1.numRegs == 80 && fa2.numRegs == 80 && per_sm >= 2) ctx->engine = 2; else ctx->pair_warps = 4;
    } else {                          // 256-thread CTAs: 128 at launch, 216 / 40 after the split
      CUDA_TRY(ctx, cudaFuncGetAttributes(&fa0, tile_pair_kernel<0, 4>));
      CUDA_TRY(ctx, cudaFuncGetAttributes(&fa1, tile_pair_kernel<1, 4>));
      CUDA_TRY(ctx, cudaFuncGetAttributes(&fa2, tile_pair_kernel<2, 4>));
      CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tile_pair_kernel<0, 4>, 256, SMEM2_BYTES));
      if (fa0.numRegs == 128 && fa1.numRegs == 128 && fa2.numRegs == 128 && per_sm >= 2) ctx->engine = 2;
      break;
    }
  }
  const char* e = getenv("IAK3D_ENGINE");
  if (e && e[0] == '1') ctx->engine = 1;
  if (e && e[0] == '3' && ctx->engine == 2) ctx->engine = 3;   // tests: single-problem launches through the pair engine too
  const char* w = getenv("IAK3D_PAIR_WIDTH");
  if (w && atoi(w) > 0) ctx->pair_min_width = atoi(w);
  const char* tqe = getenv("IAK3D_TABLE_QUEUE");
  if (tqe) ctx->no_table_queue = atoi(tqe) == 0;
  const char* dg = getenv("IAK3D_DIAG");
  if (dg) ctx->diag_chain = atoi(dg) != 0;
  return 0;
}

int launch_tile_tasks(iak3d_ctx* ctx, const ProblemDesc* d_problems, const int4* d_tasks, int32_t ntasks, int32_t* d_ticket, int mode, int latency_bound, int width) {
  if (ntasks <= 0) return 0;
  CUDA_TRY(ctx, cudaMemsetAsync(d_ticket, 0, 2 * sizeof(int32_t), ctx->stream));   // [0] ticket, [1] abort flag
  if (ctx->engine == 3 || (ctx->engine == 2 && !(mode == 0 && latency_bound) && width >= ctx->pair_min_width)) {
    // throughput launches -- at least pair_min_width independent tasks per dependency step, i.e. enough to keep two CTAs per
    // SM busy: two half-size CTAs per SM, one's epilogue under the other's K-loop (chol_pair.cuh).  Narrower launches are
    // bound by the per-column critical path, which the full-size CTA (8 MMA warps on one tile) runs faster: measured at
    // 9 x n = 2049 (width ~ 100) 18.2 vs 14.5 TFLOP/s, at 9 x n = 5000 (width ~ 270) 28.7 vs 30.3.
    int grid2 = 2 * ctx->sm_count;
    if (grid2 > ntasks) grid2 = ntasks;
    if (ctx->pair_warps == 8) {
      if (mode == 0) tile_pair_kernel<0, 8><<<grid2, 384, SMEM2_BYTES, ctx->stream>>>(d_problems, d_tasks, ntasks, d_ticket);
      else if (mode == 1) tile_pair_kernel<1, 8><<<grid2, 384, SMEM2_BYTES, ctx->stream>>>(d_problems, d_tasks, ntasks, d_ticket);
      else tile_pair_kernel<2, 8><<<grid2, 384, SMEM2_BYTES, ctx->stream>>>(d_problems, d_tasks, ntasks, d_ticket);
    } else {
      if (mode == 0) tile_pair_kernel<0, 4><<<grid2, 256, SMEM2_BYTES, ctx->stream>>>(d_problems, d_tasks, ntasks, d_ticket);
      else if (mode == 1) tile_pair_kernel<1, 4><<<grid2, 256, SMEM2_BYTES, ctx->stream>>>(d_problems, d_tasks, ntasks, d_ticket);
      else tile_pair_kernel<2, 4><<<grid2, 256, SMEM2_BYTES, ctx->stream>>>(d_problems, d_tasks, ntasks, d_ticket);
    }
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    return 0;
  }
  int grid = ctx->sm_count;
  if (grid > ntasks) grid = ntasks;
  // factorisations on the full-size CTA always take the warp-specialised diagonal block (LOOKAHEAD = 1; IAK3D_DIAG=0: the
  // four-column routine of the batched engines, kept for comparison -- same bits)
  if (mode == 0 && (latency_bound || ctx->diag_chain)) tile_task_kernel<-1, 1><<<grid, NTHREADS, SMEM_BYTES, ctx->stream>>>(d_problems, d_tasks, ntasks, d_ticket);
  else if (mode == 0) tile_task_kernel<-1, 0><<<grid, NTHREADS, SMEM_BYTES, ctx->stream>>>(d_problems, d_tasks, ntasks, d_ticket);
  else if (mode == 1) tile_task_kernel<1, 0><<<grid, NTHREADS, SMEM_BYTES, ctx->stream>>>(d_problems, d_tasks, ntasks, d_ticket);
  else tile_task_kernel<2, 0><<<grid, NTHREADS, SMEM_BYTES, ctx->stream>>>(d_problems, d_tasks, ntasks, d_ticket);
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  return 0;
}

int launch_finish(iak3d_ctx* ctx, const ProblemDesc* d_problems, int32_t count, double* d_results, int32_t stride) {
  if (count <= 0) return 0;
  finish_kernel<<<count, 128, 0, ctx->stream>>>(d_problems, count, d_results, stride);
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  return 0;
}

int launch_backsolve(iak3d_ctx* ctx, const double* d_L, int64_t nRU, int64_t Npad, const double* d_invd, int32_t q, double* d_X) {
  // d_X: Npad x q column-major out.  Two launches, pooled scratch, no synchronisation: the caller's copy follows on the stream.
  if (q <= 0 || q > 64) { ctx->err = "backsolve: 1 <= q <= 64"; return IAK3D_ERR_ARG; }
  const int nCB = (int)(Npad / TN);
  double* d_Yw = nullptr; int32_t* d_fl = nullptr;
  if (iak_pool_alloc(ctx, (void**)&d_Yw, (size_t)q * Npad * sizeof(double)) != cudaSuccess ||
      iak_pool_alloc(ctx, (void**)&d_fl, ((size_t)nCB + 2) * sizeof(int32_t)) != cudaSuccess) {
    iak_pool_free(ctx, d_Yw); ctx->err = "backsolve: allocation failed"; cudaGetLastError(); return IAK3D_ERR_CUDA;
  }
  cudaMemsetAsync(d_fl, 0, ((size_t)nCB + 2) * sizeof(int32_t), ctx->stream);
  gather_border_kernel<<<(unsigned)((Npad + 127) / 128), 128, 0, ctx->stream>>>(d_L, nRU, Npad, q, d_Yw);
  const int grid = std::min(nCB, ctx->sm_count);
  const size_t smem = (size_t)3 * TN * (q + 1) * sizeof(double);
  backsolve_sweep_kernel<<<grid, 256, smem, ctx->stream>>>(d_L, nRU, Npad, nCB, d_invd, q, d_Yw, d_X, d_fl, d_fl + nCB);
  ctx->launches += 2;
  const cudaError_t e = cudaGetLastError();
  iak_pool_free(ctx, d_Yw); iak_pool_free(ctx, d_fl);
  CUDA_TRY(ctx, e);
  return 0;
}

int launch_peak(iak3d_ctx* ctx, int kind, double* tflops) {
  const int blocks = ctx->sm_count * 4, threads = 256, iters = 20000;
  double* d_out = nullptr;
  CUDA_TRY(ctx, cudaMalloc(&d_out, (size_t)blocks * threads * sizeof(double)));
  cudaEvent_t e0, e1;
  CUDA_TRY(ctx, cudaEventCreate(&e0)); CUDA_TRY(ctx, cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {
    CUDA_TRY(ctx, cudaEventRecord(e0, ctx->stream));
    if (kind == 0) peak_dfma_kernel<<<blocks, threads, 0, ctx->stream>>>(d_out, iters);
    else if (kind == 1) peak_dmma_kernel<<<blocks, threads, 0, ctx->stream>>>(d_out, iters);
    else peak_mixed_kernel<<<blocks, threads, 0, ctx->stream>>>(d_out, iters);
    ctx->launches++;
    CUDA_TRY(ctx, cudaEventRecord(e1, ctx->stream));
    CUDA_TRY(ctx, cudaEventSynchronize(e1));
    float ms = 0; CUDA_TRY(ctx, cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d_out);
  double flops;
  if (kind == 0) flops = (double)blocks * threads * iters * 16.0 * 2.0;
  else if (kind == 1) flops = (double)blocks * (threads / 32) * iters * 8.0 * (8.0 * 8.0 * 4.0 * 2.0);
  else flops = 0.5 * (double)blocks * threads * iters * 16.0 * 2.0 + 0.5 * (double)blocks * (threads / 32) * iters * 8.0 * (8.0 * 8.0 * 4.0 * 2.0);
  *tflops = flops / (best * 1e-3) / 1e12;
  return 0;
}
