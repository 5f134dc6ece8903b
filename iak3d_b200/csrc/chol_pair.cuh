// chol_pair.cuh -- the throughput instantiation of the tile engine: TWO co-resident CTAs per SM.
// Included by chol.cu inside its anonymous namespace (shares the mbarrier / TMA / flag helpers and the constants).
//
// Why.  tile_task_kernel (one CTA of 8 MMA warps per SM) leaves the FP64 tensor pipe idle for the whole epilogue of
// every task (tile = A - acc, wait for / fetch of the inverse diagonal block, the triangular solve, copy-out, publish:
// ~11 us of a 126 us task at n = 5000) and for the pipeline fill of the next one: ncu showed the pipe active 86 % of
// the kernel's cycles.  Here a task is run by a CTA of FOUR MMA warps (each owns 32 rows x all 64 columns of the
// 128 x 64 tile: 128 accumulator registers) plus one TMA producer warp, with half the shared memory (two 52 KB stages),
// and two such CTAs live on every SM: while one is in its epilogue the other's K-loop has the tensor pipe to itself,
// and when both are in their K-loops the hardware interleaves their DMMAs, which also covers each other's operand
// waits.  One DMMA per sub-partition every 16 cycles is all the pipe takes, so one MMA warp per sub-partition and CTA
// saturates it.
//
// Registers.  128 accumulators + fragments do not fit the 128 registers per thread that two 256-thread CTAs per SM
// allow, so the CTA is launched with two warpgroups: the MMA warps raise their allocation to 216
// (setmaxnreg.inc) with what the producer warpgroup (one working warp, three that exit at once) gives back
// (setmaxnreg.dec to 40): 128 x (216 + 40) = the 32,768 registers of the CTA.
//
// Shared memory.  There is no resident copy of the task's own tile: when the K-loop has drained, the two stages
// (6 units) are re-used as  T = units 0..3 (the tile, fetched by TMA then)  and  Ys = units 4..5 (the inverse diagonal
// block, or potf2's scratch + output; later the W rows of the kriging epilogue).  The latency of that fetch is covered
// by the other CTA.
//
// Single-problem launches (the optimiser's line searches: latency, not throughput) keep tile_task_kernel<-1, 1>.

// NW = MMA warps per CTA: 4 (32 rows x 64 columns each, 128 accumulator registers; 216 / 40 registers) or 8 (16 rows each;
// 104 / 32 registers; 384-thread CTAs launched with 80 registers per thread).  One DMMA warp per sub-partition cannot keep
// its share of the pipe full on its own (a warp issues a DMMA every ~32 cycles, the pipe takes one every 16), which shows
// whenever the partner CTA is in an epilogue; with NW = 8 a CTA has two warps per sub-partition.
constexpr int STAGES2 = 2;
constexpr int SMEM2_BYTES = STAGES2 * STAGE_D * 8;      // 104,448 B = 6 units
static_assert(STAGES2 * STAGE_D == 6 * UNIT, "epilogue map needs six units");
static_assert(INVD_BLOCK == 2 * UNIT, "inverse diagonal block = two units");

template <int NW> __device__ __forceinline__ void pair_bar() { asm volatile("bar.sync 1, %0;" ::"n"(NW * 32) : "memory"); }
template <int NW> __device__ __forceinline__ void task_bar() { asm volatile("bar.sync 2, %0;" ::"n"((NW + 1) * 32) : "memory"); }

// one 32-deep chunk of a warp's CNT (1..4) live 8-row tiles x 64 columns; ao[mi] = offset of row tile mi inside a stage's A
// part.  CNT is a template parameter, not a predicate: mma.sync under a per-instruction predicate makes the compiler put a
// WARPSYNC in front of every DMMA (measured: 12.4 -> 13.9 ms for the S5 step).
template <int MI, int CNT>
__device__ __forceinline__ void warp_mma2_chunk(double (&acc)[MI][8][2], const double* __restrict__ Ab, const int (&ao)[MI],
                                                const double* __restrict__ Bb) {
#pragma unroll
  for (int k4 = 0; k4 < KC / 4; k4++) {
    double a[CNT], b[8];
#pragma unroll
    for (int mi = 0; mi < CNT; mi++) a[mi] = Ab[k4 * 4 * UL + ao[mi]];
#pragma unroll
    for (int ni = 0; ni < 8; ni++) b[ni] = Bb[k4 * 4 * UL + ni * 8];
#pragma unroll
    for (int ni = 0; ni < 8; ni++)
#pragma unroll
      for (int mi = 0; mi < CNT; mi++) dmma(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
  }
}

// E3: the eight k of block KB of  X = T inv(L_jj)'  for the column tiles KB..7 (inv(L_jj)' is upper triangular)
template <int MI, int KB>
__device__ __forceinline__ void solve2_block(double (&x)[MI][8][2], const double* __restrict__ Ta, const int (&ao)[MI],
                                             const double* __restrict__ Yb) {
#pragma unroll
  for (int h = 0; h < 2; h++) {
    const int k0 = KB * 8 + h * 4;
    const double* Ab = Ta + (k0 >> 5) * 2 * UNIT + (k0 & 31) * UL;
    const double* Bb = Yb + k0 * LDB_S;
    double a[MI], b[8];
#pragma unroll
    for (int mi = 0; mi < MI; mi++) a[mi] = Ab[ao[mi]];
#pragma unroll
    for (int ni = KB; ni < 8; ni++) b[ni] = Bb[ni * 8];
#pragma unroll
    for (int ni = KB; ni < 8; ni++)
#pragma unroll
      for (int mi = 0; mi < MI; mi++) dmma(x[mi][ni][0], x[mi][ni][1], a[mi], b[ni]);   // slots without a tile: computed, never stored
  }
}

// potf2 fused with the inverse, as potf2_inv_blk4, for 128 threads: thread (b2, bc) of an 8 x 16 grid owns the 4 x 4
// blocks (b2, bc) and (b2 + 8, bc).  scratch (2 x BLK4_BUF doubles) may alias Ys: the inverse is written there only after
// a barrier behind the last read of the scratch.
__device__ void potf2_inv_blk4_h2(double* __restrict__ At, int off, double* __restrict__ Ys, double* __restrict__ scratch,
                                  double* __restrict__ pivs /* 64 */, int tid, double* logsum_out, int32_t* status) {
  const int b2 = tid >> 4, bc = tid & 15;
  double d[2][4][4], y[2][4][4];
#pragma unroll
  for (int h = 0; h < 2; h++) {
    const int br = b2 + 8 * h;
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int jj = 0; jj < 4; jj++) {
        d[h][i][jj] = (bc <= br) ? At[AT_IDX(off + 4 * br + i, 4 * bc + jj)] : 0.0;
        y[h][i][jj] = (br == bc && i == jj) ? 1.0 : 0.0;
      }
  }
  bool bad = false;
#pragma unroll 1
  for (int cb = 0; cb < TN / 4; cb++) {
    double* buf = scratch + (cb & 1) * BLK4_BUF;
#pragma unroll
    for (int h = 0; h < 2; h++) {
      if (bc == cb && b2 + 8 * h == cb) {
        double l[6], r[4], pv[4];
        pv[0] = d[h][0][0]; r[0] = fast_rsqrt(pv[0]);
        l[0] = d[h][1][0] * r[0]; l[1] = d[h][2][0] * r[0]; l[2] = d[h][3][0] * r[0];
        pv[1] = d[h][1][1] - l[0] * l[0]; r[1] = fast_rsqrt(pv[1]);
        l[3] = (d[h][2][1] - l[1] * l[0]) * r[1]; l[4] = (d[h][3][1] - l[2] * l[0]) * r[1];
        pv[2] = d[h][2][2] - l[1] * l[1] - l[3] * l[3]; r[2] = fast_rsqrt(pv[2]);
        l[5] = (d[h][3][2] - l[2] * l[1] - l[4] * l[3]) * r[2];
        pv[3] = d[h][3][3] - l[2] * l[2] - l[4] * l[4] - l[5] * l[5]; r[3] = fast_rsqrt(pv[3]);
#pragma unroll
        for (int k = 0; k < 4; k++) { if (!(pv[k] > 0.0) || isinf(pv[k])) bad = true; pivs[4 * cb + k] = pv[k]; }
#pragma unroll
        for (int k = 0; k < 6; k++) buf[BLK4_L + k] = l[k];
#pragma unroll
        for (int k = 0; k < 4; k++) buf[BLK4_L + 6 + k] = r[k];
        d[h][0][0] = pv[0] * r[0]; d[h][0][1] = 0.0; d[h][0][2] = 0.0; d[h][0][3] = 0.0;
        d[h][1][0] = l[0]; d[h][1][1] = pv[1] * r[1]; d[h][1][2] = 0.0; d[h][1][3] = 0.0;
        d[h][2][0] = l[1]; d[h][2][1] = l[3]; d[h][2][2] = pv[2] * r[2]; d[h][2][3] = 0.0;
        d[h][3][0] = l[2]; d[h][3][1] = l[4]; d[h][3][2] = l[5]; d[h][3][3] = pv[3] * r[3];
#pragma unroll
        for (int jj = 0; jj < 4; jj++) {
          blk4_subst_row(l, r, y[h][0][jj], y[h][1][jj], y[h][2][jj], y[h][3][jj]);
#pragma unroll
          for (int k = 0; k < 4; k++) buf[BLK4_Y + k * 64 + 4 * bc + jj] = y[h][k][jj];
        }
      }
    }
    pair_bar<4>();
    {
      double l[6], r[4];
#pragma unroll
      for (int k = 0; k < 6; k++) l[k] = buf[BLK4_L + k];
#pragma unroll
      for (int k = 0; k < 4; k++) r[k] = buf[BLK4_L + 6 + k];
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int br = b2 + 8 * h;
        if (bc <= br && (bc == cb) != (br == cb)) {
          if (bc == cb) {       // panel block below the diagonal block
#pragma unroll
            for (int i = 0; i < 4; i++) {
              blk4_subst_row(l, r, d[h][i][0], d[h][i][1], d[h][i][2], d[h][i][3]);
#pragma unroll
              for (int k = 0; k < 4; k++) buf[BLK4_P + (4 * br + i) * 4 + k] = d[h][i][k];
            }
          } else {              // block row cb of the inverse, left of the diagonal
#pragma unroll
            for (int jj = 0; jj < 4; jj++) {
              blk4_subst_row(l, r, y[h][0][jj], y[h][1][jj], y[h][2][jj], y[h][3][jj]);
#pragma unroll
              for (int k = 0; k < 4; k++) buf[BLK4_Y + k * 64 + 4 * bc + jj] = y[h][k][jj];
            }
          }
        }
      }
    }
    pair_bar<4>();
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const int br = b2 + 8 * h;
      if (bc <= br && br > cb) {
        double pr[4][4];
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
          for (int k = 0; k < 4; k++) pr[i][k] = buf[BLK4_P + (4 * br + i) * 4 + k];
        if (bc > cb) {        // trailing block: d -= P_br P_bc'
#pragma unroll
          for (int jj = 0; jj < 4; jj++) {
            double pc[4];
#pragma unroll
            for (int k = 0; k < 4; k++) pc[k] = buf[BLK4_P + (4 * bc + jj) * 4 + k];
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
              for (int k = 0; k < 4; k++) d[h][i][jj] -= pr[i][k] * pc[k];
          }
        } else {              // inverse block (br, bc <= cb): y -= P_br Y(cb, bc)
#pragma unroll
          for (int jj = 0; jj < 4; jj++) {
            double yb[4];
#pragma unroll
            for (int k = 0; k < 4; k++) yb[k] = buf[BLK4_Y + k * 64 + 4 * bc + jj];
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
              for (int k = 0; k < 4; k++) y[h][i][jj] -= pr[i][k] * yb[k];
          }
        }
      }
    }
  }
  pair_bar<4>();               // the scratch may alias Ys
#pragma unroll
  for (int h = 0; h < 2; h++) {
    const int br = b2 + 8 * h;
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int jj = 0; jj < 4; jj++) {
        const int r = 4 * br + i, cc = 4 * bc + jj;
        At[AT_IDX(off + r, cc)] = (bc <= br && r >= cc) ? d[h][i][jj] : 0.0;
        Ys[cc * LDB_S + r] = (bc <= br && r >= cc) ? y[h][i][jj] : 0.0;
      }
  }
  if (bad) atomicExch(status, 1);
  pair_bar<4>();
  if (tid < 32) {   // log-determinant of the block: fixed-order tree, deterministic
    double v = log(pivs[tid]) + log(pivs[tid + 32]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (tid == 0) *logsum_out = v;
  }
}

template <int MODE, int NW>
__global__ void __launch_bounds__((NW + 4) * 32, 2)
tile_pair_kernel(const ProblemDesc* __restrict__ probs, const int4* __restrict__ tasks, int ntasks, int32_t* ticket) {
  constexpr int MI = 16 / NW;                   // 8-row tiles per MMA warp
  extern __shared__ __align__(128) double smem[];
  double* const T = smem;                       // epilogue: the task's own tile, 4 units  [col half][row half]
  double* const Ys = smem + 4 * UNIT;           // epilogue: inverse diagonal block (2 units)
  __shared__ __align__(8) uint64_t full_bar[STAGES2];
  __shared__ __align__(8) uint64_t empty_bar[STAGES2];
  __shared__ __align__(8) uint64_t at_bar;
  __shared__ __align__(8) uint64_t inv_bar;
  __shared__ __align__(16) double s_pivs[64];
  __shared__ int s_task;
  __shared__ long long s_tdep;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < STAGES2; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], NW); }
    mbar_init(&at_bar, 1);
    mbar_init(&inv_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  int32_t* const abort_flag = ticket + 1;
  uint32_t it = 0;          // pipeline chunk counter, identical in every thread
  // The two roles never share code after this point: ptxas budgets registers per branch of a setmaxnreg pair only when
  // the branches do not merge again.
  if (warp >= NW) {
    if constexpr (NW == 4) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    else asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
    if (warp > NW) return;
    // ===================== producer warp: tickets + TMA bulk copies, issued by one lane =====================
    while (true) {
      if (lane == 0) s_task = (ld_volatile(abort_flag) != 0) ? ntasks : atomicAdd(ticket, 1);
      task_bar<NW>();
      const int t = s_task;
      if (t >= ntasks) break;
      const int4 tk = tasks[t];
      const ProblemDesc* pb = probs + tk.x;
      const int r = tk.y, j = tk.z;
      const int mode = MODE;
      int32_t* status = pb->status;
      const int nchunks = (mode == 2) ? pb->kchunks : j * (TN / KC);
      const int kstart = min(tk.w, nchunks);
      const uint32_t it0 = it;
      if (lane == 0) {
        const double* __restrict__ L = pb->L; const int64_t nRUL = pb->nRUL;
        const double* __restrict__ Asrc = (mode == 2) ? pb->A : pb->P;
        const int64_t nRUA = (mode == 2) ? pb->nRUA : pb->nRUP;
        const int nCB = pb->nCB, epoch = pb->epoch;
        // the previous task's epilogue wrote the ring through the generic proxy (ordered by the end-of-task barrier)
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        uint32_t itp = it0;
        for (int c = kstart; c < nchunks; c++, itp++) {
          const int k = c >> 1;
          if ((c & 1) == 0 && mode != 2) {
            wait_flag(pb->doneP + (int64_t)r * nCB + k, epoch, status, abort_flag);
            if (mode == 0) wait_flag(pb->doneL + (int64_t)(j >> 1) * nCB + k, epoch, status, abort_flag);
            asm volatile("fence.proxy.async;" ::: "memory");
            if (pb->prof && k == j - 1) s_tdep = (long long)gtime();
          }
          const uint32_t s = itp % STAGES2, use = itp / STAGES2;
          if (use > 0) mbar_wait(&empty_bar[s], (use - 1) & 1, status, abort_flag);
          mbar_expect_tx(&full_bar[s], STAGE_D * 8);
          double* As = smem + s * STAGE_D;
          bulk_g2s(As, Asrc + ((int64_t)c * nRUA + 2 * r) * UNIT, A_STAGE * 8, &full_bar[s]);
          bulk_g2s(As + A_STAGE, L + ((int64_t)c * nRUL + j) * UNIT, B_STAGE * 8, &full_bar[s]);
        }
      }
      __syncwarp();
      it = it0 + (nchunks - kstart);
      task_bar<NW>();
    }
    return;
  }
  if constexpr (NW == 4) asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
  else asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
  // ===================== MMA warps: DMMA K-loop + epilogue =====================
  uint32_t at_phase = 0, inv_phase = 0;
  while (true) {
    task_bar<NW>();
    const int t = s_task;
    if (t >= ntasks) break;
    const int4 tk = tasks[t];
    const ProblemDesc* pb = probs + tk.x;
    const int r = tk.y, j = tk.z;
    const int mode = MODE;
    int32_t* status = pb->status;
    const int nchunks = (mode == 2) ? pb->kchunks : j * (TN / KC);
    const int kstart = min(tk.w, nchunks);
    const uint32_t it0 = it;
    double* __restrict__ P = pb->P; const int64_t nRUP = pb->nRUP;
    const int nCB = pb->nCB, epoch = pb->epoch;
    const int64_t R0 = (int64_t)r * TM;
    long long* prof = pb->prof ? pb->prof + (int64_t)t * 8 : nullptr;
    long long tp0 = 0, tp1 = 0, tp2 = 0, tp3 = 0, tp4 = 0, kwait = 0;
    if (prof && tid == 0) tp0 = (long long)gtime();
    double* const Pt0 = P + ((int64_t)(2 * j) * nRUP + 2 * r) * UNIT;
    double* const Pt1 = P + ((int64_t)(2 * j + 1) * nRUP + 2 * r) * UNIT;
    const int kq = lane & 3, rq = lane >> 2;
    const bool is_diag = (mode == 0) && (r == (j >> 1));
    // Row tiles (8 rows each, 16 per task) that carry data are DEALT to the four warps, so that a partly empty task costs
    // every warp the same: rows above the diagonal block of an odd column's diagonal tile, the identity padding [n, Npad)
    // and the rows behind the W rows never enter the K-loop (at n = 5000 the last row block holds 17 live rows of 128).
    int tl[MI], ao[MI], cnt = 0;
    {
      unsigned live = 0xffffu;
      if (mode == 0) {
        const int nl = pb->nlive, w0 = pb->w0, wq = pb->q;
        if (nl > 0) {
          live = 0u;
          for (int tt = 0; tt < 16; tt++) {
            const int64_t g0 = R0 + 8 * tt;
            if (g0 < nl || (g0 + 7 >= w0 && g0 < w0 + wq)) live |= 1u << tt;
          }
        }
        if (is_diag && (j & 1)) live &= 0xff00u;        // upper 64 rows lie above the diagonal block
        if (is_diag) live |= (j & 1) ? 0xff00u : 0x00ffu; // the diagonal block itself is always whole (identity padding included)
      }
      int k = 0;
#pragma unroll
      for (int mi = 0; mi < MI; mi++) { tl[mi] = -1; }
      for (int tt = 0; tt < 16; tt++)
        if ((live >> tt) & 1u) {
          if ((k % NW) == warp) {
            const int slot = k / NW;
#pragma unroll
            for (int mi = 0; mi < MI; mi++) if (mi == slot) tl[mi] = tt;
            cnt = slot + 1;
          }
          k++;
        }
#pragma unroll
      for (int mi = 0; mi < MI; mi++) {
        const int tt = tl[mi] < 0 ? 0 : tl[mi];
        ao[mi] = (tt >> 3) * UNIT + (tt & 7) * 8 + rq;           // unit of the row half + row inside the unit
      }
    }
    double acc[MI][8][2];
#pragma unroll
    for (int mi = 0; mi < MI; mi++)
#pragma unroll
      for (int ni = 0; ni < 8; ni++) { acc[mi][ni][0] = 0.0; acc[mi][ni][1] = 0.0; }
    uint32_t itc = it0;
    for (int c = kstart; c < nchunks; c++, itc++) {
      const uint32_t s = itc % STAGES2;
      if (prof && tid == 0) {
        const long long c0 = clock64();
        mbar_wait(&full_bar[s], (itc / STAGES2) & 1, status, abort_flag);
        kwait += clock64() - c0;
      } else {
        mbar_wait(&full_bar[s], (itc / STAGES2) & 1, status, abort_flag);
      }
      {
        const double* Ab = smem + s * STAGE_D + kq * UL;
        const double* Bb = smem + s * STAGE_D + A_STAGE + kq * UL + rq;
        if constexpr (MI == 4) {
          if (cnt == 4) warp_mma2_chunk<MI, 4>(acc, Ab, ao, Bb);
          else if (cnt == 3) warp_mma2_chunk<MI, 3>(acc, Ab, ao, Bb);
          else if (cnt == 2) warp_mma2_chunk<MI, 2>(acc, Ab, ao, Bb);
          else if (cnt == 1) warp_mma2_chunk<MI, 1>(acc, Ab, ao, Bb);
        } else {
          if (cnt == 2) warp_mma2_chunk<MI, 2>(acc, Ab, ao, Bb);
          else if (cnt == 1) warp_mma2_chunk<MI, 1>(acc, Ab, ao, Bb);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[s]);
    }
    if (prof && tid == 0) tp1 = (long long)gtime();
    // ---- the ring has drained: fetch the task's own tile (and the inverse diagonal block) into it ----
    pair_bar<NW>();
    if (tid == 0) {
      mbar_expect_tx(&at_bar, 4 * UNIT * 8);
      bulk_g2s(T, Pt0, 2 * UNIT * 8, &at_bar);
      bulk_g2s(T + 2 * UNIT, Pt1, 2 * UNIT * 8, &at_bar);
      if (mode != 2 && !is_diag) {
        if (mode == 0) wait_flag(pb->invdone + j, epoch, status, abort_flag);
        asm volatile("fence.proxy.async;" ::: "memory");
        mbar_expect_tx(&inv_bar, INVD_BLOCK * 8);
        bulk_g2s(Ys, pb->invd + (int64_t)j * INVD_BLOCK, INVD_BLOCK * 8, &inv_bar);
      }
    }
    mbar_wait(&at_bar, at_phase, status, abort_flag);
    at_phase ^= 1;
    // ---- E1: tile = A - acc (each warp touches its own 32 rows only) ----
    if (mode == 2) {
      const double alpha = pb->alpha; const bool cin = pb->use_cin != 0;
#pragma unroll
      for (int mi = 0; mi < MI; mi++)
#pragma unroll
        for (int ni = 0; ni < 8; ni++) {
          const int row = tl[mi] * 8 + rq, col = ni * 8 + kq * 2;      // mode 2: every tile is live
          const int i0 = AT_IDX(row, col), i1 = AT_IDX(row, col + 1);
          T[i0] = (cin ? T[i0] : 0.0) + alpha * acc[mi][ni][0];
          T[i1] = (cin ? T[i1] : 0.0) + alpha * acc[mi][ni][1];
        }
      pair_bar<NW>();
      for (int idx = tid; idx < 4 * (UNIT / 2); idx += NW * 32) {
        const int u = idx / (UNIT / 2), w2 = idx % (UNIT / 2);
        reinterpret_cast<double2*>(((u >> 1) ? Pt1 : Pt0) + (u & 1) * UNIT)[w2] = reinterpret_cast<const double2*>(T + u * UNIT)[w2];
      }
    } else {
#pragma unroll
      for (int mi = 0; mi < MI; mi++)
        if (mi < cnt) {
#pragma unroll
          for (int ni = 0; ni < 8; ni++) {
            const int row = tl[mi] * 8 + rq, col = ni * 8 + kq * 2;
            T[AT_IDX(row, col)] -= acc[mi][ni][0];
            T[AT_IDX(row, col + 1)] -= acc[mi][ni][1];
          }
        }
      if (prof && tid == 0) tp2 = (long long)gtime();
      // ---- E2: diagonal block, or the inverted diagonal block arrives ----
      int row_lo = 0;
      if (is_diag) {
        const int off = (j & 1) * TN;
        pair_bar<NW>();
        if constexpr (NW == 4) potf2_inv_blk4_h2(T, off, Ys, Ys, s_pivs, tid, pb->logsum + j, status);
        else potf2_inv_blk4<true>(T, off, Ys, Ys, s_pivs, tid, pb->logsum + j, status);
        double* invd = pb->invd + (int64_t)j * INVD_BLOCK;
        for (int idx = tid; idx < INVD_BLOCK / 2; idx += NW * 32)
          reinterpret_cast<double2*>(invd)[idx] = reinterpret_cast<const double2*>(Ys)[idx];
        for (int idx = tid; idx < UNIT; idx += NW * 32) {        // UNIT double2 = the two units (column halves) of L_jj
          const int ch = idx / (UNIT / 2), w2 = idx % (UNIT / 2);
          const double2 v = reinterpret_cast<const double2*>(T + (ch * 2 + (off >> 6)) * UNIT)[w2];
          reinterpret_cast<double2*>((ch ? Pt1 : Pt0) + (off >> 6) * UNIT)[w2] = v;
        }
        __threadfence();
        asm volatile("fence.proxy.async;" ::: "memory");   // later readers fetch this block with the TMA engine
        pair_bar<NW>();
        if (tid == 0) st_release(pb->invdone + j, epoch);
        row_lo = off + TN;
      } else {
        mbar_wait(&inv_bar, inv_phase, status, abort_flag);
        inv_phase ^= 1;
        __syncwarp();
      }
      if (prof && tid == 0) tp3 = (long long)gtime();
      const bool wfetch = (mode == 1) && (pb->q > 0);
      if (row_lo < TM) {
        // ---- E3: X = tile * inv(L_jj)' on the tensor pipe; a warp reads and rewrites its own 32 rows of T ----
        // row tiles below the diagonal block (diagonal tiles: the block itself was finished by potf2)
        int cnt3 = 0;
#pragma unroll
        for (int mi = 0; mi < MI; mi++) if (mi < cnt && tl[mi] * 8 >= row_lo) cnt3 = mi + 1;
        int lo3 = 0;
#pragma unroll
        for (int mi = 0; mi < MI; mi++) if (mi < cnt && tl[mi] * 8 < row_lo) lo3 = mi + 1;     // dealt in ascending order: dead ones first
        if (cnt3 > lo3) {
          double x[MI][8][2];
#pragma unroll
          for (int mi = 0; mi < MI; mi++)
#pragma unroll
            for (int ni = 0; ni < 8; ni++) { x[mi][ni][0] = 0.0; x[mi][ni][1] = 0.0; }
          const double* Ta = T + kq * UL;
          const double* Yb = Ys + kq * LDB_S + rq;
          solve2_block<MI, 0>(x, Ta, ao, Yb); solve2_block<MI, 1>(x, Ta, ao, Yb); solve2_block<MI, 2>(x, Ta, ao, Yb); solve2_block<MI, 3>(x, Ta, ao, Yb);
          solve2_block<MI, 4>(x, Ta, ao, Yb); solve2_block<MI, 5>(x, Ta, ao, Yb); solve2_block<MI, 6>(x, Ta, ao, Yb); solve2_block<MI, 7>(x, Ta, ao, Yb);
          __syncwarp();
#pragma unroll
          for (int mi = 0; mi < MI; mi++)
            if (mi >= lo3 && mi < cnt3) {
#pragma unroll
              for (int ni = 0; ni < 8; ni++) {
                const int row = tl[mi] * 8 + rq, col = ni * 8 + kq * 2;
                T[AT_IDX(row, col)] = x[mi][ni][0];
                T[AT_IDX(row, col + 1)] = x[mi][ni][1];
              }
            }
        }
        pair_bar<NW>();
        if (wfetch && tid == 0) {
          // kriging epilogue: the two units of the factor buffer that hold the W rows under this column block replace
          // the inverse diagonal block (every warp is past its solve)
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          const int64_t wu = (int64_t)pb->w0 / UR;
          mbar_expect_tx(&inv_bar, 2 * UNIT * 8);
          bulk_g2s(Ys, pb->L + ((int64_t)(2 * j) * pb->nRUL + wu) * UNIT, UNIT * 8, &inv_bar);
          bulk_g2s(Ys + UNIT, pb->L + ((int64_t)(2 * j + 1) * pb->nRUL + wu) * UNIT, UNIT * 8, &inv_bar);
        }
        if (prof && tid == 0 && mode == 1) s_tdep = (long long)gtime();
        // ---- E4: copy-out of row halves [row_lo/64, 2): whole units, coalesced 16-byte stores ----
        {
          const int rh0 = row_lo >> 6, nun = (2 - rh0) * 2;
          for (int idx = tid; idx < nun * (UNIT / 2); idx += NW * 32) {
            const int u = idx / (UNIT / 2), w2 = idx % (UNIT / 2);
            const int ch = u / (2 - rh0), rh = rh0 + u % (2 - rh0);
            const double2 v = reinterpret_cast<const double2*>(T + (ch * 2 + rh) * UNIT)[w2];
            reinterpret_cast<double2*>((ch ? Pt1 : Pt0) + rh * UNIT)[w2] = v;
          }
        }
        // ---- E5: border reductions (ordered by the flag chain => deterministic, no atomics) ----
        const int q = pb->q;
        if (mode == 0) {
          const int wl0 = pb->w0 - (int)R0;
          if (q > 0 && wl0 >= row_lo && wl0 < TM) {
            for (int e = tid; e < q * q; e += NW * 32) {
              const int a = e % q, b = e / q;
              double s = 0.0;
              for (int c = 0; c < TN; c++) s += T[AT_IDX(wl0 + a, c)] * T[AT_IDX(wl0 + b, c)];
              double* G = pb->G + e;            // __ldcg: the previous column's CTA wrote it through L2
              *G = (j == 0 ? 0.0 : __ldcg(G)) + s;
            }
          }
        } else if (wfetch) {
          // kriging rows: G(:,0..q-1) += X Yw', G(:,q) += rowsum(X^2) on the tensor pipe: the warp's 32 rows are four
          // m8 tiles, the W rows the B operand (element (w, c) at Ws[(c>>5)*UNIT + (c&31)*UL + w]), 16 W rows per pass
          mbar_wait(&inv_bar, inv_phase, status, abort_flag);
          inv_phase ^= 1;
          const double* Ws = Ys;
          double* const Gr = pb->G + R0 + rq;                    // + 8 * tl[mi]: every tile of a kriging panel is live
          const int64_t ldG = pb->ldG;
          double ssq[MI];
#pragma unroll
          for (int mi = 0; mi < MI; mi++) ssq[mi] = 0.0;
          for (int wb = 0; wb < q; wb += 16) {
            double g[MI][2][2], gold[MI][2][2];
#pragma unroll
            for (int mi = 0; mi < MI; mi++)
#pragma unroll
              for (int ni = 0; ni < 2; ni++)
#pragma unroll
                for (int e = 0; e < 2; e++) {
                  g[mi][ni][e] = 0.0;
                  const int col = wb + ni * 8 + kq * 2 + e;
                  gold[mi][ni][e] = (j == 0 || col >= q) ? 0.0 : __ldcg(Gr + tl[mi] * 8 + (int64_t)col * ldG);
                }
#pragma unroll 4
            for (int k4 = 0; k4 < TN / 4; k4++) {
              const int k = k4 * 4 + kq;
              const double* tk_ = T + (k >> 5) * 2 * UNIT + (k & 31) * UL;
              double a[MI];
#pragma unroll
              for (int mi = 0; mi < MI; mi++) a[mi] = tk_[ao[mi]];
              const double* wk = Ws + (k >> 5) * UNIT + (k & 31) * UL + wb + rq;
              const double b0 = wk[0], b1 = wk[8];
              if (wb == 0) {
#pragma unroll
                for (int mi = 0; mi < MI; mi++) ssq[mi] = fma(a[mi], a[mi], ssq[mi]);
              }
#pragma unroll
              for (int mi = 0; mi < MI; mi++) {
                dmma(g[mi][0][0], g[mi][0][1], a[mi], b0);
                dmma(g[mi][1][0], g[mi][1][1], a[mi], b1);
              }
            }
#pragma unroll
            for (int mi = 0; mi < MI; mi++)
#pragma unroll
              for (int ni = 0; ni < 2; ni++)
#pragma unroll
                for (int e = 0; e < 2; e++) {
                  const int col = wb + ni * 8 + kq * 2 + e;
                  if (col < q) Gr[tl[mi] * 8 + (int64_t)col * ldG] = gold[mi][ni][e] + g[mi][ni][e];
                }
          }
#pragma unroll
          for (int mi = 0; mi < MI; mi++) {                       // the four k-lanes of a row: fixed-order butterfly
            ssq[mi] += __shfl_xor_sync(0xffffffffu, ssq[mi], 1);
            ssq[mi] += __shfl_xor_sync(0xffffffffu, ssq[mi], 2);
          }
          if (kq == 0) {
#pragma unroll
            for (int mi = 0; mi < MI; mi++) {
              double* Gq = Gr + tl[mi] * 8 + (int64_t)q * ldG;
              *Gq = (j == 0 ? 0.0 : __ldcg(Gq)) + ssq[mi];
            }
          }
        }
      }
      // ---- E6: publish ----
      if (prof && tid == 0) tp4 = (long long)gtime();
      __threadfence();
      asm volatile("fence.proxy.async;" ::: "memory");     // the tile is re-read by other CTAs' TMA bulk copies
      pair_bar<NW>();
      if (tid == 0) st_release(pb->doneP + (int64_t)r * nCB + j, epoch);
      if (prof && tid == 0) {
        prof[0] = tp0; prof[1] = tp1; prof[2] = tp2; prof[3] = tp3; prof[4] = tp4; prof[5] = (long long)gtime();
        prof[6] = (nchunks > kstart) ? s_tdep : tp0;
        prof[7] = ((long long)blockIdx.x << 48) | (kwait << 1) | (is_diag ? 1 : 0);
      }
    }
    it = it0 + (nchunks - kstart);
    task_bar<NW>();
  }
}
