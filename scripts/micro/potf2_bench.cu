// Micro-benchmark of potf2_inv_blk4 (chol.cu) on one CTA: cycles per call, and a breakdown by disabling parts.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I. scripts/micro/potf2_bench.cu -o gpurun_out/potf2_bench
#ifndef NO_POTF2_PROF
#define POTF2_PROF 1
#endif
#ifndef POTF2_PROF_HELPER
#define POTF2_PROF_HELPER 64
#endif
#include "../../iak3d_b200/csrc/chol.cu"
cudaError_t iak_pool_alloc(iak3d_ctx*, void** p, size_t bytes) { return cudaMalloc(p, bytes); }   // chol.cu's host side is not used here
void iak_pool_free(iak3d_ctx*, void* p) { cudaFree(p); }

template <int V> __global__ void __launch_bounds__(NTHREADS, 1) bench_kernel(const double* A, double* out, long long* cyc, int reps) {
  extern __shared__ __align__(128) double smem[];
  double* At = smem + STAGES * STAGE_D;
  double* Ys = smem;
  __shared__ __align__(16) double s_comm[4 * 64 + 64];
  __shared__ int32_t status[2];
  const int tid = threadIdx.x;
  if (tid >= NCW * 32) return;
  long long total = 0;
  for (int r = 0; r < reps; r++) {
    for (int idx = tid; idx < 64 * 64; idx += NCW * 32) At[AT_IDX(idx & 63, idx >> 6)] = ((idx & 63) >= (idx >> 6)) ? A[idx] : nan("");
    consumer_bar();
    const long long t0 = clock64();
    double ls;
    if (V == 0) potf2_inv_blk4<false>(At, 0, Ys, smem + 2 * STAGE_D, s_comm + 256, tid, &ls, status);
    else potf2_inv_chain<(V >= 100 ? V - 100 : 0)>(At, 0, Ys, smem + 2 * STAGE_D, s_comm + 256, tid, &ls, status);
    consumer_bar();
    total += clock64() - t0;
    if (r == reps - 1 && tid == 0) out[64 * 64 * 2] = ls;
  }
  for (int idx = tid; idx < 64 * 64; idx += NCW * 32) { out[idx] = At[AT_IDX(idx & 63, idx >> 6)]; out[4096 + idx] = Ys[(idx >> 6) * LDB_S + (idx & 63)]; }
  if (tid == 0) *cyc = total / reps;
}

template <int V> void run(const char* name, const std::vector<double>& A, const double* dA, double* dout, long long* dc) {
  const int n = 64;
  std::vector<double> out(2 * n * n + 1);
  cudaMemset(dout, 0xff, out.size() * 8);
  cudaFuncSetAttribute(bench_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  bench_kernel<V><<<1, NTHREADS, SMEM_BYTES>>>(dA, dout, dc, 20);
  cudaError_t e = cudaDeviceSynchronize();
  long long c = 0; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
  cudaMemcpy(out.data(), dout, out.size() * 8, cudaMemcpyDeviceToHost);
  // check L L' = A, L Y = I, zeros above the diagonals, the log-determinant
  double e1 = 0, e2 = 0, e3 = 0, ld = 0;
  for (int i = 0; i < n; i++) for (int j = 0; j <= i; j++) {
    double s = 0, t = 0;
    for (int k = 0; k <= j; k++) s += out[k * n + i] * out[k * n + j];
    for (int k = j; k <= i; k++) t += out[k * n + i] * out[4096 + j * n + k];
    e1 = fmax(e1, fabs(s - A[j * n + i])); e2 = fmax(e2, fabs(t - (i == j ? 1.0 : 0.0)));
    if (j < i) e3 = fmax(e3, fmax(fabs(out[i * n + j]), fabs(out[4096 + i * n + j])));
    if (!(e3 == e3)) e3 = 1e300;
    if (i == j) ld += 2.0 * log(out[i * n + i]);
  }
  printf("%-22s %s: %6lld cycles per call (%.2f us at 1.965 GHz), %.1f per column; |LL'-A| %.2e |LY-I| %.2e upper %.1e logdet err %.2e\n", name,
         cudaGetErrorString(e), c, c / 1965.0, c / 64.0, e1, e2, e3, fabs(ld - out[2 * n * n]));
}

#ifdef POTF2_PROF
static void prof_reset() { long long z[16] = {0}; cudaMemcpyToSymbol(g_potf2_prof, z, sizeof(z)); }
static void prof_print() {
  long long h[16]; cudaMemcpyFromSymbol(h, g_potf2_prof, sizeof(h));
  printf("   chain warp, cycles per 4-column step: barrier %.0f | early update + factor %.0f | subst + stores %.0f | publish %.0f | late update %.0f\n",
         h[1] / 300.0, h[0] / 320.0, h[2] / 320.0, h[3] / 320.0, h[4] / 320.0);
  printf("   helper thread %d: wait for the panel %.0f | update (+ deposit) %.0f | arrive %.0f | copy / inverse row %.0f\n", POTF2_PROF_HELPER,
         h[5] / 320.0, h[6] / 320.0, h[7] / 320.0, h[8] / 320.0);
}
#else
static void prof_reset() {}
static void prof_print() {}
#endif

int main() {
  setvbuf(stdout, nullptr, _IONBF, 0);
  const int n = 64;
  std::vector<double> A(n * n);
  for (int j = 0; j < n; j++) for (int i = 0; i < n; i++) A[j * n + i] = (i == j ? 3.0 : 0.0) + 1.0 / (1.0 + abs(i - j)) + 0.3 * sin(0.37 * (i + 1) * (j + 1)) * sin(0.11 * (i - j));
  for (int j = 0; j < n; j++) for (int i = 0; i < j; i++) A[j * n + i] = A[i * n + j];
  for (int i = 0; i < n; i++) A[i * n + i] += 4.0;
  double *dA, *dout; long long* dc;
  cudaMalloc(&dA, A.size() * 8); cudaMalloc(&dout, (2 * n * n + 1) * 8); cudaMalloc(&dc, 8);
  cudaMemcpy(dA, A.data(), A.size() * 8, cudaMemcpyHostToDevice);
  run<0>("potf2_inv_blk4", A, dA, dout, dc);
  prof_reset(); run<2>("potf2_inv_chain", A, dA, dout, dc); prof_print();
  prof_reset(); run<148>("chain: helpers idle", A, dA, dout, dc); prof_print();
  run<101>("chain: no panel stores", A, dA, dout, dc);
  run<102>("chain: no lr publish", A, dA, dout, dc);
  run<104>("chain: no dn update", A, dA, dout, dc);
  run<108>("chain: no nx update", A, dA, dout, dc);
  run<116>("chain: helpers no updates", A, dA, dout, dc);
  run<132>("chain: helpers no inverse", A, dA, dout, dc);
  return 0;
}
