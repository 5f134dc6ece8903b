// Micro-benchmark of potf2_inv_blk4 (chol.cu) on one CTA: cycles per call, and a breakdown by disabling parts.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I. scripts/micro/potf2_bench.cu -o gpurun_out/potf2_bench
#include "../../iak3d_b200/csrc/chol.cu"

__global__ void __launch_bounds__(NTHREADS, 1) bench_kernel(const double* A, double* out, long long* cyc, int reps) {
  extern __shared__ __align__(128) double smem[];
  double* At = smem + STAGES * STAGE_D;
  double* Ys = smem;
  __shared__ __align__(16) double s_comm[4 * 64 + 64];
  __shared__ int32_t status[2];
  const int tid = threadIdx.x;
  if (tid >= NCW * 32) return;
  long long total = 0;
  for (int r = 0; r < reps; r++) {
    for (int idx = tid; idx < 64 * 64; idx += NCW * 32) At[AT_IDX(idx & 63, idx >> 6)] = A[idx];
    consumer_bar();
    const long long t0 = clock64();
    double ls;
    potf2_inv_blk4(At, 0, Ys, smem + 2 * STAGE_D, s_comm + 256, tid, &ls, status);
    consumer_bar();
    total += clock64() - t0;
    if (r == reps - 1 && tid == 0) out[64 * 64 * 2] = ls;
  }
  for (int idx = tid; idx < 64 * 64; idx += NCW * 32) { out[idx] = At[AT_IDX(idx & 63, idx >> 6)]; out[4096 + idx] = Ys[(idx >> 6) * LDB_S + (idx & 63)]; }
  if (tid == 0) *cyc = total / reps;
}

int main() {
  const int n = 64;
  std::vector<double> A(n * n), out(2 * n * n + 1);
  for (int j = 0; j < n; j++) for (int i = 0; i < n; i++) A[j * n + i] = (i == j ? 3.0 : 0.0) + 1.0 / (1.0 + abs(i - j));
  double *dA, *dout; long long* dc;
  cudaMalloc(&dA, A.size() * 8); cudaMalloc(&dout, out.size() * 8); cudaMalloc(&dc, 8);
  cudaMemcpy(dA, A.data(), A.size() * 8, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  bench_kernel<<<1, NTHREADS, SMEM_BYTES>>>(dA, dout, dc, 20);
  cudaError_t e = cudaDeviceSynchronize();
  long long c = 0; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
  cudaMemcpy(out.data(), dout, out.size() * 8, cudaMemcpyDeviceToHost);
  // check L L' = A and L Y = I
  double e1 = 0, e2 = 0;
  for (int i = 0; i < n; i++) for (int j = 0; j <= i; j++) {
    double s = 0, t = 0;
    for (int k = 0; k <= j; k++) s += out[k * n + i] * out[k * n + j];
    for (int k = j; k <= i; k++) t += out[k * n + i] * out[4096 + j * n + k];
    e1 = fmax(e1, fabs(s - A[j * n + i])); e2 = fmax(e2, fabs(t - (i == j ? 1.0 : 0.0)));
  }
  printf("%s: potf2_inv %lld cycles per call (%.2f us at 1.965 GHz), %.1f cycles per column; |LL'-A| %.2e |LY-I| %.2e\n",
         cudaGetErrorString(e), c, c / 1965.0, c / 64.0, e1, e2);
  return 0;
}
