// Accuracy of the MUFU seed and of 1, 2, 3 Newton steps of fast_rsqrt (chol.cu) against 1/sqrt in long double on the host.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 scripts/micro/rsqrt_acc.cu -o scripts/micro/rsqrt_acc
#include <cstdio>
#include <cmath>
#include <vector>
__global__ void k(const double* x, double* out, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double y; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x[i]));
  out[i] = y;
  const double hx = 0.5 * x[i];
  for (int s = 1; s <= 3; s++) { double e = fma(-hx * y, y, 0.5); y = fma(y, e, y); out[s * n + i] = y; }
}
int main() {
  const int n = 1 << 20;
  std::vector<double> x(n), out(4 * n);
  unsigned long long st = 88172645463325252ull;
  for (int i = 0; i < n; i++) { st ^= st << 13; st ^= st >> 7; st ^= st << 17; double u = (st >> 11) * (1.0 / 9007199254740992.0); x[i] = exp((u - 0.5) * 60.0); }
  double *dx, *dout; cudaMalloc(&dx, n * 8); cudaMalloc(&dout, 4 * n * 8);
  cudaMemcpy(dx, x.data(), n * 8, cudaMemcpyHostToDevice);
  k<<<n / 256, 256>>>(dx, dout, n);
  cudaMemcpy(out.data(), dout, 4 * n * 8, cudaMemcpyDeviceToHost);
  for (int s = 0; s < 4; s++) {
    long double worst = 0;
    for (int i = 0; i < n; i++) { long double t = 1.0L / sqrtl((long double)x[i]); long double e = fabsl((long double)out[s * n + i] - t) / t; if (e > worst) worst = e; }
    printf("%d Newton steps: max relative error %.3Le = %.2Lf ulp(2^-53)\n", s, worst, worst / 1.1102230246251565e-16L);
  }
  return 0;
}
