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
  const double y0 = y;
  for (int s = 1; s <= 3; s++) { double e = fma(-hx * y, y, 0.5); y = fma(y, e, y); out[s * n + i] = y; }
  // one third-order step from the seed: y (1 + e/2 + 3 e^2 / 8), e = 1 - x y^2   (four dependent operations instead of six)
  { const double t = x[i] * y0; const double e = fma(-t, y0, 1.0); const double w = fma(e, 0.375, 0.5); const double ye = y0 * e; out[4 * n + i] = fma(ye, w, y0); }
}
int main() {
  const int n = 1 << 20;
  std::vector<double> x(n), out(5 * n);
  unsigned long long st = 88172645463325252ull;
  for (int i = 0; i < n; i++) { st ^= st << 13; st ^= st >> 7; st ^= st << 17; double u = (st >> 11) * (1.0 / 9007199254740992.0); x[i] = exp((u - 0.5) * 60.0); }
  double *dx, *dout; cudaMalloc(&dx, n * 8); cudaMalloc(&dout, 5 * n * 8);
  cudaMemcpy(dx, x.data(), n * 8, cudaMemcpyHostToDevice);
  k<<<n / 256, 256>>>(dx, dout, n);
  cudaMemcpy(out.data(), dout, 5 * n * 8, cudaMemcpyDeviceToHost);
  for (int s = 0; s < 5; s++) {
    long double worst = 0;
    for (int i = 0; i < n; i++) { long double t = 1.0L / sqrtl((long double)x[i]); long double e = fabsl((long double)out[s * n + i] - t) / t; if (e > worst) worst = e; }
    printf("%s%d: max relative error %.3Le = %.2Lf ulp(2^-53)\n", s < 4 ? "Newton steps " : "one third-order step, variant ", s < 4 ? s : 1, worst, worst / 1.1102230246251565e-16L);
  }
  return 0;
}
