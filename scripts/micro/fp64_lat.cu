// Dependent-issue latencies on B200 (cycles): DFMA, DADD, rsqrt (MUFU seed + 3 Newton), double SHFL, LDS, DMMA m8n8k4.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ double fast_rsqrt(double x) {
  double y; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double hx = 0.5 * x; double e = fma(-hx * y, y, 0.5); y = fma(y, e, y); e = fma(-hx * y, y, 0.5); y = fma(y, e, y); e = fma(-hx * y, y, 0.5); y = fma(y, e, y); return y;
}
__global__ void k(double* out, long long* cyc, int iters) {
  __shared__ double sm[64];
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0000001, c = 1e-9;
  sm[threadIdx.x & 63] = a; __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) { a = fma(a, b, c); a = fma(a, b, c); a = fma(a, b, c); a = fma(a, b, c); }
  long long t1 = clock64();
  for (int i = 0; i < iters; i++) { a = a + c; a = a + c; a = a + c; a = a + c; }
  long long t2 = clock64();
  for (int i = 0; i < iters; i++) { a = fast_rsqrt(a + 2.0); }
  long long t3 = clock64();
  for (int i = 0; i < iters; i++) { a = __shfl_sync(0xffffffffu, a, (threadIdx.x + 1) & 31); a = __shfl_sync(0xffffffffu, a, (threadIdx.x + 3) & 31); }
  long long t4 = clock64();
  int idx = threadIdx.x & 63;
  for (int i = 0; i < iters; i++) { double v = sm[idx]; idx = (int)(v * 0.0) + ((idx + 1) & 63); a += v; }
  long long t5 = clock64();
  double c0 = 0, c1 = 0;
  for (int i = 0; i < iters; i++) { asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b)); }
  long long t6 = clock64();
  double s4[4] = {a, a + 1, a + 2, a + 3};
  long long t7 = clock64();
  for (int i = 0; i < iters; i++) { s4[0] = fma(s4[0], b, c); s4[1] = fma(s4[1], b, c); s4[2] = fma(s4[2], b, c); s4[3] = fma(s4[3], b, c); }
  long long t8 = clock64();
  out[threadIdx.x] = a + c0 + c1 + s4[0] + s4[1] + s4[2] + s4[3];
  if (threadIdx.x == 0) { cyc[0] = (t1 - t0) / (4 * iters); cyc[1] = (t2 - t1) / (4 * iters); cyc[2] = (t3 - t2) / iters; cyc[3] = (t4 - t3) / (2 * iters); cyc[4] = (t5 - t4) / iters; cyc[5] = (t6 - t5) / iters; cyc[6] = (t8 - t7) / (4 * iters); }
}
int main() {
  double* o; long long* c; cudaMalloc(&o, 8 * 1024); cudaMalloc(&c, 64);
  for (int warps = 1; warps <= 8; warps *= 2) {
    k<<<1, 32 * warps>>>(o, c, 2000); cudaDeviceSynchronize();
    long long h[8]; cudaMemcpy(h, c, 56, cudaMemcpyDeviceToHost);
    printf("%d warp(s)/CTA: dependent DFMA %lld, DADD %lld, rsqrt(+add) %lld, SHFL.f64 %lld, LDS chain(+cvt) %lld, DMMA %lld, 4 independent DFMA chains: %lld per DFMA\n", warps, h[0], h[1], h[2], h[3], h[4], h[5], h[6]);
  }
  return 0;
}
