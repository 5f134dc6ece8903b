// Where do the cycles of a potf2 column step go?  Variants of potf2_step with parts switched off (timing only).
#include "../../iak3d_b200/csrc/chol.cu"

template <int CI, int F>
__device__ __forceinline__ void vstep(int cb, int br, int bc, bool live, double (&d)[4][4], double (&y)[4][4],
                                      double* __restrict__ colbuf, double* __restrict__ yrow, double* __restrict__ pivs, bool& bad) {
  const int c = 4 * cb + CI, buf = (CI & 1) * 64;
  if (bc == cb && br >= cb) {
#pragma unroll
    for (int i = 0; i < 4; i++) colbuf[buf + 4 * br + i] = d[i][CI];
  }
  if (br == cb && bc <= cb) {
#pragma unroll
    for (int jj = 0; jj < 4; jj++) yrow[buf + 4 * bc + jj] = y[CI][jj];
  }
  if (!(F & 1)) consumer_bar();
  if (F & 32) return;
  if (live && br >= cb) {
    const double piv = colbuf[buf + c];
    if (!(piv > 0.0) || isinf(piv)) bad = true;
    const double rs = (F & 2) ? 0.577 : fast_rsqrt(piv);
    if (br == cb && bc == cb) pivs[c] = piv;
    double li[4];
#pragma unroll
    for (int i = 0; i < 4; i++) li[i] = (4 * br + i > c) ? colbuf[buf + 4 * br + i] * rs : 0.0;
    if (bc >= cb && !(F & 4)) {
      double lj[4];
#pragma unroll
      for (int jj = 0; jj < 4; jj++) lj[jj] = (4 * bc + jj > c) ? colbuf[buf + 4 * bc + jj] * rs : 0.0;
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int jj = 0; jj < 4; jj++) d[i][jj] -= li[i] * lj[jj];
    }
    if (bc <= cb && !(F & 8)) {
      double yk[4];
#pragma unroll
      for (int jj = 0; jj < 4; jj++) yk[jj] = (4 * bc + jj <= c) ? yrow[buf + 4 * bc + jj] * rs : 0.0;
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int jj = 0; jj < 4; jj++) y[i][jj] -= li[i] * yk[jj];
      if (br == cb) {
#pragma unroll
        for (int jj = 0; jj < 4; jj++) y[CI][jj] = yk[jj];
      }
    }
    if (bc == cb && !(F & 16)) {
#pragma unroll
      for (int i = 0; i < 4; i++) { const int r = 4 * br + i; d[i][CI] = (r > c) ? li[i] : ((r == c) ? piv * rs : 0.0); }
    }
  }
}

template <int F>
__global__ void __launch_bounds__(NTHREADS, 1) vkernel(const double* A, double* out, long long* cyc, int reps) {
  extern __shared__ __align__(128) double smem[];
  __shared__ __align__(16) double s_comm[4 * 64 + 64];
  const int tid = threadIdx.x;
  if (tid >= NCW * 32) return;
  double* colbuf = s_comm; double* yrow = s_comm + 128; double* pivs = s_comm + 256;
  const int br = tid >> 4, bc = tid & 15; const bool live = bc <= br;
  long long total = 0; double sink = 0;
  for (int r = 0; r < reps; r++) {
    double d[4][4], y[4][4];
    for (int i = 0; i < 4; i++) for (int jj = 0; jj < 4; jj++) { d[i][jj] = live ? A[(4 * bc + jj) * 64 + 4 * br + i] : 0.0; y[i][jj] = (br == bc && i == jj) ? 1.0 : 0.0; }
    bool bad = false;
    consumer_bar();
    const long long t0 = clock64();
#pragma unroll 1
    for (int cb = 0; cb < 16; cb++) {
      vstep<0, F>(cb, br, bc, live, d, y, colbuf, yrow, pivs, bad); vstep<1, F>(cb, br, bc, live, d, y, colbuf, yrow, pivs, bad);
      vstep<2, F>(cb, br, bc, live, d, y, colbuf, yrow, pivs, bad); vstep<3, F>(cb, br, bc, live, d, y, colbuf, yrow, pivs, bad);
    }
    consumer_bar();
    total += clock64() - t0;
    for (int i = 0; i < 4; i++) for (int jj = 0; jj < 4; jj++) sink += d[i][jj] + y[i][jj];
    if (bad) sink += 1;
  }
  out[tid] = sink;
  if (tid == 0) *cyc = total / reps;
}

template <int F> void run(const char* name, const double* dA, double* dout, long long* dc) {
  cudaFuncSetAttribute(vkernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, 1024);
  vkernel<F><<<1, NTHREADS, 1024>>>(dA, dout, dc, 20);
  cudaError_t e = cudaDeviceSynchronize();
  long long c = 0; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
  printf("%-44s %8lld cycles  %7.1f per column  (%s)\n", name, c, c / 64.0, cudaGetErrorString(e));
}

int main() {
  const int n = 64;
  std::vector<double> A(n * n);
  for (int j = 0; j < n; j++) for (int i = 0; i < n; i++) A[j * n + i] = (i == j ? 3.0 : 0.0) + 1.0 / (1.0 + abs(i - j));
  double *dA, *dout; long long* dc;
  cudaMalloc(&dA, A.size() * 8); cudaMalloc(&dout, 4096 * 8); cudaMalloc(&dc, 8);
  cudaMemcpy(dA, A.data(), A.size() * 8, cudaMemcpyHostToDevice);
  run<0>("full", dA, dout, dc);
  run<2>("no rsqrt", dA, dout, dc);
  run<4>("no D update", dA, dout, dc);
  run<8>("no Y update", dA, dout, dc);
  run<12>("no D, no Y update", dA, dout, dc);
  run<14>("no D, no Y, no rsqrt", dA, dout, dc);
  run<30>("no D, no Y, no rsqrt, no finalise", dA, dout, dc);
  run<32>("publish + barrier only", dA, dout, dc);
  run<33>("publish only", dA, dout, dc);
  run<1>("full without barrier (wrong, timing only)", dA, dout, dc);
  return 0;
}
