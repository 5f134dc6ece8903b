// Write-only roofline of the blocked factor layout (64 x 32 units, column stride 68 doubles): how fast can ANY kernel fill
// the lower triangle?  Variants: (0) plain contiguous stream, (1) blocked pattern, padding rows not written (what the
// covariance build does), (2) blocked pattern with the 4 padding rows written too (full 128-byte lines), (3) as 1 with
// evict-first stores.   nvcc -O3 -gencode arch=compute_100a,code=sm_100a store_pattern.cu -o store_pattern
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int UL = 68, UNIT = 32 * UL;
__device__ __forceinline__ void st_first(double* p, double a, double b, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1,%2}, %3;" ::"l"(p), "d"(a), "d"(b), "l"(pol) : "memory");
}
__global__ void __launch_bounds__(256) plain(double2* out, int64_t n2) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n2; i += (int64_t)gridDim.x * blockDim.x) out[i] = make_double2(1.0, 2.0);
}
template <int MODE>
__global__ void __launch_bounds__(256) blocked(double* L, int nRB, int nCB) {
  // tile (r, j): 128 rows x 64 cols; thread: 2 rows x 16 cols
  int j = blockIdx.y, r = (j >> 1) + blockIdx.x;
  if (r >= nRB) return;
  const int tid = threadIdx.x, rp = tid & 63, cg = tid >> 6, lr = rp * 2;
  uint64_t pol; asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  double* obase = L + (((int64_t)(2 * j + (cg >> 1))) * (nRB * 2) + 2 * r + (lr >> 6)) * UNIT + ((cg & 1) * 16) * UL + (lr & 63);
#pragma unroll
  for (int cc = 0; cc < 16; cc++) {
    if (MODE == 3) st_first(obase + cc * UL, 1.0 + tid, 2.0 + cc, pol);
    else *reinterpret_cast<double2*>(obase + cc * UL) = make_double2(1.0 + tid, 2.0 + cc);
    if (MODE == 2 && (lr & 63) == 62) {
      *reinterpret_cast<double2*>(obase + cc * UL + 2) = make_double2(0.0, 0.0);
      *reinterpret_cast<double2*>(obase + cc * UL + 4) = make_double2(0.0, 0.0);
    }
  }
}
int main(int argc, char** argv) {
  int n = argc > 1 ? atoi(argv[1]) : 20000;
  int Npad = (n + 63) / 64 * 64, ld = (Npad + 9 + 127) / 128 * 128, nCB = Npad / 64, nRB = ld / 128;
  size_t doubles = (size_t)(Npad / 32) * (ld / 64) * UNIT;
  double* L; cudaMalloc(&L, doubles * 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  double tri_bytes = 8.0 * n * (n + 1.0) / 2.0;
  int64_t tiles = 0; for (int c = 0; c < nCB; c++) tiles += nRB - c / 2;
  double tile_bytes = (double)tiles * 128 * 64 * 8;
  for (int mode = 0; mode < 4; mode++) {
    float best = 1e9;
    for (int rep = 0; rep < 6; rep++) {
      cudaEventRecord(e0);
      if (mode == 0) plain<<<148 * 8, 256>>>((double2*)L, (int64_t)(tile_bytes / 16));
      else if (mode == 1) blocked<1><<<dim3(nRB, nCB), 256>>>(L, nRB, nCB);
      else if (mode == 2) blocked<2><<<dim3(nRB, nCB), 256>>>(L, nRB, nCB);
      else blocked<3><<<dim3(nRB, nCB), 256>>>(L, nRB, nCB);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep > 0 && ms < best) best = ms;
    }
    printf("n=%d mode %d: %.3f ms  tile bytes %.0f MB -> %.0f GB/s written;  lower-triangle bytes / time = %.0f GB/s  (%s)\n", n, mode, best,
           tile_bytes / 1e6, tile_bytes / best / 1e6, tri_bytes / best / 1e6, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
