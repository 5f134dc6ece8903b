// peer.cu -- the one exchange step of the composite likelihood (Reduce('+') over the workers, compositeLikIAK3D.R:63-64) as a
// kernel over NVLink peer memory instead of a library collective.  What is exchanged is tiny (q^2 + 2 doubles per parameter
// set: ~700 B), so the cost of a collective is its latency: NCCL's all-reduce of this buffer measured 106-270 us per
// evaluation on 2-8 B200s (profiles/scale_bench_N*_r02.json).  Here every rank owns a small cudaMalloc'ed mailbox that its
// peers map through CUDA IPC (one process per GPU); an all-reduce is ONE single-CTA kernel per rank:
//   1. copy the rank's values into its own mailbox (bank = call parity), release the bank's flag with the call number
//      (system scope: the readers are other GPUs);
//   2. acquire every peer's flag (bounded spin), then read the peers' banks through NVLink and add them IN RANK ORDER --
//      every rank forms the same sum bit for bit (the optimiser differentiates the result).
// Two banks suffice: a rank can only be two calls ahead of a reader if that reader has posted the call in between, and it
// posts only after it has finished reading the earlier one (kernels of one stream run in order).
#include "common.cuh"

constexpr int PEER_MAX_WORLD = 16;
constexpr unsigned long long PEER_TIMEOUT_NS = 20000000000ull;     // 20 s: ranks may arrive seconds apart

struct PeerBox {                       // layout of one rank's mailbox (device memory)
  unsigned long long flag[2];
  unsigned long long pad[14];
  // double data[2][cap] follows
};

struct iak3d_peer {
  iak3d_ctx* ctx = nullptr;
  int rank = 0, world = 1, cap = 0;
  unsigned long long calls = 0;
  void* own = nullptr;                 // this rank's mailbox
  void* box[PEER_MAX_WORLD] = {};      // every rank's mailbox as mapped here (box[rank] == own)
  void** d_box = nullptr;              // device copy of box[]
  int32_t* d_status = nullptr; int32_t* h_status = nullptr;
  double* h_stage = nullptr;           // pinned: the reduced values on their way to the caller
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool connected = false;
};

namespace {
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v; asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ double ld_relaxed_sys(const double* p) {
  double v; asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ unsigned long long gtime_ns() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }

__global__ void __launch_bounds__(128) peer_allreduce_kernel(void* const* __restrict__ boxes, int rank, int world, int cap, int n,
                                                             unsigned long long call, double* __restrict__ buf, int32_t* __restrict__ status) {
  const int bank = (int)(call & 1ull);
  PeerBox* mine = (PeerBox*)boxes[rank];
  double* mydata = (double*)(mine + 1) + (size_t)bank * cap;
  for (int i = threadIdx.x; i < n; i += blockDim.x) mydata[i] = buf[i];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) st_release_sys(&mine->flag[bank], call);
  __shared__ int s_bad;
  if (threadIdx.x == 0) s_bad = 0;
  __syncthreads();
  if (threadIdx.x < world && threadIdx.x != rank) {
    const unsigned long long* f = &((const PeerBox*)boxes[threadIdx.x])->flag[bank];
    const unsigned long long t0 = gtime_ns();
    while (ld_acquire_sys(f) < call) {
      __nanosleep(200);
      if (gtime_ns() - t0 > PEER_TIMEOUT_NS) { s_bad = 1; break; }
    }
  }
  __syncthreads();
  if (s_bad) { if (threadIdx.x == 0) atomicExch(status, 1); return; }
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double s = 0.0;
    for (int r = 0; r < world; r++) {
      const double* d = (const double*)((const PeerBox*)boxes[r] + 1) + (size_t)bank * cap;
      s += (r == rank) ? mydata[i] : ld_relaxed_sys(d + i);
    }
    buf[i] = s;
  }
}
}  // namespace

extern "C" int iak3d_peer_create(iak3d_ctx* ctx, int32_t rank, int32_t world, int32_t max_doubles, iak3d_peer** out, void* handle_out) {
  if (!ctx || !out || !handle_out || world < 1 || world > PEER_MAX_WORLD || rank < 0 || rank >= world || max_doubles < 1) return IAK3D_ERR_ARG;
  *out = nullptr;
  cudaSetDevice(ctx->device);
  iak3d_peer* p = new iak3d_peer();
  p->ctx = ctx; p->rank = rank; p->world = world; p->cap = max_doubles;
  const size_t bytes = sizeof(PeerBox) + (size_t)2 * max_doubles * sizeof(double);
  cudaError_t e = cudaMalloc(&p->own, bytes);                      // plain cudaMalloc: pooled allocations cannot be exported
  if (e == cudaSuccess) e = cudaMemset(p->own, 0, bytes);
  if (e == cudaSuccess) e = cudaMalloc(&p->d_box, PEER_MAX_WORLD * sizeof(void*));
  if (e == cudaSuccess) e = cudaMalloc(&p->d_status, sizeof(int32_t));
  if (e == cudaSuccess) e = cudaMemset(p->d_status, 0, sizeof(int32_t));
  if (e == cudaSuccess) e = cudaMallocHost(&p->h_status, sizeof(int32_t));
  if (e == cudaSuccess) e = cudaMallocHost(&p->h_stage, (size_t)max_doubles * sizeof(double));
  if (e == cudaSuccess) e = cudaEventCreate(&p->ev0);
  if (e == cudaSuccess) e = cudaEventCreate(&p->ev1);
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p->own);
  if (e != cudaSuccess) {
    ctx->err = std::string("peer_create: ") + cudaGetErrorString(e); cudaGetLastError();
    cudaFree(p->own); cudaFree(p->d_box); cudaFree(p->d_status); if (p->h_status) cudaFreeHost(p->h_status); if (p->h_stage) cudaFreeHost(p->h_stage);
    if (p->ev0) cudaEventDestroy(p->ev0); if (p->ev1) cudaEventDestroy(p->ev1); delete p;
    return IAK3D_ERR_CUDA;
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  memcpy(handle_out, &h, 64);
  p->box[rank] = p->own;
  *out = p;
  return IAK3D_OK;
}

extern "C" int iak3d_peer_connect(iak3d_peer* p, const void* handles) {
  if (!p || !handles) return IAK3D_ERR_ARG;
  iak3d_ctx* ctx = p->ctx;
  cudaSetDevice(ctx->device);
  for (int r = 0; r < p->world; r++) {
    if (r == p->rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char*)handles + (size_t)r * 64, 64);
    const cudaError_t e = cudaIpcOpenMemHandle(&p->box[r], h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { ctx->err = std::string("peer_connect: cudaIpcOpenMemHandle: ") + cudaGetErrorString(e); cudaGetLastError(); return IAK3D_ERR_CUDA; }
  }
  CUDA_TRY(ctx, cudaMemcpy(p->d_box, p->box, PEER_MAX_WORLD * sizeof(void*), cudaMemcpyHostToDevice));
  p->connected = true;
  return IAK3D_OK;
}

// In-place sum over the ranks of the n doubles at d_buf (a DEVICE pointer), enqueued on the context's stream; the stream is
// synchronised before returning and the sums are also written to h_out (host, may be NULL).  Every rank must call it the same
// number of times with the same n.  us (may be NULL) <- device time of the exchange kernel, the wait for the slowest rank included.
extern "C" int iak3d_peer_allreduce(iak3d_peer* p, double* d_buf, int32_t n, double* h_out, double* us) {
  if (!p || !d_buf || n < 0) return IAK3D_ERR_ARG;
  iak3d_ctx* ctx = p->ctx;
  if (!p->connected) { ctx->err = "peer_allreduce: not connected"; return IAK3D_ERR_ARG; }
  if (n > p->cap) { ctx->err = "peer_allreduce: more values than the mailbox holds"; return IAK3D_ERR_ARG; }
  cudaSetDevice(ctx->device);
  p->calls++;
  if (n > 0) {
    cudaEventRecord(p->ev0, ctx->stream);
    peer_allreduce_kernel<<<1, 128, 0, ctx->stream>>>(p->d_box, p->rank, p->world, p->cap, n, p->calls, d_buf, p->d_status);
    cudaEventRecord(p->ev1, ctx->stream);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    if (h_out) CUDA_TRY(ctx, cudaMemcpyAsync(p->h_stage, d_buf, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  }
  CUDA_TRY(ctx, cudaMemcpyAsync(p->h_status, p->d_status, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  if (*p->h_status != 0) { ctx->err = "peer_allreduce: a peer did not arrive within 20 s"; return IAK3D_ERR_TIMEOUT; }
  if (n > 0 && h_out) memcpy(h_out, p->h_stage, (size_t)n * sizeof(double));
  if (us) { float ms = 0.f; if (n > 0) cudaEventElapsedTime(&ms, p->ev0, p->ev1); *us = 1e3 * (double)ms; }
  return IAK3D_OK;
}

extern "C" int iak3d_peer_destroy(iak3d_peer* p) {
  if (!p) return IAK3D_OK;
  cudaSetDevice(p->ctx->device);
  cudaStreamSynchronize(p->ctx->stream);
  for (int r = 0; r < p->world; r++) if (r != p->rank && p->box[r]) cudaIpcCloseMemHandle(p->box[r]);
  cudaFree(p->own); cudaFree(p->d_box); cudaFree(p->d_status);
  if (p->h_status) cudaFreeHost(p->h_status);
  if (p->h_stage) cudaFreeHost(p->h_stage);
  cudaEventDestroy(p->ev0); cudaEventDestroy(p->ev1);
  delete p;
  return IAK3D_OK;
}
