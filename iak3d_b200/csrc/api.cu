// api.cu -- the extern "C" boundary of libiak3d.so (include/iak3d.h): contexts, datasets (setupIAK3D as id
// vectors), evaluation planning from parsBTfmd, and the likelihood-term pipeline
//     tables -> covariance build into the factor layout -> ticketed tile Cholesky -> finish.
// Host logic only; all arithmetic on matrices happens in the kernels of tables.cu / covbuild.cu / chol.cu.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <thread>
#include <unordered_map>

#include "common.cuh"

// ================================================================================================
// helpers
// ================================================================================================
namespace {

struct Key2 {
  uint64_t a, b;
  bool operator==(const Key2& o) const { return a == o.a && b == o.b; }
};
struct Key2Hash {
  size_t operator()(const Key2& k) const {
    uint64_t h = k.a * 0x9E3779B97F4A7C15ull;
    h ^= (h >> 29);
    h += k.b * 0xC2B2AE3D27D4EB4Full;
    h ^= (h >> 32);
    return (size_t)h;
  }
};
inline uint64_t dbits(double v) {
  if (v == 0.0) v = 0.0;   // -0.0 == 0.0 in R's `==`
  uint64_t u;
  memcpy(&u, &v, 8);
  return u;
}

}  // namespace

// rows (c0[i], c1[i]) -> ids in first-occurrence order (`!duplicated`, iaCovMatern.R:418-419) + unique table
void iak_unique_first(int64_t n, const double* c0, const double* c1, std::vector<int32_t>& id, std::vector<double>& U) {
  std::unordered_map<Key2, int32_t, Key2Hash> map;
  map.reserve((size_t)std::min<int64_t>(n, 1 << 22));
  std::vector<double> u0, u1;
  id.resize((size_t)n);
  Key2 prev{0, 0}; int32_t prev_id = -1;               // runs of equal keys (a prediction grid is ordered interval by interval)
  for (int64_t i = 0; i < n; i++) {
    const Key2 k{dbits(c0[i]), dbits(c1[i])};
    if (prev_id >= 0 && k.a == prev.a && k.b == prev.b) { id[(size_t)i] = prev_id; continue; }
    prev = k;
    auto it = map.find(k);
    if (it == map.end()) {
      const int32_t v = (int32_t)u0.size();
      map.emplace(k, v);
      u0.push_back(c0[i]); u1.push_back(c1[i]);
      id[(size_t)i] = v;
    } else {
      id[(size_t)i] = it->second;
    }
    prev_id = id[(size_t)i];
  }
  const size_t nu = u0.size();
  U.resize(2 * nu);
  for (size_t i = 0; i < nu; i++) { U[i] = u0[i]; U[nu + i] = u1[i]; }
}

// parsBTfmd -> plan.  `square`: setCIAK3D (the cd1 quirk of fitIAK3D.R:1052 applies); otherwise setCIAK3D2 (:1211).
// Returns <0 for unsupported requests (spline sdfd), otherwise 0 with plan.status in {0, IAK3D_BAD_PARS}.
int iak_make_plan(iak3d_ctx* ctx, const iak3d_pars* p, bool square, EvalPlan* pl) {
  *pl = EvalPlan();
  for (int c = 0; c < 3; c++) {
    const int t = p->sdfdType[c];
    if (t > IAK3D_MAX_SPL) { ctx->err = "sdfdType: more internal knots than IAK3D_MAX_SPL"; return IAK3D_ERR_ARG; }
    if (t > 0) pl->approx = true;                                     // fitIAK3D.R:959-962 / 1121-1124
    if (!(t >= -1 || (t == -9 && c == 0))) { ctx->err = "sdfdType must be 0, -1, k > 0 (or -9 for cd1)"; return IAK3D_ERR_ARG; }
  }
  if (p->modelx < IAK3D_MODELX_NUGGET || p->modelx > IAK3D_MODELX_WENDLAND) { ctx->err = "unknown modelx"; return IAK3D_ERR_ARG; }
  const bool nugget = (p->modelx == IAK3D_MODELX_NUGGET);
  pl->cx0 = p->cx0; pl->cx1 = nugget ? 0.0 : p->cx1; pl->cd1 = p->cd1; pl->cxd0 = p->cxd0; pl->cxd1 = nugget ? 0.0 : p->cxd1;
  pl->cme = p->cme;
  if (p->nssre < 0 || p->nssre > IAK3D_MAX_SSRE) { ctx->err = "nssre out of range (0 .. IAK3D_MAX_SSRE)"; return IAK3D_ERR_ARG; }
  pl->nssre = p->nssre;
  for (int k = 0; k < p->nssre; k++) { pl->cssre[k] = p->cssre[k]; if (!isfinite(p->cssre[k])) { pl->status = IAK3D_BAD_PARS; return 0; } }
  // NA / Inf guards of nllIAK3D (fitIAK3D.R:718-741)
  const double vs[6] = {pl->cx0, pl->cx1, pl->cd1, pl->cxd0, pl->cxd1, pl->cme};
  for (double v : vs) if (!isfinite(v)) { pl->status = IAK3D_BAD_PARS; return 0; }
  if (!isfinite(p->ad) || !(p->ad > 0.0)) { pl->status = IAK3D_BAD_PARS; return 0; }
  for (int c = 0; c < 3; c++) {
    if (p->sdfdType[c] == -1 && !(isfinite(p->sdfdPars[c][0]) && isfinite(p->sdfdPars[c][1]))) { pl->status = IAK3D_BAD_PARS; return 0; }
    for (int k = 0; k < p->sdfdType[c]; k++) if (!isfinite(p->sdfdSpl[c][k])) { pl->status = IAK3D_BAD_PARS; return 0; }
  }
  // depth tables: one shared stationary table (fitIAK3D.R:986-993) + one per non-stationary component.  With the
  // approximation every table starts from the stationary one and is scaled by its component's averaged sd (:1960).
  int stat = -1;
  for (int c = 0; c < 3; c++) {
    const int t = p->sdfdType[c];
    if (c == 2 && nugget) { pl->tidx[c] = -1; continue; }           // :1018-1020
    if (t == -9) { pl->tidx[c] = -1; continue; }                     // :998-1000
    if (t == 0) {
      if (stat < 0) {
        stat = pl->ntab++;
        if (iak_prm_init(&pl->prm[stat], p->ad, p->nud, 0, 0.0, 0.0) != 0) { pl->status = IAK3D_BAD_PARS; return 0; }
      }
      pl->tidx[c] = stat;
    } else if (!pl->approx) {
      const int k = pl->ntab++;
      if (iak_prm_init(&pl->prm[k], p->ad, p->nud, -1, p->sdfdPars[c][0], p->sdfdPars[c][1]) != 0) { pl->status = IAK3D_BAD_PARS; return 0; }
      pl->tidx[c] = k;
    } else {
      const int k = pl->ntab++;
      if (iak_prm_init(&pl->prm[k], p->ad, p->nud, 0, 0.0, 0.0) != 0) { pl->status = IAK3D_BAD_PARS; return 0; }
      pl->sd_type[k] = t; pl->sd_comp[k] = c;
      if (t == -1) { pl->sd_coef[k][0] = p->sdfdPars[c][0]; pl->sd_coef[k][1] = p->sdfdPars[c][1]; }
      else for (int e = 0; e < t; e++) pl->sd_coef[k][e] = p->sdfdSpl[c][e];
      pl->tidx[c] = k;
    }
  }
  pl->kd1 = pl->cd1;
  if (square && (p->sdfdType[0] == -1 || p->sdfdType[0] > 0) && p->quirk_cd1_unscaled)   // fitIAK3D.R:1052, 2026 (sic)
    pl->kd1 = (p->quirk_scale != 0.0) ? p->quirk_scale : 1.0;
  pl->has_phix = !nugget;
  if (!nugget) {
    const int st = hx_init(&pl->H, p->modelx, p->ax, p->nux);
    if (st != 0) { pl->status = IAK3D_BAD_PARS; return 0; }
  }
  return 0;
}

// iasdfd (fitIAK3D.R:1448-1468) of table k on a set of unique intervals (host: ndIU values).  XsplU: nd x ncol spline
// basis of the table's component on that set (type > 0).  Returns IAK3D_BAD_PARS when min(avsd) < 0 (:1994, 2172).
int iak_avsd_host(iak3d_ctx* ctx, const EvalPlan& pl, int k, const double* dIU, int64_t nd, const double* XsplU, int ncol,
                  std::vector<double>& s) {
  const int t = pl.sd_type[k];
  s.assign((size_t)nd, 1.0);
  if (t == -1) {
    const double t1 = pl.sd_coef[k][0], t2 = pl.sd_coef[k][1];
    for (int64_t i = 0; i < nd; i++) {
      const double dU = dIU[i], dL = dIU[nd + i];
      s[(size_t)i] = (1.0 - t1) - (t1 / t2) * (exp(-t2 * dL) - exp(-t2 * dU)) / (dL - dU);
    }
  } else if (t > 0) {
    if (XsplU == nullptr || ncol != t + 1) {
      ctx->err = "spline sd function: the basis of this component is missing or has the wrong number of columns "
                 "(iak3d_dataset_set_sdfd_spline / Xspl1; ncol must be sdfdType + 1)";
      return IAK3D_ERR_ARG;
    }
    for (int64_t i = 0; i < nd; i++) {
      double v = XsplU[i];                                            // c(1, sdfdPars): first column times 1
      for (int e = 0; e < t; e++) v += XsplU[(size_t)(e + 1) * nd + i] * pl.sd_coef[k][e];
      s[(size_t)i] = v;
    }
  }
  for (int64_t i = 0; i < nd; i++) if (!(s[(size_t)i] >= 0.0) || !isfinite(s[(size_t)i])) return IAK3D_BAD_PARS;
  return 0;
}

// parsOK of setCApproxIAK3D (fitIAK3D.R:1993-1994) decided at planning time: a negative averaged sd -> C = NA
int iak_plan_check_avsd(iak3d_ctx* ctx, const iak3d_ds* ds, EvalPlan* pl) {
  if (!pl->approx || pl->status != 0) return 0;
  std::vector<double> sv;
  for (int k = 0; k < pl->ntab; k++) {
    if (pl->sd_type[k] == 0) continue;
    const int c = pl->sd_comp[k];
    const int st = iak_avsd_host(ctx, *pl, k, ds->h_dIU.data(), ds->ndIU, ds->h_spl[c].empty() ? nullptr : ds->h_spl[c].data(),
                                 ds->spl_ncol[c], sv);
    if (st == IAK3D_BAD_PARS) { pl->status = IAK3D_BAD_PARS; return 0; }
    if (st != 0) return st;
  }
  return 0;
}

// ================================================================================================
// context
// ================================================================================================
struct Batch {
  int32_t count = 0, q = 0;
  std::vector<iak3d_ds*> ds;
  std::vector<Slot*> slots;
  std::vector<int> plan_status;
  std::vector<int64_t> key;          // geometry key of the cached task list
  ProblemDesc* d_probs = nullptr; int32_t probs_cap = 0;
  int4* d_tasks = nullptr; int64_t tasks_cap = 0; int32_t ntasks = 0;
  int32_t* d_ticket = nullptr;
  int32_t* d_status = nullptr; int32_t status_cap = 0;
  double* d_results = nullptr; double* h_results = nullptr; int64_t results_cap = 0;
  int32_t stride = 0;
  bool pending = false;
};

bool iak3d_ctx::batch_pending() const { return batch != nullptr && batch->pending; }

extern "C" int iak3d_ctx_create(int device, iak3d_ctx** out) {
  if (out == nullptr) return IAK3D_ERR_ARG;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) return IAK3D_ERR_CUDA;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return IAK3D_ERR_CUDA;
  if (prop.major != 10) return IAK3D_ERR_CUDA;   // sm_100-class only: no fallback of any kind
  if (cudaSetDevice(device) != cudaSuccess) return IAK3D_ERR_CUDA;
  iak3d_ctx* ctx = new iak3d_ctx();
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  ctx->cc_major = prop.major; ctx->cc_minor = prop.minor;
  ctx->total_mem = prop.totalGlobalMem;
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return IAK3D_ERR_CUDA; }
  cudaEventCreate(&ctx->ev0); cudaEventCreate(&ctx->ev1);
  for (int i = 0; i < 5; i++) cudaEventCreate(&ctx->evs[i]);
  ctx->batch = new Batch();
  if (cudaMalloc(&ctx->batch->d_ticket, 2 * sizeof(int32_t)) != cudaSuccess) { delete ctx->batch; delete ctx; return IAK3D_ERR_CUDA; }
  if (chol_configure(ctx) != 0) { delete ctx->batch; delete ctx; return IAK3D_ERR_CUDA; }
  *out = ctx;
  return IAK3D_OK;
}

extern "C" int iak3d_ctx_destroy(iak3d_ctx* ctx) {
  if (ctx == nullptr) return IAK3D_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  Batch* b = ctx->batch;
  if (b != nullptr) {
    cudaFree(b->d_probs); cudaFree(b->d_tasks); cudaFree(b->d_ticket); cudaFree(b->d_status); cudaFree(b->d_results);
    if (b->h_results) cudaFreeHost(b->h_results);
    delete b;
  }
  if (ctx->flush_buf) cudaFree(ctx->flush_buf);
  cudaFree(ctx->d_fbjobs); cudaFree(ctx->d_tabjobs); if (ctx->h_tabjobs) cudaFreeHost(ctx->h_tabjobs);
  cudaFree(ctx->d_prof);
  cudaEventDestroy(ctx->ev0); cudaEventDestroy(ctx->ev1);
  for (int i = 0; i < 5; i++) cudaEventDestroy(ctx->evs[i]);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
  return IAK3D_OK;
}

cudaError_t iak_pool_alloc(iak3d_ctx* ctx, void** p, size_t bytes) {
  static int configured_device = -1;
  if (configured_device != ctx->device) {
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, ctx->device) == cudaSuccess) {
      uint64_t keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    configured_device = ctx->device;
  }
  return cudaMallocAsync(p, std::max<size_t>(bytes, 16), ctx->stream);
}
void iak_pool_free(iak3d_ctx* ctx, void* p) {
  if (p) cudaFreeAsync(p, ctx->stream);
}
cudaError_t iak_malloc(iak3d_ctx* ctx, void** p, size_t bytes) {
  cudaError_t e = cudaMalloc(p, bytes);
  if (e == cudaErrorMemoryAllocation) {
    cudaGetLastError();
    cudaStreamSynchronize(ctx->stream);
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, ctx->device) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
    e = cudaMalloc(p, bytes);
  }
  return e;
}

extern "C" const char* iak3d_last_error(iak3d_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

extern "C" int iak3d_device_info(iak3d_ctx* ctx, int* sm_count, int* cc_major, int* cc_minor, int64_t* total_mem) {
  if (!ctx) return IAK3D_ERR_ARG;
  if (sm_count) *sm_count = ctx->sm_count;
  if (cc_major) *cc_major = ctx->cc_major;
  if (cc_minor) *cc_minor = ctx->cc_minor;
  if (total_mem) *total_mem = (int64_t)ctx->total_mem;
  return IAK3D_OK;
}

extern "C" int64_t iak3d_launch_count(iak3d_ctx* ctx) { return ctx ? ctx->launches : -1; }

// ================================================================================================
// dataset
// ================================================================================================
static void free_slot(Slot* s) {
  if (!s) return;
  cudaFree(s->d_L); cudaFree(s->d_phix); cudaFree(s->d_phid); cudaFree(s->d_invd); cudaFree(s->d_flags);
  cudaFree(s->d_logsum); cudaFree(s->d_G); cudaFree(s->d_lncol); cudaFree(s->d_E);
  delete s;
}

// factor-buffer workspace for an n x n problem with q bordered rows (tables optional: nxU = ndIU = 0)
// tables_only: no factor buffer (the perturbed parameter sets of the analytic gradient need tables, not matrices)
int iak_alloc_slot(iak3d_ctx* ctx, int64_t n, int32_t q, int64_t nxU, int64_t ndIU, Slot** out, bool tables_only) {
  Slot* s = new Slot();
  s->n = n; s->q = q; s->nxU = nxU; s->ndIU = ndIU;
  s->Npad = round_up(std::max<int64_t>(n, 1), TN);
  s->ld = round_up(s->Npad + q, TM);
  s->nCB = (int32_t)(s->Npad / TN); s->nRB = (int32_t)(s->ld / TM);
  cudaError_t e = cudaSuccess;
  auto A = [&](void** p, size_t bytes) { if (e == cudaSuccess) e = iak_malloc(ctx, p, std::max<size_t>(bytes, 16)); };
  if (!tables_only) A((void**)&s->d_L, (size_t)ubuf_doubles(s->ld, s->Npad) * sizeof(double));
  A((void**)&s->d_phix, (size_t)nxU * nxU * sizeof(double));
  // tables | avVar vectors | avsd vectors | (16-byte aligned) combined tables of the factor build
  A((void**)&s->d_phid, (size_t)(NTAB_BUF * ndIU * ndIU + round_up((NTAB_BUF + 3) * ndIU, 2) + 3 * ndIU * ndIU) * sizeof(double));
  const size_t nflags = (size_t)s->nRB * s->nCB + s->nCB;
  if (!tables_only) {
    A((void**)&s->d_invd, (size_t)s->nCB * INVD_BLOCK * sizeof(double));
    A((void**)&s->d_flags, nflags * sizeof(int32_t));
    A((void**)&s->d_logsum, (size_t)s->nCB * sizeof(double));
    A((void**)&s->d_G, 64 * 64 * sizeof(double));
    if (ndIU > 0) A((void**)&s->d_E, (size_t)2 * ndIU * round_up(n, 16) * sizeof(double));
    if (e == cudaSuccess) e = cudaMemsetAsync(s->d_flags, 0, nflags * sizeof(int32_t), ctx->stream);
  }
  if (e != cudaSuccess) {
    ctx->err = std::string("slot allocation failed: ") + cudaGetErrorString(e);
    free_slot(s);
    cudaGetLastError();
    return IAK3D_ERR_CUDA;
  }
  *out = s;
  return 0;
}
void iak_free_slot(Slot* s) { free_slot(s); }

static void ds_geometry(iak3d_ds* ds) {
  ds->Npad = round_up(std::max<int64_t>(ds->n, 1), TN);
  ds->ld = round_up(ds->Npad + ds->q, TM);
  ds->nCB = (int32_t)(ds->Npad / TN); ds->nRB = (int32_t)(ds->ld / TM);
}

extern "C" int iak3d_dataset_set_W(iak3d_ds* ds, const double* W, int32_t q) {
  if (!ds) return IAK3D_ERR_ARG;
  iak3d_ctx* ctx = ds->ctx;
  if (q < 0 || q > 64 || (q > 0 && W == nullptr)) { ctx->err = "set_W: q must be in [0,64] (p <= 63 fixed effects)"; return IAK3D_ERR_ARG; }
  cudaSetDevice(ctx->device);
  if (q != ds->q) {
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    for (Slot* s : ds->slots) free_slot(s);
    ds->slots.clear();
    cudaFree(ds->d_W); ds->d_W = nullptr;
    if (ds->h_W) { cudaFreeHost(ds->h_W); ds->h_W = nullptr; }
    ds->q = q;
    ds_geometry(ds);
    if (q > 0) {
      CUDA_TRY(ctx, cudaMalloc(&ds->d_W, (size_t)ds->n * q * sizeof(double)));
      CUDA_TRY(ctx, cudaMallocHost(&ds->h_W, (size_t)ds->n * q * sizeof(double)));
    }
  }
  if (q > 0) {
    // caller's (pageable, R-owned) buffer -> pinned staging -> device; the staging copy doubles as the host copy.
    // A previous asynchronous copy out of the staging buffer must have finished before it is overwritten.
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    memcpy(ds->h_W, W, (size_t)ds->n * q * sizeof(double));
    CUDA_TRY(ctx, cudaMemcpyAsync(ds->d_W, ds->h_W, (size_t)ds->n * q * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    ds->w_set = true;
  }
  return IAK3D_OK;
}

// W unchanged since the last iak3d_dataset_set_W?  (optim() passes the same XData / zData to every callback: the R glue asks
// before it uploads n x q doubles again; bit-wise comparison against the pinned host copy.)
extern "C" int iak3d_dataset_same_W(iak3d_ds* ds, const double* W, int32_t q, int32_t* same) {
  if (!ds || !same) return IAK3D_ERR_ARG;
  *same = (q == ds->q && q > 0 && W != nullptr && ds->h_W != nullptr && ds->w_set &&
           memcmp(ds->h_W, W, (size_t)ds->n * q * sizeof(double)) == 0) ? 1 : 0;
  return IAK3D_OK;
}

extern "C" int iak3d_dataset_create(iak3d_ctx* ctx, int64_t n, const double* xy, const double* dI, const double* W, int32_t q,
                                    iak3d_ds** out) {
  if (!ctx || !out) return IAK3D_ERR_ARG;
  *out = nullptr;
  if (n <= 0 || xy == nullptr || dI == nullptr) { ctx->err = "dataset_create: n > 0, xy and dI are required"; return IAK3D_ERR_ARG; }
  if (n > (int64_t)1 << 30) { ctx->err = "dataset_create: n too large"; return IAK3D_ERR_ARG; }
  cudaSetDevice(ctx->device);
  iak3d_ds* ds = new iak3d_ds();
  ds->ctx = ctx; ds->n = n; ds->q = -1;
  iak_unique_first(n, xy, xy + n, ds->h_xid, ds->h_xU);
  iak_unique_first(n, dI, dI + n, ds->h_did, ds->h_dIU);
  ds->nxU = (int64_t)ds->h_xU.size() / 2; ds->ndIU = (int64_t)ds->h_dIU.size() / 2;
  cudaError_t e = cudaSuccess;
  auto up = [&](void** p, const void* src, size_t bytes) {
    if (e == cudaSuccess) e = cudaMalloc(p, bytes);
    if (e == cudaSuccess) e = cudaMemcpyAsync(*p, src, bytes, cudaMemcpyHostToDevice, ctx->stream);
  };
  up((void**)&ds->d_xid, ds->h_xid.data(), (size_t)n * 4);
  up((void**)&ds->d_did, ds->h_did.data(), (size_t)n * 4);
  up((void**)&ds->d_xU, ds->h_xU.data(), ds->h_xU.size() * 8);
  up((void**)&ds->d_dIU, ds->h_dIU.data(), ds->h_dIU.size() * 8);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) {
    ctx->err = std::string("dataset_create: ") + cudaGetErrorString(e);
    iak3d_dataset_destroy(ds);
    return IAK3D_ERR_CUDA;
  }
  const int st = iak3d_dataset_set_W(ds, W, W ? q : 0);
  if (st != 0) { iak3d_dataset_destroy(ds); return st; }
  *out = ds;
  return IAK3D_OK;
}

extern "C" int iak3d_dataset_destroy(iak3d_ds* ds) {
  if (!ds) return IAK3D_OK;
  cudaSetDevice(ds->ctx->device);
  cudaStreamSynchronize(ds->ctx->stream);
  for (Slot* s : ds->slots) free_slot(s);
  for (Slot* s : ds->tslots) free_slot(s);
  cudaFree(ds->d_gradU); cudaFree(ds->d_gradiC);
  cudaFree(ds->d_xid); cudaFree(ds->d_did); cudaFree(ds->d_xU); cudaFree(ds->d_dIU); cudaFree(ds->d_W);
  cudaFree(ds->d_sid); cudaFree(ds->d_Xs);
  if (ds->h_W) cudaFreeHost(ds->h_W);
  delete ds;
  return IAK3D_OK;
}

extern "C" int iak3d_dataset_info(iak3d_ds* ds, int64_t* n, int64_t* nxU, int64_t* ndIU, int32_t* q) {
  if (!ds) return IAK3D_ERR_ARG;
  if (n) *n = ds->n;
  if (nxU) *nxU = ds->nxU;
  if (ndIU) *ndIU = ds->ndIU;
  if (q) *q = ds->q;
  return IAK3D_OK;
}

extern "C" int iak3d_dataset_ids(iak3d_ds* ds, int32_t* xid, int32_t* did) {
  if (!ds) return IAK3D_ERR_ARG;
  if (xid) memcpy(xid, ds->h_xid.data(), (size_t)ds->n * 4);
  if (did) memcpy(did, ds->h_did.data(), (size_t)ds->n * 4);
  return IAK3D_OK;
}

extern "C" int iak3d_dataset_set_sdfd_spline(iak3d_ds* ds, int32_t comp, int32_t ncol, const double* XsplU) {
  if (!ds) return IAK3D_ERR_ARG;
  if (comp < 0 || comp > 2 || ncol < 0 || ncol > IAK3D_MAX_SPL + 1 || (ncol > 0 && !XsplU)) {
    ds->ctx->err = "dataset_set_sdfd_spline: comp in 0..2, ncol in 0..IAK3D_MAX_SPL+1";
    return IAK3D_ERR_ARG;
  }
  ds->spl_ncol[comp] = ncol;
  ds->h_spl[comp].assign(XsplU, XsplU + (size_t)ncol * ds->ndIU);
  return IAK3D_OK;
}

int iak_get_table_slot(iak3d_ds* ds, size_t k, Slot** out) {
  while (ds->tslots.size() <= k) {
    Slot* s = nullptr;
    const int st = iak_alloc_slot(ds->ctx, ds->n, 0, ds->nxU, ds->ndIU, &s, true);
    if (st != 0) return st;
    ds->tslots.push_back(s);
  }
  *out = ds->tslots[k];
  return 0;
}

int iak_get_slot(iak3d_ds* ds, size_t k, Slot** out) {
  while (ds->slots.size() <= k) {
    Slot* s = nullptr;
    const int st = iak_alloc_slot(ds->ctx, ds->n, ds->q, ds->nxU, ds->ndIU, &s, false);
    if (st != 0) return st;
    ds->slots.push_back(s);
  }
  *out = ds->slots[k];
  return 0;
}

// ================================================================================================
// tables for one evaluation of a dataset (square) into a slot; fills CovArgs
// ================================================================================================
// `cache` (optional) de-duplicates tables inside one batch: the npar+1 evaluations of a forward-difference gradient
// differ in one parameter each, so all but one share the horizontal table and all but one the depth tables.
struct TableCache {
  // per dataset: a composite-likelihood batch holds hundreds of datasets (times npar + 1 parameter sets), and one flat list
  // made every look-up a walk over all of them
  struct Entry { int kind; double k[6]; const double* ptr; const double* av; };
  std::unordered_map<const iak3d_ds*, std::vector<Entry>> e;
  const double* find(const iak3d_ds* ds, int kind, const double* k, const double** av = nullptr) const {
    const auto it = e.find(ds);
    if (it == e.end()) return nullptr;
    for (const Entry& x : it->second)
      if (x.kind == kind && memcmp(x.k, k, sizeof(x.k)) == 0) { if (av) *av = x.av; return x.ptr; }
    return nullptr;
  }
  void add(const iak3d_ds* ds, int kind, const double* k, const double* ptr, const double* av = nullptr) {
    Entry x; x.kind = kind; memcpy(x.k, k, sizeof(x.k)); x.ptr = ptr; x.av = av; e[ds].push_back(x);
  }
};

TableCache* iak_table_cache_new() { return new TableCache(); }
void iak_table_cache_free(TableCache* c) { delete c; }

int iak_square_tables(iak3d_ctx* ctx, const iak3d_ds* ds, const EvalPlan& pl, Slot* s, CovArgs* a, TableCache* cache = nullptr,
                      TableQueue* tq = nullptr /* batches: queue the plain tables instead of launching them (flush_table_queue) */) {
  if (pl.approx) tq = nullptr;         // the spline-sd tables are scaled from a source table right here: launch in order
  const int64_t nd = ds->ndIU, nx = ds->nxU;
  double* const avbase = s->d_phid + (size_t)NTAB_BUF * nd * nd;
  for (int k = 0; k < 3; k++) { a->T[k] = nullptr; a->av[k] = nullptr; }
  a->comb = s->d_phid + (size_t)NTAB_BUF * nd * nd + (size_t)round_up((NTAB_BUF + 3) * nd, 2);
  a->Ebuf = s->d_E; a->E_shared = nullptr;
  auto plain_table = [&](int k, int buf, const double** out, const double** avout) -> int {   // iaCovMatern on the unique intervals
    const IaPrm& P = pl.prm[k];
    const double key[6] = {P.psi, P.tau0, P.tau1, P.tau2, P.g11[1], P.g11[2]};     // (ad, nud, sd function) determine the table
    const double* hit = cache ? cache->find(ds, 1, key, avout) : nullptr;
    if (hit != nullptr) { *out = hit; return 0; }
    double* T = s->d_phid + (size_t)buf * nd * nd;
    if (tq != nullptr) {
      if (nd > 0) tq->ia.push_back(IaTabJob{P, ds->d_dIU, nd, T, avbase + (size_t)buf * nd});
    } else {
      const int st = launch_iacov_table(ctx, P, ds->d_dIU, nd, ds->d_dIU, nd, T, avbase + (size_t)buf * nd);
      if (st != 0) return st;
    }
    *out = T; *avout = avbase + (size_t)buf * nd;
    if (cache) cache->add(ds, 1, key, T, *avout);
    return 0;
  };
  if (!pl.approx) {
    for (int k = 0; k < pl.ntab; k++) {
      const int st = plain_table(k, k, &a->T[k], &a->av[k]);
      if (st != 0) return st;
    }
  } else {
    // setCApproxIAK3D (fitIAK3D.R:1919-2085): the stationary table, then avCor * avsd(I) * avsd(J) per component
    int ksrc = -1;
    for (int k = 0; k < pl.ntab; k++) if (pl.sd_type[k] == 0) ksrc = k;
    const double* src = nullptr; const double* srcav = nullptr;
    const int st0 = plain_table(ksrc >= 0 ? ksrc : 0, ksrc >= 0 ? ksrc : NTAB_BUF - 1, &src, &srcav);
    if (st0 != 0) return st0;
    if (ksrc >= 0) { a->T[ksrc] = src; a->av[ksrc] = srcav; }
    std::vector<double> sv;
    for (int k = 0; k < pl.ntab; k++) {
      if (pl.sd_type[k] == 0) continue;
      const int c = pl.sd_comp[k];
      const int st = iak_avsd_host(ctx, pl, k, ds->h_dIU.data(), nd, ds->h_spl[c].empty() ? nullptr : ds->h_spl[c].data(),
                                   ds->spl_ncol[c], sv);
      if (st != 0) return st;                                         // IAK3D_BAD_PARS: C = NA (:2079)
      double* d_s = avbase + (size_t)NTAB_BUF * nd + (size_t)k * nd;
      CUDA_TRY(ctx, cudaMemcpyAsync(d_s, sv.data(), (size_t)nd * 8, cudaMemcpyHostToDevice, ctx->stream));
      double* T = s->d_phid + (size_t)k * nd * nd;
      const int st2 = launch_scale_table(ctx, src, d_s, nd, d_s, nd, 1, T, avbase + (size_t)k * nd);
      if (st2 != 0) return st2;
      a->T[k] = T; a->av[k] = avbase + (size_t)k * nd;
    }
  }
  for (int c = 0; c < 3; c++) a->tidx[c] = pl.tidx[c];
  a->ntab = pl.ntab;
  a->phix = nullptr; a->same = nullptr;
  if (pl.has_phix) {
    const double key[6] = {(double)pl.H.modelx, pl.H.a, pl.H.nu, 0, 0, 0};
    const double* hit = cache ? cache->find(ds, 2, key) : nullptr;
    if (hit != nullptr) {
      a->phix = hit;
    } else {
      if (tq != nullptr) {
        if (nx > 0) tq->hx.push_back(HxTabJob{pl.H, ds->d_xU, nx, s->d_phix});
      } else {
        const int st = launch_phix_table(ctx, pl.H, ds->d_xU, nx, ds->d_xU, nx, s->d_phix, nullptr);
        if (st != 0) return st;
      }
      a->phix = s->d_phix;
      if (cache) cache->add(ds, 2, key, s->d_phix);
    }
  }
  a->cx0 = pl.cx0; a->cx1 = pl.cx1; a->kd1 = pl.kd1; a->cxd0 = pl.cxd0; a->cxd1 = pl.cxd1; a->cme = pl.cme;
  a->xid_r = ds->d_xid; a->did_r = ds->d_did; a->xid_c = ds->d_xid; a->did_c = ds->d_did;
  a->nr = ds->n; a->nc = ds->n; a->nxU_r = nx; a->ndIU_r = nd;
  a->nssre = 0;
  if (pl.nssre > 0) {      // site-specific random effects (fitIAK3D.R:1091-1101)
    if (ds->nssre != pl.nssre || !ds->d_sid) { ctx->err = "site-specific random effects: pars->nssre does not match iak3d_dataset_set_ssre"; return IAK3D_ERR_ARG; }
    a->nssre = pl.nssre;
    for (int k = 0; k < pl.nssre; k++) a->cssre[k] = pl.cssre[k];
    a->sid_r = a->sid_c = ds->d_sid; a->Xs_r = a->Xs_c = ds->d_Xs; a->ldXs_r = a->ldXs_c = ds->n;
  }
  return 0;
}

extern "C" int iak3d_dataset_set_ssre(iak3d_ds* ds, int32_t nssre, const int32_t* siteid, const double* Xs) {
  if (!ds || nssre < 0 || nssre > IAK3D_MAX_SSRE || (nssre > 0 && (!siteid || !Xs))) return IAK3D_ERR_ARG;
  iak3d_ctx* ctx = ds->ctx;
  cudaSetDevice(ctx->device);
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  cudaFree(ds->d_sid); cudaFree(ds->d_Xs); ds->d_sid = nullptr; ds->d_Xs = nullptr; ds->nssre = 0;
  if (nssre == 0) return IAK3D_OK;
  CUDA_TRY(ctx, cudaMalloc(&ds->d_sid, (size_t)ds->n * 4));
  CUDA_TRY(ctx, cudaMalloc(&ds->d_Xs, (size_t)ds->n * nssre * 8));
  CUDA_TRY(ctx, cudaMemcpyAsync(ds->d_sid, siteid, (size_t)ds->n * 4, cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(ctx, cudaMemcpyAsync(ds->d_Xs, Xs, (size_t)ds->n * nssre * 8, cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  ds->nssre = nssre;
  return IAK3D_OK;
}

// sigma2Vec - diag(C) of one evaluation into s->d_lncol[0..n)
int iak_lnN_column(iak3d_ctx* ctx, const iak3d_ds* ds, const EvalPlan& pl, const CovArgs& a, Slot* s) {
  const int64_t n = ds->n;
  if (s->d_lncol == nullptr) CUDA_TRY(ctx, cudaMalloc(&s->d_lncol, (size_t)2 * n * sizeof(double)));
  int st = launch_sigma2(ctx, pl, a.av, ds->d_did, n, s->d_lncol);
  if (st != 0) return st;
  st = launch_ckk(ctx, pl, a.T, ds->ndIU, ds->d_did, n, s->d_lncol + n);
  if (st != 0) return st;
  if (pl.nssre > 0) { ctx->err = "lognormal data with site-specific random effects: sigma2Vec is undefined (fitIAK3D.R:1092)"; return IAK3D_ERR_ARG; }
  return launch_vec_sub(ctx, s->d_lncol, s->d_lncol + n, n, s->d_lncol);
}

extern "C" int iak3d_dataset_set_lnN_column(iak3d_ds* ds, int32_t col) {
  if (!ds) return IAK3D_ERR_ARG;
  if (col >= 64) { ds->ctx->err = "dataset_set_lnN_column: column out of range"; return IAK3D_ERR_ARG; }
  ds->lncol = col < 0 ? -1 : col;
  return IAK3D_OK;
}

// ================================================================================================
// depth / horizontal kernels on caller-supplied unique sets
// ================================================================================================
extern "C" int iak3d_iacov(iak3d_ctx* ctx, int64_t n1, const double* dI1, int64_t n2, const double* dI2, double ad, double nud,
                           int32_t sdfdType, const double* sdfdPars, double* out, double* avVar) {
  if (!ctx || n1 < 0 || n2 < 0 || (n1 > 0 && !dI1) || (n2 > 0 && !dI2) || !out) return IAK3D_ERR_ARG;
  IaPrm P;
  const double s1 = (sdfdType == -1 && sdfdPars) ? sdfdPars[0] : 0.0, s2 = (sdfdType == -1 && sdfdPars) ? sdfdPars[1] : 0.0;
  if (sdfdType == -1 && !sdfdPars) return IAK3D_ERR_ARG;
  if (iak_prm_init(&P, ad, nud, sdfdType, s1, s2) != 0) return IAK3D_BAD_PARS;
  if (n1 == 0 || n2 == 0) return IAK3D_OK;
  cudaSetDevice(ctx->device);
  double *d1 = nullptr, *d2 = nullptr, *dout = nullptr, *dav = nullptr;
  int rc = IAK3D_OK;
  do {
    if (cudaMalloc(&d1, (size_t)n1 * 16) != cudaSuccess || cudaMalloc(&d2, (size_t)n2 * 16) != cudaSuccess ||
        cudaMalloc(&dout, (size_t)n1 * n2 * 8) != cudaSuccess || cudaMalloc(&dav, (size_t)n1 * 8) != cudaSuccess) {
      ctx->err = "iacov: allocation failed"; rc = IAK3D_ERR_CUDA; break;
    }
    cudaMemcpyAsync(d1, dI1, (size_t)n1 * 16, cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync(d2, dI2, (size_t)n2 * 16, cudaMemcpyHostToDevice, ctx->stream);
    rc = launch_iacov_table(ctx, P, d1, n1, d2, n2, dout, dav);
    if (rc != 0) break;
    cudaMemcpyAsync(out, dout, (size_t)n1 * n2 * 8, cudaMemcpyDeviceToHost, ctx->stream);
    if (avVar) cudaMemcpyAsync(avVar, dav, (size_t)n1 * 8, cudaMemcpyDeviceToHost, ctx->stream);
    const cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { ctx->err = std::string("iacov: ") + cudaGetErrorString(e); rc = IAK3D_ERR_CUDA; }
  } while (0);
  cudaFree(d1); cudaFree(d2); cudaFree(dout); cudaFree(dav);
  return rc;
}

extern "C" int iak3d_spatial_cov(iak3d_ctx* ctx, int64_t m, const double* D, int32_t modelx, double c1, double a, double nu,
                                 double* out) {
  if (!ctx || m < 0 || (m > 0 && (!D || !out))) return IAK3D_ERR_ARG;
  if (modelx < IAK3D_MODELX_SPHERICAL || modelx > IAK3D_MODELX_WENDLAND) return IAK3D_ERR_ARG;
  if (!(c1 > 0.0)) return IAK3D_BAD_PARS;                              // iaCovMatern.R:335,354,379
  HxPrm H;
  if (hx_init(&H, modelx, a, nu) != 0) return IAK3D_BAD_PARS;
  if (m == 0) return IAK3D_OK;
  cudaSetDevice(ctx->device);
  double *dD = nullptr, *dout = nullptr;
  int rc = IAK3D_OK;
  do {
    if (cudaMalloc(&dD, (size_t)m * 8) != cudaSuccess || cudaMalloc(&dout, (size_t)m * 8) != cudaSuccess) {
      ctx->err = "spatial_cov: allocation failed"; rc = IAK3D_ERR_CUDA; break;
    }
    cudaMemcpyAsync(dD, D, (size_t)m * 8, cudaMemcpyHostToDevice, ctx->stream);
    rc = launch_phix_vec(ctx, H, dD, m, dout);
    if (rc != 0) break;
    cudaMemcpyAsync(out, dout, (size_t)m * 8, cudaMemcpyDeviceToHost, ctx->stream);
    const cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { ctx->err = std::string("spatial_cov: ") + cudaGetErrorString(e); rc = IAK3D_ERR_CUDA; break; }
    if (c1 != 1.0) for (int64_t i = 0; i < m; i++) out[i] = c1 * out[i];
  } while (0);
  cudaFree(dD); cudaFree(dout);
  return rc;
}

// ================================================================================================
// covariance assembly returned to the caller
// ================================================================================================
extern "C" int iak3d_cov_build(iak3d_ds* ds, const iak3d_pars* pars, double* C, double* sigma2Vec) {
  if (!ds || !pars) return IAK3D_ERR_ARG;
  iak3d_ctx* ctx = ds->ctx;
  cudaSetDevice(ctx->device);
  EvalPlan pl;
  int st = iak_make_plan(ctx, pars, true, &pl);
  if (st == 0) st = iak_plan_check_avsd(ctx, ds, &pl);
  if (st != 0) return st;
  if (pl.status != 0) return pl.status;
  Slot* s = nullptr;
  st = iak_get_slot(ds, 0, &s);
  if (st != 0) return st;
  CovArgs a;
  st = iak_square_tables(ctx, ds, pl, s, &a);
  if (st != 0) return st;
  const int64_t n = ds->n;
  double *dC = nullptr, *ds2 = nullptr;
  int rc = IAK3D_OK;
  do {
    if (C) {
      if (cudaMalloc(&dC, (size_t)n * n * 8) != cudaSuccess) { ctx->err = "cov_build: allocation failed"; rc = IAK3D_ERR_CUDA; cudaGetLastError(); break; }
      rc = launch_cov_rect(ctx, a, dC, n, n, n, 1);
      if (rc != 0) break;
      cudaMemcpyAsync(C, dC, (size_t)n * n * 8, cudaMemcpyDeviceToHost, ctx->stream);
    }
    if (sigma2Vec) {
      if (cudaMalloc(&ds2, (size_t)n * 8) != cudaSuccess) { ctx->err = "cov_build: allocation failed"; rc = IAK3D_ERR_CUDA; cudaGetLastError(); break; }
      rc = launch_sigma2(ctx, pl, a.av, ds->d_did, n, ds2);
      if (rc != 0) break;
      cudaMemcpyAsync(sigma2Vec, ds2, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream);
      if (pl.nssre > 0) {       // "sigma2Vec <- NA * sigma2Vec # just to remind that this shouldn't be used with ssre" (fitIAK3D.R:1092)
        if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { ctx->err = "cov_build: copy failed"; rc = IAK3D_ERR_CUDA; break; }
        for (int64_t i = 0; i < n; i++) sigma2Vec[i] = nan("");
      }
    }
    const cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { ctx->err = std::string("cov_build: ") + cudaGetErrorString(e); rc = IAK3D_ERR_CUDA; }
  } while (0);
  cudaFree(dC); cudaFree(ds2);
  return rc;
}

// setCIAK3D as the factor build forms it (cov_factor_kernel into the blocked factor buffer), copied back un-blocked
extern "C" int iak3d_factor_build_fetch(iak3d_ds* ds, const iak3d_pars* pars, int32_t batched, double* A, double* Wrows) {
  if (!ds || !pars || !A) return IAK3D_ERR_ARG;
  iak3d_ctx* ctx = ds->ctx;
  cudaSetDevice(ctx->device);
  if (ctx->batch_pending()) { ctx->err = "factor_build_fetch: a batch is pending"; return IAK3D_ERR_ARG; }
  EvalPlan pl;
  int st = iak_make_plan(ctx, pars, true, &pl);
  if (st == 0) st = iak_plan_check_avsd(ctx, ds, &pl);
  if (st != 0) return st;
  if (pl.status != 0) return pl.status;
  const int njobs = batched ? 2 : 1;
  std::vector<FactorBuildJob> jobs((size_t)njobs);
  Slot* s = nullptr;
  for (int k = 0; k < njobs; k++) {
    st = iak_get_slot(ds, (size_t)k, &s);
    if (st != 0) return st;
    CovArgs a;
    st = iak_square_tables(ctx, ds, pl, s, &a);
    if (st != 0) return st;
    FactorBuildJob& jb = jobs[(size_t)k];
    jb.a = a; jb.d_W = ds->d_W; jb.q = ds->q; jb.n = s->n; jb.Npad = s->Npad; jb.ld = s->ld; jb.nCB = s->nCB; jb.d_L = s->d_L;
  }
  st = launch_cov_factor_jobs(ctx, jobs.data(), njobs);
  if (st != 0) return st;
  const int64_t n = ds->n;
  double* d_out = nullptr; double* d_w = nullptr;
  int rc = IAK3D_OK;
  do {
    if (iak_pool_alloc(ctx, (void**)&d_out, (size_t)n * n * 8) != cudaSuccess) { ctx->err = "factor_build_fetch: allocation failed"; rc = IAK3D_ERR_CUDA; cudaGetLastError(); break; }
    rc = launch_unblock_rect(ctx, s->d_L, s->ld / UR, 0, n, n, 1, d_out, n);
    if (rc != 0) break;
    cudaMemcpyAsync(A, d_out, (size_t)n * n * 8, cudaMemcpyDeviceToHost, ctx->stream);
    if (Wrows && ds->q > 0) {
      if (iak_pool_alloc(ctx, (void**)&d_w, (size_t)ds->q * n * 8) != cudaSuccess) { ctx->err = "factor_build_fetch: allocation failed"; rc = IAK3D_ERR_CUDA; cudaGetLastError(); break; }
      rc = launch_unblock_rect(ctx, s->d_L, s->ld / UR, s->Npad, ds->q, n, 0, d_w, ds->q);
      if (rc != 0) break;
      cudaMemcpyAsync(Wrows, d_w, (size_t)ds->q * n * 8, cudaMemcpyDeviceToHost, ctx->stream);
    }
    const cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { ctx->err = std::string("factor_build_fetch: ") + cudaGetErrorString(e); rc = IAK3D_ERR_CUDA; }
  } while (0);
  iak_pool_free(ctx, d_out); iak_pool_free(ctx, d_w);
  return rc;
}

extern "C" int iak3d_cov_diag(iak3d_ds* ds, const iak3d_pars* pars, double* diagC) {
  if (!ds || !pars || !diagC) return IAK3D_ERR_ARG;
  iak3d_ctx* ctx = ds->ctx;
  cudaSetDevice(ctx->device);
  EvalPlan pl;
  int st = iak_make_plan(ctx, pars, true, &pl);
  if (st == 0) st = iak_plan_check_avsd(ctx, ds, &pl);
  if (st != 0) return st;
  if (pl.status != 0) return pl.status;
  Slot* s = nullptr;
  st = iak_get_slot(ds, 0, &s);
  if (st != 0) return st;
  CovArgs a;
  st = iak_square_tables(ctx, ds, pl, s, &a);
  if (st != 0) return st;
  const int64_t n = ds->n;
  if (s->d_lncol == nullptr) CUDA_TRY(ctx, cudaMalloc(&s->d_lncol, (size_t)2 * n * sizeof(double)));
  st = launch_ckk(ctx, pl, a.T, ds->ndIU, ds->d_did, n, s->d_lncol + n);
  if (st != 0) return st;
  st = launch_ssre_diag(ctx, a.nssre, a.cssre, a.Xs_r, a.ldXs_r, n, s->d_lncol + n);
  if (st != 0) return st;
  CUDA_TRY(ctx, cudaMemcpyAsync(diagC, s->d_lncol + n, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return IAK3D_OK;
}

// Cross-covariance tables between set 1 (given as unique tables + ids on the device) and a dataset (set 2).
// Allocates into *bufs (freed by the caller with cross_free).
struct CrossBufs {
  double* d_x1 = nullptr; double* d_d1 = nullptr; int32_t* d_xid1 = nullptr; int32_t* d_did1 = nullptr;
  double* d_phix = nullptr; uint8_t* d_same = nullptr; double* d_T = nullptr; double* d_s = nullptr;
  int64_t nx1 = 0, nd1 = 0;
};
void iak_cross_free(CrossBufs* b) {
  cudaFree(b->d_x1); cudaFree(b->d_d1); cudaFree(b->d_xid1); cudaFree(b->d_did1); cudaFree(b->d_phix); cudaFree(b->d_same);
  cudaFree(b->d_T); cudaFree(b->d_s);
  *b = CrossBufs();
}
// Depth tables between the unique intervals of set 1 (device d_d1, host dIU1) and of a dataset: iaCovMatern2 per table
// (fitIAK3D.R:1156-1184) or, with spline sd functions, the stationary table scaled by avsd1(I) * avsd2(J) (:2128-2168).
// d_T: NTAB_BUF x nd1 x nd2; d_s: 3 x (nd1 + nd2) scratch.  spl1U[c]: nd1 x ncol basis of set 1 (may be null).
// Returns IAK3D_BAD_PARS when an averaged sd is negative (C = NA, :2172).
int iak_cross_depth_tables(iak3d_ctx* ctx, const iak3d_ds* ds2, const EvalPlan& pl, const double* d_d1, const double* dIU1,
                           int64_t nd1, const std::vector<double>* spl1U, double* d_T, double* d_s, const double** Tout) {
  const int64_t nd2 = ds2->ndIU;
  for (int k = 0; k < 3; k++) Tout[k] = nullptr;
  if (!pl.approx) {
    for (int k = 0; k < pl.ntab; k++) {
      double* T = d_T + (size_t)k * nd1 * nd2;
      const int st = launch_iacov_table(ctx, pl.prm[k], d_d1, nd1, ds2->d_dIU, nd2, T, nullptr);
      if (st != 0) return st;
      Tout[k] = T;
    }
    return 0;
  }
  int ksrc = -1;
  for (int k = 0; k < pl.ntab; k++) if (pl.sd_type[k] == 0) ksrc = k;
  double* src = d_T + (size_t)(ksrc >= 0 ? ksrc : NTAB_BUF - 1) * nd1 * nd2;
  int st = launch_iacov_table(ctx, pl.prm[0], d_d1, nd1, ds2->d_dIU, nd2, src, nullptr);
  if (st != 0) return st;
  if (ksrc >= 0) Tout[ksrc] = src;
  std::vector<double> s1, s2;
  for (int k = 0; k < pl.ntab; k++) {
    if (pl.sd_type[k] == 0) continue;
    const int c = pl.sd_comp[k];
    const std::vector<double>* b1 = spl1U ? &spl1U[c] : nullptr;
    const int ncol1 = (b1 && nd1 > 0) ? (int)(b1->size() / (size_t)nd1) : 0;
    st = iak_avsd_host(ctx, pl, k, dIU1, nd1, (b1 && !b1->empty()) ? b1->data() : nullptr, ncol1, s1);
    if (st != 0) return st;
    st = iak_avsd_host(ctx, pl, k, ds2->h_dIU.data(), nd2, ds2->h_spl[c].empty() ? nullptr : ds2->h_spl[c].data(), ds2->spl_ncol[c], s2);
    if (st != 0) return st;
    double* ds1 = d_s + (size_t)k * (nd1 + nd2); double* ds2v = ds1 + nd1;
    CUDA_TRY(ctx, cudaMemcpyAsync(ds1, s1.data(), (size_t)nd1 * 8, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(ds2v, s2.data(), (size_t)nd2 * 8, cudaMemcpyHostToDevice, ctx->stream));
    double* T = d_T + (size_t)k * nd1 * nd2;
    st = launch_scale_table(ctx, src, ds1, nd1, ds2v, nd2, 0, T, nullptr);
    if (st != 0) return st;
    Tout[k] = T;
  }
  return 0;
}

// rows of a per-row basis (n1 x ncol) at the first occurrence of each unique interval -> nd1 x ncol
void iak_spl_unique_rows(const double* Xrow, int64_t n1, int ncol, const std::vector<int32_t>& did1, int64_t nd1, std::vector<double>& out) {
  out.assign((size_t)nd1 * ncol, 0.0);
  std::vector<char> seen((size_t)nd1, 0);
  for (int64_t i = 0; i < n1; i++) {
    const int32_t u = did1[(size_t)i];
    if (seen[(size_t)u]) continue;
    seen[(size_t)u] = 1;
    for (int e = 0; e < ncol; e++) out[(size_t)e * nd1 + u] = Xrow[(size_t)e * n1 + i];
  }
}

int iak_cross_tables(iak3d_ctx* ctx, const iak3d_ds* ds2, const EvalPlan& pl, int64_t n1, const std::vector<int32_t>& xid1,
                     const std::vector<double>& xU1, const std::vector<int32_t>& did1, const std::vector<double>& dIU1,
                     const std::vector<double>* spl1U, CrossBufs* b, CovArgs* a) {
  const int64_t nx1 = (int64_t)xU1.size() / 2, nd1 = (int64_t)dIU1.size() / 2;
  b->nx1 = nx1; b->nd1 = nd1;
  cudaError_t e = cudaSuccess;
  auto up = [&](void** p, const void* src, size_t bytes) {
    if (e == cudaSuccess) e = cudaMalloc(p, std::max<size_t>(bytes, 16));
    if (e == cudaSuccess && bytes) e = cudaMemcpyAsync(*p, src, bytes, cudaMemcpyHostToDevice, ctx->stream);
  };
  up((void**)&b->d_x1, xU1.data(), xU1.size() * 8);
  up((void**)&b->d_d1, dIU1.data(), dIU1.size() * 8);
  up((void**)&b->d_xid1, xid1.data(), (size_t)n1 * 4);
  up((void**)&b->d_did1, did1.data(), (size_t)n1 * 4);
  if (e == cudaSuccess) e = cudaMalloc(&b->d_phix, std::max<size_t>((size_t)nx1 * ds2->nxU * 8, 16));
  if (e == cudaSuccess) e = cudaMalloc(&b->d_same, std::max<size_t>((size_t)nx1 * ds2->nxU, 16));
  if (e == cudaSuccess) e = cudaMalloc(&b->d_T, std::max<size_t>((size_t)NTAB_BUF * nd1 * ds2->ndIU * 8, 16));
  if (e == cudaSuccess) e = cudaMalloc(&b->d_s, std::max<size_t>((size_t)3 * (nd1 + ds2->ndIU) * 8, 16));
  if (e != cudaSuccess) { ctx->err = std::string("cross tables: ") + cudaGetErrorString(e); cudaGetLastError(); return IAK3D_ERR_CUDA; }
  int st = iak_cross_depth_tables(ctx, ds2, pl, b->d_d1, dIU1.data(), nd1, spl1U, b->d_T, b->d_s, a->T);
  if (st != 0) return st;
  for (int c = 0; c < 3; c++) a->tidx[c] = pl.tidx[c];
  a->ntab = pl.ntab;
  // phix0 = 1[Dx == 0] (fitIAK3D.R:1138-1141) always; phix1 unless nugget
  st = launch_phix_table(ctx, pl.H, b->d_x1, nx1, ds2->d_xU, ds2->nxU, pl.has_phix ? b->d_phix : nullptr, b->d_same);
  if (st != 0) return st;
  a->phix = pl.has_phix ? b->d_phix : nullptr;
  a->same = b->d_same;
  a->cx0 = pl.cx0; a->cx1 = pl.cx1; a->kd1 = pl.kd1; a->cxd0 = pl.cxd0; a->cxd1 = pl.cxd1; a->cme = 0.0;   // no cme (:1243)
  a->xid_r = b->d_xid1; a->did_r = b->d_did1; a->xid_c = ds2->d_xid; a->did_c = ds2->d_did;
  a->nr = n1; a->nc = ds2->n; a->nxU_r = nx1; a->ndIU_r = nd1;
  return 0;
}

static int cov_cross_impl(iak3d_ds* ds2, const iak3d_pars* pars, int64_t n1, const double* xy1, const double* dI1,
                          const double* const* Xspl1, const int32_t* siteid1, const double* Xs1, double* out);
extern "C" int iak3d_cov_cross_spl(iak3d_ds* ds2, const iak3d_pars* pars, int64_t n1, const double* xy1, const double* dI1,
                                   const double* const* Xspl1, double* out) {
  return cov_cross_impl(ds2, pars, n1, xy1, dI1, Xspl1, nullptr, nullptr, out);
}
extern "C" int iak3d_cov_cross_ssre(iak3d_ds* ds2, const iak3d_pars* pars, int64_t n1, const double* xy1, const double* dI1,
                                    const int32_t* siteid1, const double* Xs1, double* out) {
  if (pars && pars->nssre > 0 && (!siteid1 || !Xs1)) return IAK3D_ERR_ARG;
  return cov_cross_impl(ds2, pars, n1, xy1, dI1, nullptr, siteid1, Xs1, out);
}
static int cov_cross_impl(iak3d_ds* ds2, const iak3d_pars* pars, int64_t n1, const double* xy1, const double* dI1,
                          const double* const* Xspl1, const int32_t* siteid1, const double* Xs1, double* out) {
  if (!ds2 || !pars || n1 < 0 || (n1 > 0 && (!xy1 || !dI1 || !out))) return IAK3D_ERR_ARG;
  iak3d_ctx* ctx = ds2->ctx;
  cudaSetDevice(ctx->device);
  EvalPlan pl;
  int st = iak_make_plan(ctx, pars, false, &pl);
  if (st != 0) return st;
  if (pl.status != 0) return pl.status;
  if (n1 == 0) return IAK3D_OK;
  std::vector<int32_t> xid1, did1; std::vector<double> xU1, dIU1;
  iak_unique_first(n1, xy1, xy1 + n1, xid1, xU1);
  iak_unique_first(n1, dI1, dI1 + n1, did1, dIU1);
  std::vector<double> spl1U[3];
  for (int c = 0; c < 3; c++)
    if (pars->sdfdType[c] > 0 && Xspl1 && Xspl1[c]) iak_spl_unique_rows(Xspl1[c], n1, pars->sdfdType[c] + 1, did1, (int64_t)dIU1.size() / 2, spl1U[c]);
  CrossBufs b; CovArgs a;
  double* dout = nullptr; int32_t* d_sid1 = nullptr; double* d_Xs1 = nullptr;
  int rc = iak_cross_tables(ctx, ds2, pl, n1, xid1, xU1, did1, dIU1, spl1U, &b, &a);
  do {
    if (rc != 0) break;
    if (pl.nssre > 0) {       // setCIAK3D2 with siteIDData / siteIDData2 (fitIAK3D.R:1233-1240)
      if (!siteid1 || !Xs1 || ds2->nssre != pl.nssre) { ctx->err = "cov_cross: site ids / columns of set 1 (iak3d_cov_cross_ssre) and iak3d_dataset_set_ssre are needed"; rc = IAK3D_ERR_ARG; break; }
      if (cudaMalloc(&d_sid1, (size_t)n1 * 4) != cudaSuccess || cudaMalloc(&d_Xs1, (size_t)n1 * pl.nssre * 8) != cudaSuccess) { ctx->err = "cov_cross: allocation failed"; rc = IAK3D_ERR_CUDA; cudaGetLastError(); break; }
      cudaMemcpyAsync(d_sid1, siteid1, (size_t)n1 * 4, cudaMemcpyHostToDevice, ctx->stream);
      cudaMemcpyAsync(d_Xs1, Xs1, (size_t)n1 * pl.nssre * 8, cudaMemcpyHostToDevice, ctx->stream);
      a.nssre = pl.nssre;
      for (int k = 0; k < pl.nssre; k++) a.cssre[k] = pl.cssre[k];
      a.sid_r = d_sid1; a.Xs_r = d_Xs1; a.ldXs_r = n1; a.sid_c = ds2->d_sid; a.Xs_c = ds2->d_Xs; a.ldXs_c = ds2->n;
    }
    if (cudaMalloc(&dout, (size_t)n1 * ds2->n * 8) != cudaSuccess) { ctx->err = "cov_cross: allocation failed"; rc = IAK3D_ERR_CUDA; cudaGetLastError(); break; }
    rc = launch_cov_rect(ctx, a, dout, n1, n1, ds2->n, 0);
    if (rc != 0) break;
    cudaMemcpyAsync(out, dout, (size_t)n1 * ds2->n * 8, cudaMemcpyDeviceToHost, ctx->stream);
    const cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { ctx->err = std::string("cov_cross: ") + cudaGetErrorString(e); rc = IAK3D_ERR_CUDA; }
  } while (0);
  cudaStreamSynchronize(ctx->stream);
  cudaFree(dout); cudaFree(d_sid1); cudaFree(d_Xs1);
  iak_cross_free(&b);
  return rc;
}

extern "C" int iak3d_cov_cross(iak3d_ds* ds2, const iak3d_pars* pars, int64_t n1, const double* xy1, const double* dI1, double* out) {
  return iak3d_cov_cross_spl(ds2, pars, n1, xy1, dI1, nullptr, out);
}

// ================================================================================================
// likelihood terms: the batch pipeline
// ================================================================================================
static int batch_reserve(iak3d_ctx* ctx, Batch* b, int32_t count, int32_t q) {
  if (count > b->probs_cap) {
    cudaFree(b->d_probs); b->d_probs = nullptr;
    CUDA_TRY(ctx, cudaMalloc(&b->d_probs, (size_t)count * sizeof(ProblemDesc)));
    b->probs_cap = count;
  }
  if (count > b->status_cap) {
    cudaFree(b->d_status); b->d_status = nullptr;
    CUDA_TRY(ctx, cudaMalloc(&b->d_status, (size_t)count * 2 * sizeof(int32_t)));
    b->status_cap = count;
  }
  const int32_t stride = 2 + q * q;
  if ((int64_t)count * stride > b->results_cap) {
    cudaFree(b->d_results); b->d_results = nullptr;
    if (b->h_results) { cudaFreeHost(b->h_results); b->h_results = nullptr; }
    CUDA_TRY(ctx, cudaMalloc(&b->d_results, (size_t)count * stride * sizeof(double)));
    CUDA_TRY(ctx, cudaMallocHost(&b->h_results, (size_t)count * stride * sizeof(double)));
    b->results_cap = (int64_t)count * stride;
  }
  b->stride = stride;
  return 0;
}

// Task order: by column block, then problem; inside a (column j, problem) group the tile that the NEXT diagonal
// block needs comes first, then the next column's diagonal task itself (look-ahead: its long K-loop and its
// potf2 overlap with the rest of column j, so column j+1's tiles find the inverse diagonal block ready), then the
// remaining row blocks.  Every dependency of a task (same problem: earlier columns of its row block and of the
// column block's rows; the column's own diagonal task) precedes it, which is all the ticket scheduler needs.
static int batch_tasks(iak3d_ctx* ctx, Batch* b, const std::vector<ProblemDesc>& probs, const std::vector<int>& live) {
  std::vector<int64_t> key;
  key.reserve(live.size() * 3);
  for (size_t i = 0; i < live.size(); i++) { key.push_back(live[i]); key.push_back(probs[i].nCB); key.push_back(probs[i].nRB); }
  if (key == b->key && b->d_tasks != nullptr) return 0;
  std::vector<int4> tasks;
  int maxCB = 0;
  for (size_t i = 0; i < live.size(); i++) if (live[i]) maxCB = std::max(maxCB, (int)probs[i].nCB);
  for (int j = 0; j < maxCB; j++)
    for (size_t i = 0; i < live.size(); i++) {
      if (!live[i] || j >= probs[i].nCB) continue;
      const int nCB = probs[i].nCB, nRB = probs[i].nRB;
      const int rd = j >> 1;                       // row block of this column's diagonal task
      if (j == 0) tasks.push_back(make_int4((int)i, 0, 0, 0));
      const int rn = (j + 1) >> 1;                 // row block holding the next diagonal block
      const bool have_next = (j + 1 < nCB);
      if (have_next && rn != rd) tasks.push_back(make_int4((int)i, rn, j, 0));
      if (have_next) tasks.push_back(make_int4((int)i, rn, j + 1, 0));          // look-ahead diagonal task
      for (int r = rd + 1; r < nRB; r++)
        if (!(have_next && r == rn)) tasks.push_back(make_int4((int)i, r, j, 0));
    }
  if ((int64_t)tasks.size() > b->tasks_cap) {
    cudaFree(b->d_tasks); b->d_tasks = nullptr;
    CUDA_TRY(ctx, cudaMalloc(&b->d_tasks, std::max<size_t>(tasks.size(), 1) * sizeof(int4)));
    b->tasks_cap = (int64_t)tasks.size();
  }
  if (!tasks.empty())
    CUDA_TRY(ctx, cudaMemcpyAsync(b->d_tasks, tasks.data(), tasks.size() * sizeof(int4), cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  b->ntasks = (int32_t)tasks.size();
  b->key = key;
  return 0;
}

static void fill_problem(ProblemDesc* pd, Slot* s, int32_t* d_status) {
  pd->L = s->d_L; pd->nRUL = s->ld / UR; pd->P = s->d_L; pd->nRUP = s->ld / UR;
  pd->invd = s->d_invd;
  pd->doneL = s->d_flags; pd->doneP = s->d_flags; pd->invdone = s->d_flags + (size_t)s->nRB * s->nCB;
  pd->nCB = s->nCB; pd->nRB = s->nRB; pd->mode = 0; pd->epoch = s->epoch;
  pd->w0 = (int32_t)s->Npad; pd->q = s->q; pd->nlive = (int32_t)s->n;
  pd->G = s->d_G; pd->ldG = s->q;
  pd->logsum = s->d_logsum; pd->status = d_status; pd->prof = nullptr;
}

extern "C" int iak3d_nll_terms_batch_enqueue(iak3d_ctx* ctx, int32_t count, iak3d_ds* const* dss, const iak3d_pars* pars) {
  if (!ctx || count <= 0 || !dss || !pars) return IAK3D_ERR_ARG;
  cudaSetDevice(ctx->device);
  Batch* b = ctx->batch;
  if (b->pending) { ctx->err = "nll_terms_batch_enqueue: previous batch not fetched"; return IAK3D_ERR_ARG; }
  const int32_t q = dss[0]->q;
  for (int i = 0; i < count; i++) {
    if (!dss[i] || dss[i]->ctx != ctx) { ctx->err = "batch: dataset belongs to another context"; return IAK3D_ERR_ARG; }
    if (dss[i]->q != q) { ctx->err = "batch: all datasets must share q"; return IAK3D_ERR_ARG; }
  }
  int st = batch_reserve(ctx, b, count, q);
  if (st != 0) return st;
  // plans (host) and slots
  std::vector<EvalPlan> plans((size_t)count);
  b->ds.assign(dss, dss + count);
  b->slots.assign((size_t)count, nullptr);
  b->plan_status.assign((size_t)count, 0);
  std::unordered_map<iak3d_ds*, size_t> uses;
  std::vector<int> live((size_t)count, 0);
  for (int i = 0; i < count; i++) {
    st = iak_make_plan(ctx, &pars[i], true, &plans[i]);
    if (st == 0) st = iak_plan_check_avsd(ctx, dss[i], &plans[i]);
    if (st != 0) return st;
    b->plan_status[i] = plans[i].status;
    const size_t k = uses[dss[i]]++;
    st = iak_get_slot(dss[i], k, &b->slots[i]);
    if (st != 0) return st;
    live[i] = (plans[i].status == 0);
  }
  std::vector<ProblemDesc> probs((size_t)count);
  double flops = 0.0, bytes = 0.0;
  for (int i = 0; i < count; i++) {
    Slot* s = b->slots[i];
    s->epoch++;
    fill_problem(&probs[i], s, b->d_status + 2 * i);
    if (live[i]) { const double n = (double)s->n; flops += n * n * n / 3.0; bytes += 8.0 * n * (n + 1.0) / 2.0; }
  }
  st = batch_tasks(ctx, b, probs, live);
  if (st != 0) return st;
  ctx->last_chol_flops = flops; ctx->last_cov_bytes = bytes;
  if (getenv("IAK3D_PROF") != nullptr) {        // diagnostics: per-task timestamps (iak3d_last_task_profile)
    if (b->ntasks > ctx->prof_cap) {
      cudaFree(ctx->d_prof); ctx->d_prof = nullptr;
      CUDA_TRY(ctx, cudaMalloc(&ctx->d_prof, (size_t)b->ntasks * 8 * sizeof(long long)));
      ctx->prof_cap = b->ntasks;
    }
    ctx->prof_ntasks = b->ntasks; ctx->d_prof_tasks = b->d_tasks;
    for (int i = 0; i < count; i++) probs[i].prof = ctx->d_prof;
  }

  CUDA_TRY(ctx, cudaEventRecord(ctx->evs[0], ctx->stream));
  std::vector<CovArgs> cargs((size_t)count);
  TableCache cache;
  TableQueue tq;                                  // the plain tables of all problems: one launch per kind (tables.cu)
  TableQueue* const tqp = (count > 16 && !ctx->no_table_queue) ? &tq : nullptr;   // small batches: nothing to gain, one sync to lose
  for (int i = 0; i < count; i++) {
    if (!live[i]) continue;
    st = iak_square_tables(ctx, dss[i], plans[i], b->slots[i], &cargs[i], &cache, tqp);
    if (st != 0) return st;
  }
  st = flush_table_queue(ctx, tqp);
  if (st != 0) return st;
  CUDA_TRY(ctx, cudaEventRecord(ctx->evs[1], ctx->stream));
  std::vector<FactorBuildJob> jobs;
  jobs.reserve((size_t)count);
  std::unordered_map<const iak3d_ds*, std::vector<int>> seen;
  for (int i = 0; i < count; i++) {
    if (!live[i]) continue;
    Slot* s = b->slots[i];
    if (cov_one_table(cargs[i])) {           // the expansion depends on (dataset, table) only: share it inside the batch
      std::vector<int>& prev = seen[dss[i]];   // earlier problems of the same dataset (not a walk over the whole batch)
      for (int k : prev)
        if (cov_one_table(cargs[k]) && cargs[k].T[0] == cargs[i].T[0]) {
          cargs[i].E_shared = cargs[k].E_shared ? cargs[k].E_shared : cargs[k].Ebuf;
          break;
        }
      prev.push_back(i);
    }
    if (dss[i]->lncol >= 0 && dss[i]->lncol < q) {
      // lognormal data: column lncol of W is sigma2Vec - diag(C) of THIS evaluation (nrUpdatesIAK3DlnN.R:15, 194)
      st = iak_lnN_column(ctx, dss[i], plans[i], cargs[i], s);
      if (st != 0) return st;
    }
    FactorBuildJob jb;
    jb.a = cargs[i]; jb.d_W = dss[i]->d_W; jb.q = q; jb.n = s->n; jb.Npad = s->Npad; jb.ld = s->ld; jb.nCB = s->nCB; jb.d_L = s->d_L;
    jb.d_lncol = dss[i]->lncol >= 0 ? s->d_lncol : nullptr; jb.lncol = dss[i]->lncol;
    jobs.push_back(jb);
  }
  st = launch_cov_factor_jobs(ctx, jobs.data(), (int)jobs.size());   // ONE launch for the whole batch
  if (st != 0) return st;
  CUDA_TRY(ctx, cudaEventRecord(ctx->evs[2], ctx->stream));
  CUDA_TRY(ctx, cudaMemsetAsync(b->d_status, 0, (size_t)count * 2 * sizeof(int32_t), ctx->stream));
  CUDA_TRY(ctx, cudaMemcpyAsync(b->d_probs, probs.data(), (size_t)count * sizeof(ProblemDesc), cudaMemcpyHostToDevice, ctx->stream));
  int maxCB = 1;
  for (int i = 0; i < count; i++) if (live[i]) maxCB = std::max(maxCB, (int)probs[i].nCB);
  st = launch_tile_tasks(ctx, b->d_probs, b->d_tasks, b->ntasks, b->d_ticket, 0, /*latency_bound=*/count <= 2 ? 1 : 0, b->ntasks / maxCB);
  if (st != 0) return st;
  CUDA_TRY(ctx, cudaEventRecord(ctx->evs[3], ctx->stream));
  st = launch_finish(ctx, b->d_probs, count, b->d_results, b->stride);
  if (st != 0) return st;
  CUDA_TRY(ctx, cudaMemcpyAsync(b->h_results, b->d_results, (size_t)count * b->stride * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaEventRecord(ctx->evs[4], ctx->stream));
  ctx->stage_events_valid = true;
  b->count = count; b->q = q; b->pending = true;
  return IAK3D_OK;
}

extern "C" int iak3d_nll_terms_batch_fetch(iak3d_ctx* ctx, double* lndetA, double* WiAW, int32_t* status) {
  if (!ctx) return IAK3D_ERR_ARG;
  Batch* b = ctx->batch;
  if (!b->pending) { ctx->err = "nll_terms_batch_fetch: nothing enqueued"; return IAK3D_ERR_ARG; }
  b->pending = false;
  cudaSetDevice(ctx->device);
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  const int q = b->q, qq = q * q;
  int rc = IAK3D_OK;
  for (int i = 0; i < b->count; i++) {
    const double* r = b->h_results + (size_t)i * b->stride;
    int st = b->plan_status[i];
    if (st == 0) {
      const int flags = (int)r[1];
      if (flags & 2) { ctx->err = "factorisation: dependency wait timed out (internal error)"; rc = IAK3D_ERR_TIMEOUT; st = IAK3D_ERR_TIMEOUT; }
      else if ((flags & 1) || !isfinite(r[0])) st = IAK3D_NOT_SPD;
    }
    if (status) status[i] = st;
    const double nanv = nan("");
    if (lndetA) lndetA[i] = (st == 0) ? r[0] : nanv;
    if (WiAW) for (int e = 0; e < qq; e++) WiAW[(size_t)i * qq + e] = (st == 0) ? r[2 + e] : nanv;
  }
  return rc;
}

extern "C" int iak3d_nll_terms_batch(iak3d_ctx* ctx, int32_t count, iak3d_ds* const* ds, const iak3d_pars* pars, double* lndetA,
                                     double* WiAW, int32_t* status) {
  const int st = iak3d_nll_terms_batch_enqueue(ctx, count, ds, pars);
  if (st != 0) return st;
  return iak3d_nll_terms_batch_fetch(ctx, lndetA, WiAW, status);
}

// ---- weighted on-device sum of a batch's terms (the composite likelihood's exchange buffer) ----
// one block per group; thread e owns entry e of the q x q sum and walks the units in order: deterministic
__global__ void wsum_kernel(const double* __restrict__ results, int stride, int count, const double* __restrict__ weight,
                            const int32_t* __restrict__ group, const int32_t* __restrict__ plan_bad, int q, double* __restrict__ out) {
  const int g = blockIdx.x, qq = q * q;
  double* o = out + (int64_t)g * (qq + 2);
  for (int e = threadIdx.x; e < qq + 2; e += blockDim.x) {
    double s = 0.0;
    for (int i = 0; i < count; i++) {
      if (group[i] != g) continue;
      const double* r = results + (int64_t)i * stride;
      const bool bad = plan_bad[i] != 0 || r[1] != 0.0 || !isfinite(r[0]);
      if (e == qq + 1) { s += bad ? 1.0 : 0.0; continue; }
      if (bad) continue;
      if (e == qq) { s += weight[i] * r[0]; continue; }
      const int a = e % q, b = e / q;                        // column-major (a, b); keep the upper triangle: G(min, max)
      const int lo = a < b ? a : b, hi = a < b ? b : a;
      s += weight[i] * r[2 + (int64_t)hi * q + lo];
    }
    o[e] = s;
  }
}

extern "C" int iak3d_nll_terms_batch_wsum(iak3d_ctx* ctx, int32_t count, iak3d_ds* const* ds, const iak3d_pars* pars,
                                          const double* weight, const int32_t* group, int32_t ngroups, double* d_out, double* h_out) {
  if (!ctx || !weight || !group || ngroups <= 0 || (!d_out && !h_out)) return IAK3D_ERR_ARG;
  for (int i = 0; i < count; i++) if (group[i] < 0 || group[i] >= ngroups) { ctx->err = "batch_wsum: group out of range"; return IAK3D_ERR_ARG; }
  int st = iak3d_nll_terms_batch_enqueue(ctx, count, ds, pars);
  if (st != 0) return st;
  Batch* b = ctx->batch;
  b->pending = false;
  const int q = b->q, qq = q * q;
  double* d_w = nullptr; int32_t* d_g = nullptr; int32_t* d_bad = nullptr; double* d_own = nullptr;
  int rc = IAK3D_OK;
  do {
    cudaError_t e = iak_pool_alloc(ctx, (void**)&d_w, (size_t)count * 8);
    if (e == cudaSuccess && !d_out) { e = iak_pool_alloc(ctx, (void**)&d_own, (size_t)ngroups * (qq + 2) * 8); d_out = d_own; }
    if (e == cudaSuccess) e = iak_pool_alloc(ctx, (void**)&d_g, (size_t)count * 4);
    if (e == cudaSuccess) e = iak_pool_alloc(ctx, (void**)&d_bad, (size_t)count * 4);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_w, weight, (size_t)count * 8, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_g, group, (size_t)count * 4, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_bad, b->plan_status.data(), (size_t)count * 4, cudaMemcpyHostToDevice, ctx->stream);
    if (e != cudaSuccess) { ctx->err = std::string("batch_wsum: ") + cudaGetErrorString(e); rc = IAK3D_ERR_CUDA; cudaGetLastError(); break; }
    wsum_kernel<<<ngroups, 128, 0, ctx->stream>>>(b->d_results, b->stride, count, d_w, d_g, d_bad, q, d_out);
    ctx->launches++;
    if (h_out) cudaMemcpyAsync(h_out, d_out, (size_t)ngroups * (qq + 2) * 8, cudaMemcpyDeviceToHost, ctx->stream);
    e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { ctx->err = std::string("batch_wsum: ") + cudaGetErrorString(e); rc = IAK3D_ERR_CUDA; break; }
    for (int i = 0; i < count; i++)
      if (b->plan_status[i] == 0 && ((int)b->h_results[(size_t)i * b->stride + 1] & 2)) {
        ctx->err = "factorisation: dependency wait timed out (internal error)"; rc = IAK3D_ERR_TIMEOUT; break;
      }
  } while (0);
  cudaStreamSynchronize(ctx->stream);
  iak_pool_free(ctx, d_w); iak_pool_free(ctx, d_g); iak_pool_free(ctx, d_bad); iak_pool_free(ctx, d_own);
  return rc;
}

// The composite likelihood over several GPUs from ONE host process.  R is single-threaded and a CUDA context does not survive
// fork() (the reference's mcmapply, compositeLikIAK3D.R:54), so the library runs one host thread per device: device d
// evaluates its count[d] (dataset, parameter) units (datasets created on ctxs[d]) with iak3d_nll_terms_batch_wsum and the
// per-device sums are added on the host in device order (deterministic).  h_out: ngroups * (q*q + 2) doubles.
extern "C" int iak3d_nll_terms_batch_wsum_multi(int32_t ndev, iak3d_ctx* const* ctxs, const int32_t* count, iak3d_ds* const* ds,
                                                const iak3d_pars* pars, const double* weight, const int32_t* group, int32_t ngroups,
                                                int32_t q, double* h_out) {
  if (ndev < 1 || !ctxs || !count || !h_out || ngroups <= 0 || q < 0) return IAK3D_ERR_ARG;
  const size_t len = (size_t)ngroups * ((size_t)q * q + 2);
  std::vector<std::vector<double>> part((size_t)ndev, std::vector<double>(len, 0.0));
  std::vector<int> rc((size_t)ndev, 0);
  std::vector<int64_t> off((size_t)ndev + 1, 0);
  for (int d = 0; d < ndev; d++) { if (!ctxs[d] || count[d] < 0) return IAK3D_ERR_ARG; off[d + 1] = off[d] + count[d]; }
  std::vector<std::thread> th;
  for (int d = 0; d < ndev; d++) {
    if (count[d] == 0) continue;
    th.emplace_back([&, d]() {
      const int64_t o = off[d];
      rc[d] = iak3d_nll_terms_batch_wsum(ctxs[d], count[d], ds + o, pars + o, weight + o, group + o, ngroups, nullptr, part[d].data());
    });
  }
  for (auto& t : th) t.join();
  for (int d = 0; d < ndev; d++) if (rc[d] != 0) return rc[d];
  for (size_t k = 0; k < len; k++) { double sum = 0.0; for (int d = 0; d < ndev; d++) sum += part[d][k]; h_out[k] = sum; }
  return IAK3D_OK;
}

extern "C" int iak3d_dataset_slot_bytes(iak3d_ds* ds, int64_t* slot_bytes, int64_t* free_bytes) {
  if (!ds) return IAK3D_ERR_ARG;
  cudaSetDevice(ds->ctx->device);
  if (slot_bytes) {
    const int64_t nd = ds->ndIU, nx = ds->nxU;
    *slot_bytes = 8 * (ubuf_doubles(ds->ld, ds->Npad) + nx * nx + (NTAB_BUF + 3) * nd * nd + 2 * nd * round_up(ds->n, 16) +
                       (int64_t)ds->nCB * INVD_BLOCK) + 4 * ((int64_t)ds->nRB * ds->nCB + ds->nCB);
  }
  if (free_bytes) {
    size_t fr = 0, tot = 0;
    CUDA_TRY(ds->ctx, cudaMemGetInfo(&fr, &tot));
    *free_bytes = (int64_t)fr;
  }
  return IAK3D_OK;
}

// iAW = A^-1 W from the slot's factor (rows Npad.. hold Y' = (L^-1 W)').  out: n x q column-major (host).
int iak_backsolve_to_host(iak3d_ctx* ctx, Slot* s, double* out_host, double* d_out_opt) {
  const int q = s->q;
  if (q <= 0) return 0;
  double* dX = nullptr;
  int rc = 0;
  do {
    if (cudaMalloc(&dX, (size_t)s->Npad * q * 8) != cudaSuccess) {
      ctx->err = "backsolve: allocation failed"; rc = IAK3D_ERR_CUDA; cudaGetLastError(); break;
    }
    rc = launch_backsolve(ctx, s->d_L, s->ld / UR, s->Npad, s->d_invd, q, dX);
    if (rc != 0) break;
    if (out_host)
      cudaMemcpy2DAsync(out_host, (size_t)s->n * 8, dX, (size_t)s->Npad * 8, (size_t)s->n * 8, (size_t)q, cudaMemcpyDeviceToHost, ctx->stream);
    if (d_out_opt)
      cudaMemcpy2DAsync(d_out_opt, (size_t)s->n * 8, dX, (size_t)s->Npad * 8, (size_t)s->n * 8, (size_t)q, cudaMemcpyDeviceToDevice, ctx->stream);
    const cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { ctx->err = std::string("backsolve: ") + cudaGetErrorString(e); rc = IAK3D_ERR_CUDA; }
  } while (0);
  cudaFree(dX);
  return rc;
}

extern "C" int iak3d_nll_terms(iak3d_ds* ds, const iak3d_pars* pars, double* lndetA, double* WiAW, double* iAW) {
  if (!ds || !pars) return IAK3D_ERR_ARG;
  iak3d_ctx* ctx = ds->ctx;
  iak3d_ds* one[1] = {ds};
  int32_t status = 0;
  int st = iak3d_nll_terms_batch(ctx, 1, one, pars, lndetA, WiAW, &status);
  if (st != 0) return st;
  if (status != 0) return status;
  if (iAW) st = iak_backsolve_to_host(ctx, ds->slots[0], iAW, nullptr);
  return st;
}

// ================================================================================================
// lndetANDinvCb on a caller-supplied matrix (optim_functions.R:268-312)
// ================================================================================================
int iak_factor_slot(iak3d_ctx* ctx, Slot* s, int32_t* status_out, double* lndet_out, double* G_out /* q*q or null */) {
  Batch* b = ctx->batch;
  if (b->pending) { ctx->err = "factor: a batch is pending"; return IAK3D_ERR_ARG; }
  int st = batch_reserve(ctx, b, 1, s->q);
  if (st != 0) return st;
  s->epoch++;
  std::vector<ProblemDesc> probs(1);
  fill_problem(&probs[0], s, b->d_status);
  std::vector<int> live(1, 1);
  st = batch_tasks(ctx, b, probs, live);
  if (st != 0) return st;
  CUDA_TRY(ctx, cudaMemsetAsync(b->d_status, 0, 2 * sizeof(int32_t), ctx->stream));
  CUDA_TRY(ctx, cudaMemcpyAsync(b->d_probs, probs.data(), sizeof(ProblemDesc), cudaMemcpyHostToDevice, ctx->stream));
  st = launch_tile_tasks(ctx, b->d_probs, b->d_tasks, b->ntasks, b->d_ticket, 0, /*latency_bound=*/1);
  if (st != 0) return st;
  st = launch_finish(ctx, b->d_probs, 1, b->d_results, b->stride);
  if (st != 0) return st;
  CUDA_TRY(ctx, cudaMemcpyAsync(b->h_results, b->d_results, (size_t)b->stride * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->stage_events_valid = false;
  const int flags = (int)b->h_results[1];
  if (flags & 2) { ctx->err = "factorisation: dependency wait timed out (internal error)"; return IAK3D_ERR_TIMEOUT; }
  *status_out = ((flags & 1) || !isfinite(b->h_results[0])) ? IAK3D_NOT_SPD : 0;
  if (lndet_out) *lndet_out = b->h_results[0];
  if (G_out) for (int e = 0; e < s->q * s->q; e++) G_out[e] = b->h_results[2 + e];
  return 0;
}

int iak_inverse_from_factor(iak3d_ctx* ctx, Slot* s, double* d_iC /* n x n */);   // gemm.cu

extern "C" int iak3d_lndet_invCb(iak3d_ctx* ctx, int64_t n, const double* C, int32_t nrhs, const double* b, double* lndetC,
                                 double* invCb) {
  if (!ctx || n <= 0 || !C || nrhs < 0 || (nrhs > 0 && !b)) return IAK3D_ERR_ARG;
  cudaSetDevice(ctx->device);
  double* dC = nullptr; double* db = nullptr; double* d_iC = nullptr;
  std::vector<Slot*> made;
  int rc = IAK3D_OK;
  const double nanv = nan("");
  do {
    if (cudaMalloc(&dC, (size_t)n * n * 8) != cudaSuccess) { ctx->err = "lndet_invCb: allocation failed"; rc = IAK3D_ERR_CUDA; cudaGetLastError(); break; }
    cudaMemcpyAsync(dC, C, (size_t)n * n * 8, cudaMemcpyHostToDevice, ctx->stream);
    if (nrhs > 0) {
      if (cudaMalloc(&db, (size_t)n * nrhs * 8) != cudaSuccess) { ctx->err = "lndet_invCb: allocation failed"; rc = IAK3D_ERR_CUDA; cudaGetLastError(); break; }
      cudaMemcpyAsync(db, b, (size_t)n * nrhs * 8, cudaMemcpyHostToDevice, ctx->stream);
    }
    // right-hand sides go through the bordered rows 64 at a time; the first pass also factorises
    const int npass = nrhs > 0 ? (nrhs + 63) / 64 : 1;
    for (int ps = 0; ps < npass && rc == 0; ps++) {
      const int q = nrhs > 0 ? std::min(64, nrhs - ps * 64) : 0;
      Slot* s = nullptr;
      rc = iak_alloc_slot(ctx, n, q, 0, 0, &s, false);
      if (rc != 0) break;
      made.push_back(s);
      rc = launch_pack_factor_layout(ctx, dC, n, db ? db + (size_t)ps * 64 * n : nullptr, q, s->Npad, s->ld, s->d_L);
      if (rc != 0) break;
      int32_t status = 0; double lndet = 0.0;
      rc = iak_factor_slot(ctx, s, &status, &lndet, nullptr);
      if (rc != 0) break;
      if (status != 0) {                                            // NA, NA (optim_functions.R:292-294)
        if (lndetC) *lndetC = nanv;
        rc = IAK3D_NOT_SPD;
        break;
      }
      if (lndetC) *lndetC = lndet;
      if (q > 0 && invCb) rc = iak_backsolve_to_host(ctx, s, invCb + (size_t)ps * 64 * n, nullptr);
      if (nrhs == 0 && invCb) {                                     // chol2inv (optim_functions.R:297)
        if (cudaMalloc(&d_iC, (size_t)n * n * 8) != cudaSuccess) { ctx->err = "lndet_invCb: allocation failed"; rc = IAK3D_ERR_CUDA; cudaGetLastError(); break; }
        rc = iak_inverse_from_factor(ctx, s, d_iC);
        if (rc != 0) break;
        cudaMemcpyAsync(invCb, d_iC, (size_t)n * n * 8, cudaMemcpyDeviceToHost, ctx->stream);
        if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { ctx->err = "lndet_invCb: copy failed"; rc = IAK3D_ERR_CUDA; }
      }
      if (ps + 1 < npass) { free_slot(s); made.pop_back(); }
    }
  } while (0);
  cudaStreamSynchronize(ctx->stream);
  for (Slot* s : made) free_slot(s);
  cudaFree(dC); cudaFree(db); cudaFree(d_iC);
  return rc;
}

// ================================================================================================
// measurement helpers
// ================================================================================================
extern "C" int iak3d_sync(iak3d_ctx* ctx) {
  if (!ctx) return IAK3D_ERR_ARG;
  cudaSetDevice(ctx->device);
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return IAK3D_OK;
}
extern "C" int iak3d_timer_start(iak3d_ctx* ctx) {
  if (!ctx) return IAK3D_ERR_ARG;
  cudaSetDevice(ctx->device);
  CUDA_TRY(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
  return IAK3D_OK;
}
extern "C" int iak3d_timer_stop(iak3d_ctx* ctx, double* ms) {
  if (!ctx || !ms) return IAK3D_ERR_ARG;
  cudaSetDevice(ctx->device);
  CUDA_TRY(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
  CUDA_TRY(ctx, cudaEventSynchronize(ctx->ev1));
  float f = 0;
  CUDA_TRY(ctx, cudaEventElapsedTime(&f, ctx->ev0, ctx->ev1));
  *ms = (double)f;
  return IAK3D_OK;
}
extern "C" int iak3d_flush_l2(iak3d_ctx* ctx) {
  if (!ctx) return IAK3D_ERR_ARG;
  cudaSetDevice(ctx->device);
  const size_t bytes = (size_t)256 << 20;
  if (!ctx->flush_buf) { CUDA_TRY(ctx, cudaMalloc(&ctx->flush_buf, bytes)); ctx->flush_bytes = bytes; }
  CUDA_TRY(ctx, cudaMemsetAsync(ctx->flush_buf, 0x5a, ctx->flush_bytes, ctx->stream));
  return IAK3D_OK;
}
extern "C" int iak3d_fp64_peak(iak3d_ctx* ctx, int32_t kind, double* tflops) {
  if (!ctx || !tflops || kind < 0 || kind > 2) return IAK3D_ERR_ARG;
  cudaSetDevice(ctx->device);
  return launch_peak(ctx, kind, tflops);
}
extern "C" int iak3d_last_task_profile(iak3d_ctx* ctx, int64_t* out, int64_t cap_tasks, int64_t* ntasks) {
  if (!ctx || !ntasks) return IAK3D_ERR_ARG;
  *ntasks = ctx->prof_ntasks;
  if (!out || ctx->prof_ntasks == 0) return IAK3D_OK;
  cudaSetDevice(ctx->device);
  const int64_t nt = std::min<int64_t>(cap_tasks, ctx->prof_ntasks);
  std::vector<int4> tasks((size_t)nt);
  std::vector<long long> tm((size_t)nt * 8);
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  CUDA_TRY(ctx, cudaMemcpy(tasks.data(), ctx->d_prof_tasks, (size_t)nt * sizeof(int4), cudaMemcpyDeviceToHost));
  CUDA_TRY(ctx, cudaMemcpy(tm.data(), ctx->d_prof, (size_t)nt * 8 * sizeof(long long), cudaMemcpyDeviceToHost));
  for (int64_t t = 0; t < nt; t++) {
    out[t * 11 + 0] = tasks[(size_t)t].x; out[t * 11 + 1] = tasks[(size_t)t].y; out[t * 11 + 2] = tasks[(size_t)t].z;
    for (int k = 0; k < 8; k++) out[t * 11 + 3 + k] = tm[(size_t)t * 8 + k];
  }
  return IAK3D_OK;
}

extern "C" int iak3d_last_stage_ms(iak3d_ctx* ctx, double* ms4, double* chol_flops, double* cov_bytes) {
  if (!ctx || !ms4) return IAK3D_ERR_ARG;
  if (!ctx->stage_events_valid) { ctx->err = "last_stage_ms: no batch has been run"; return IAK3D_ERR_ARG; }
  cudaSetDevice(ctx->device);
  CUDA_TRY(ctx, cudaEventSynchronize(ctx->evs[4]));
  for (int i = 0; i < 4; i++) {
    float f = 0;
    CUDA_TRY(ctx, cudaEventElapsedTime(&f, ctx->evs[i], ctx->evs[i + 1]));
    ms4[i] = (double)f;
  }
  if (chol_flops) *chol_flops = ctx->last_chol_flops;
  if (cov_bytes) *cov_bytes = ctx->last_cov_bytes;
  return IAK3D_OK;
}
