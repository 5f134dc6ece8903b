// voronoi.cu -- block construction for the composite likelihood: the data-parallel parts of setVoronoiBlocks
// (compositeLikIAK3D.R:829-935), getBlocksFromLocations (:803-812) and calcnPerCluster (:1033-1046, the objective that
// balanceNeighbours2's random search evaluates 15,000+ times, :1048-1133), plus the two host algorithms around them: the greedy
// centroid seeding (:853-884) and the Dirichlet-tile adjacency the reference takes from deldir (:907-916).
//
//   nearest_centroid_kernel   D = xyDist(x, vcentres) (iaCovMatern.R:888-920) row by row and which.min (first minimum): one
//                             thread per location, centroids in shared memory; optional per-centroid counts (integer
//                             atomics: order-independent, so deterministic)
// Compiled with -fmad=false: the distance is sqrt(dx^2 + dy^2) rounded like R's, so ties and near-ties fall the same way.
#include <math.h>

#include <algorithm>
#include <map>
#include <numeric>
#include <set>

#include "common.cuh"

struct iak3d_points {
  iak3d_ctx* ctx = nullptr;
  int64_t m = 0;
  double* d_xy = nullptr;          // m x 2 column-major
  int32_t* d_blk = nullptr;        // m
  double* d_cen = nullptr; int32_t cen_cap = 0;
  unsigned long long* d_cnt = nullptr;
};

namespace {

__global__ void __launch_bounds__(256) nearest_centroid_kernel(const double* __restrict__ xy, int64_t m, const double* __restrict__ cen,
                                                               int nB, int32_t* __restrict__ blk, unsigned long long* __restrict__ cnt) {
  extern __shared__ double sc[];                 // nB x 2 column-major
  for (int k = threadIdx.x; k < 2 * nB; k += blockDim.x) sc[k] = cen[k];
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const double x = xy[i], y = xy[m + i];
  double best = INFINITY; int bi = 0;
  bool any = false;
  for (int b = 0; b < nB; b++) {
    const double dx = x - sc[b], dy = y - sc[nB + b];
    const double d = sqrt(dx * dx + dy * dy);
    // which.min: first minimum, NaN skipped (a block without a centroid never wins)
    if (!isnan(d) && (!any || d < best)) { best = d; bi = b; any = true; }
  }
  blk[i] = any ? bi : -1;
  if (cnt && any) atomicAdd(cnt + bi, 1ull);
}

}  // namespace

extern "C" int iak3d_points_create(iak3d_ctx* ctx, int64_t m, const double* xy, iak3d_points** out) {
  if (!ctx || m < 0 || (m > 0 && !xy) || !out) return IAK3D_ERR_ARG;
  *out = nullptr;
  cudaSetDevice(ctx->device);
  iak3d_points* P = new iak3d_points();
  P->ctx = ctx; P->m = m;
  if (m > 0) {
    cudaError_t e = cudaMalloc(&P->d_xy, (size_t)m * 16);
    if (e == cudaSuccess) e = cudaMalloc(&P->d_blk, (size_t)m * 4);
    if (e == cudaSuccess) e = cudaMemcpyAsync(P->d_xy, xy, (size_t)m * 16, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
      ctx->err = std::string("points_create: ") + cudaGetErrorString(e); cudaGetLastError();
      cudaFree(P->d_xy); cudaFree(P->d_blk); delete P; return IAK3D_ERR_CUDA;
    }
  }
  *out = P;
  return IAK3D_OK;
}

extern "C" int iak3d_points_destroy(iak3d_points* P) {
  if (!P) return IAK3D_OK;
  cudaSetDevice(P->ctx->device);
  cudaStreamSynchronize(P->ctx->stream);
  cudaFree(P->d_xy); cudaFree(P->d_blk); cudaFree(P->d_cen); cudaFree(P->d_cnt);
  delete P;
  return IAK3D_OK;
}

extern "C" int iak3d_points_nearest(iak3d_points* P, int32_t nB, const double* vcentres, int32_t* block, int64_t* counts) {
  if (!P || nB < 1 || !vcentres) return IAK3D_ERR_ARG;
  iak3d_ctx* ctx = P->ctx;
  if (nB > 2800) { ctx->err = "points_nearest: at most 2800 centroids (shared-memory staging)"; return IAK3D_ERR_ARG; }
  cudaSetDevice(ctx->device);
  if (nB > P->cen_cap) {
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(P->d_cen); cudaFree(P->d_cnt); P->d_cen = nullptr; P->d_cnt = nullptr;
    CUDA_TRY(ctx, cudaMalloc(&P->d_cen, (size_t)nB * 16));
    CUDA_TRY(ctx, cudaMalloc(&P->d_cnt, (size_t)nB * 8));
    P->cen_cap = nB;
  }
  CUDA_TRY(ctx, cudaMemcpyAsync(P->d_cen, vcentres, (size_t)nB * 16, cudaMemcpyHostToDevice, ctx->stream));
  if (counts) CUDA_TRY(ctx, cudaMemsetAsync(P->d_cnt, 0, (size_t)nB * 8, ctx->stream));
  if (P->m > 0) {
    nearest_centroid_kernel<<<(unsigned)((P->m + 255) / 256), 256, (size_t)nB * 16, ctx->stream>>>(P->d_xy, P->m, P->d_cen, nB, P->d_blk,
                                                                                                counts ? P->d_cnt : nullptr);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
  }
  if (block && P->m > 0) CUDA_TRY(ctx, cudaMemcpyAsync(block, P->d_blk, (size_t)P->m * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (counts) CUDA_TRY(ctx, cudaMemcpyAsync(counts, P->d_cnt, (size_t)nB * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return IAK3D_OK;
}

// ---- greedy seeding of the centroids (compositeLikIAK3D.R:840-884) ---------------------------------------------------
// nBlocks = floor(nU / nPerBlock); the first nBigBlocks blocks take nPerBlock + 1 locations, the last one the rest.  Every
// step starts from the extreme location of the remaining set along its longer axis (which.max: first maximum), takes the k
// nearest remaining locations (order(): ties by position) and records their mean (colMeans accumulates in long double).
extern "C" int iak3d_voronoi_seed(int64_t nU, const double* xU, int32_t nPerBlock, int32_t cap_blocks, int32_t* nBlocks_out, double* vcentres) {
  if (nU < 1 || !xU || nPerBlock < 1 || !nBlocks_out || !vcentres) return IAK3D_ERR_ARG;
  const int64_t nBlocks = nU / nPerBlock;
  *nBlocks_out = (int32_t)nBlocks;
  if (nBlocks < 1 || nBlocks > cap_blocks) return IAK3D_ERR_ARG;
  const int64_t nSmall = nBlocks * ((int64_t)nPerBlock + 1) - nU;
  const int64_t nBig = nBlocks - nSmall;
  std::vector<int64_t> rem((size_t)nU);
  std::iota(rem.begin(), rem.end(), (int64_t)0);
  std::vector<double> d;
  std::vector<int64_t> ord;
  for (int64_t b = 0; b < nBlocks; b++) {
    const int64_t nr = (int64_t)rem.size();
    if (nr == 0) { vcentres[b] = NAN; vcentres[nBlocks + b] = NAN; continue; }
    double xmin = INFINITY, xmax = -INFINITY, ymin = INFINITY, ymax = -INFINITY;
    int64_t ix = 0, iy = 0;
    for (int64_t k = 0; k < nr; k++) {
      const double x = xU[rem[k]], y = xU[nU + rem[k]];
      if (x > xmax) { xmax = x; ix = k; }
      if (y > ymax) { ymax = y; iy = k; }
      xmin = std::min(xmin, x); ymin = std::min(ymin, y);
    }
    const int64_t it = ((ymax - ymin) > (xmax - xmin)) ? iy : ix;
    const double px = xU[rem[it]], py = xU[nU + rem[it]];
    d.resize((size_t)nr); ord.resize((size_t)nr);
    for (int64_t k = 0; k < nr; k++) {
      const double dx = xU[rem[k]] - px, dy = xU[nU + rem[k]] - py;
      volatile double sx = dx * dx, sy = dy * dy;      // no contraction: (dx)^2 + (dy)^2 as two rounded products
      d[k] = sqrt(sx + sy);
      ord[k] = k;
    }
    int64_t take = (b < nBig) ? std::min<int64_t>(nPerBlock + 1, nr) : (b < nBlocks - 1 ? std::min<int64_t>(nPerBlock, nr) : nr);
    auto cmp = [&](int64_t a, int64_t c) { return d[a] < d[c] || (d[a] == d[c] && a < c); };
    if (take < nr) std::partial_sort(ord.begin(), ord.begin() + take, ord.end(), cmp);
    else std::sort(ord.begin(), ord.end(), cmp);
    long double sx = 0.0L, sy = 0.0L;
    for (int64_t k = 0; k < take; k++) { sx += (long double)xU[rem[ord[k]]]; sy += (long double)xU[nU + rem[ord[k]]]; }
    vcentres[b] = (double)(sx / (long double)take);
    vcentres[nBlocks + b] = (double)(sy / (long double)take);
    std::vector<char> gone((size_t)nr, 0);
    for (int64_t k = 0; k < take; k++) gone[(size_t)ord[k]] = 1;
    std::vector<int64_t> next;
    next.reserve((size_t)(nr - take));
    for (int64_t k = 0; k < nr; k++) if (!gone[(size_t)k]) next.push_back(rem[k]);
    rem.swap(next);
  }
  return IAK3D_OK;
}

// ---- adjacency of the Dirichlet tiles (deldir(x, y)$dirsgs, compositeLikIAK3D.R:907-916) ---------------------------------
// Bowyer-Watson Delaunay triangulation of the centroids; two centroids are adjacent when the Dirichlet (Voronoi) edge
// between them -- the segment between the circumcentres of the two triangles on their Delaunay edge, or the ray from the
// one circumcentre of a hull edge -- meets deldir's default window: the bounding box of the points extended by 10 % on
// every side.  Duplicated centroids are collapsed onto their first occurrence (deldir drops duplicates).
namespace {
struct Tri { int a, b, c; double cx, cy, r2; bool dead; };
static void circum(const std::vector<double>& X, const std::vector<double>& Y, Tri& t) {
  const double ax = X[t.a], ay = Y[t.a], bx = X[t.b] - ax, by = Y[t.b] - ay, cx = X[t.c] - ax, cy = Y[t.c] - ay;
  const double dd = 2.0 * (bx * cy - by * cx);
  const double b2 = bx * bx + by * by, c2 = cx * cx + cy * cy;
  const double ux = (cy * b2 - by * c2) / dd, uy = (bx * c2 - cx * b2) / dd;
  t.cx = ax + ux; t.cy = ay + uy; t.r2 = ux * ux + uy * uy;
}
// does {p + t d : t in [0, tmax]} meet the closed rectangle?
static bool meets_window(double px, double py, double dx, double dy, double tmax, const double rw[4]) {
  double t0 = 0.0, t1 = tmax;
  const double pp[4] = {-dx, dx, -dy, dy};
  const double qq[4] = {px - rw[0], rw[1] - px, py - rw[2], rw[3] - py};
  for (int k = 0; k < 4; k++) {
    if (pp[k] == 0.0) { if (qq[k] < 0.0) return false; continue; }
    const double r = qq[k] / pp[k];
    if (pp[k] < 0.0) { if (r > t1) return false; if (r > t0) t0 = r; }
    else { if (r < t0) return false; if (r < t1) t1 = r; }
  }
  return t0 <= t1;
}
}  // namespace

extern "C" int iak3d_dirichlet_adjacency(int32_t nB, const double* vcentres, int32_t* pairs, int64_t cap_pairs, int64_t* npairs) {
  if (nB < 1 || !vcentres || !npairs) return IAK3D_ERR_ARG;
  *npairs = 0;
  // de-duplicate (first occurrence kept); NaN centroids (blocks that lost every location) take no part
  std::vector<int> orig;
  std::vector<double> X, Y;
  {
    std::map<std::pair<double, double>, int> seen;
    for (int i = 0; i < nB; i++) {
      const double x = vcentres[i], y = vcentres[nB + i];
      if (std::isnan(x) || std::isnan(y)) continue;
      if (seen.emplace(std::make_pair(x, y), i).second) { orig.push_back(i); X.push_back(x); Y.push_back(y); }
    }
  }
  const int n = (int)X.size();
  std::set<std::pair<int, int>> adj;
  auto emit = [&](int i, int j) { const int a = orig[i], b = orig[j]; adj.insert({std::min(a, b), std::max(a, b)}); };
  if (n == 2) emit(0, 1);
  if (n >= 3) {
    double xmin = X[0], xmax = X[0], ymin = Y[0], ymax = Y[0];
    for (int i = 1; i < n; i++) { xmin = std::min(xmin, X[i]); xmax = std::max(xmax, X[i]); ymin = std::min(ymin, Y[i]); ymax = std::max(ymax, Y[i]); }
    const double rw[4] = {xmin - 0.1 * (xmax - xmin), xmax + 0.1 * (xmax - xmin), ymin - 0.1 * (ymax - ymin), ymax + 0.1 * (ymax - ymin)};
    // super-triangle far outside the data
    const double span = std::max(xmax - xmin, ymax - ymin), mx = 0.5 * (xmin + xmax), my = 0.5 * (ymin + ymax);
    const double big = 1e5 * (span > 0 ? span : 1.0);
    X.push_back(mx - 2 * big); Y.push_back(my - big);
    X.push_back(mx + 2 * big); Y.push_back(my - big);
    X.push_back(mx); Y.push_back(my + 2 * big);
    std::vector<Tri> T;
    { Tri t{n, n + 1, n + 2, 0, 0, 0, false}; circum(X, Y, t); T.push_back(t); }
    for (int p = 0; p < n; p++) {
      std::map<std::pair<int, int>, int> edge_count;
      std::vector<std::pair<int, int>> edges;
      for (auto& t : T) {
        if (t.dead) continue;
        const double dx = X[p] - t.cx, dy = Y[p] - t.cy;
        if (dx * dx + dy * dy < t.r2) {
          t.dead = true;
          const int v[3] = {t.a, t.b, t.c};
          for (int e = 0; e < 3; e++) {
            const int u = v[e], w = v[(e + 1) % 3];
            edges.push_back({u, w});
            edge_count[{std::min(u, w), std::max(u, w)}]++;
          }
        }
      }
      std::vector<Tri> keep;
      keep.reserve(T.size() + 8);
      for (auto& t : T) if (!t.dead) keep.push_back(t);
      for (auto& e : edges)
        if (edge_count[{std::min(e.first, e.second), std::max(e.first, e.second)}] == 1) {
          Tri t{e.first, e.second, p, 0, 0, 0, false};
          // keep counter-clockwise orientation
          const double o = (X[t.b] - X[t.a]) * (Y[t.c] - Y[t.a]) - (Y[t.b] - Y[t.a]) * (X[t.c] - X[t.a]);
          if (o == 0.0) continue;
          if (o < 0) std::swap(t.a, t.b);
          circum(X, Y, t);
          keep.push_back(t);
        }
      T.swap(keep);
    }
    // real triangles and, per Delaunay edge, the triangles on it
    std::map<std::pair<int, int>, std::vector<int>> on_edge;
    std::vector<Tri> R;
    for (auto& t : T) if (t.a < n && t.b < n && t.c < n) R.push_back(t);
    for (int k = 0; k < (int)R.size(); k++) {
      const int v[3] = {R[k].a, R[k].b, R[k].c};
      for (int e = 0; e < 3; e++) on_edge[{std::min(v[e], v[(e + 1) % 3]), std::max(v[e], v[(e + 1) % 3])}].push_back(k);
    }
    if (R.empty()) {
      // collinear centroids: consecutive points along the line are neighbours
      std::vector<int> o(n);
      std::iota(o.begin(), o.end(), 0);
      std::sort(o.begin(), o.end(), [&](int a, int b) { return X[a] < X[b] || (X[a] == X[b] && Y[a] < Y[b]); });
      for (int k = 0; k + 1 < n; k++) emit(o[k], o[k + 1]);
    }
    for (auto& kv : on_edge) {
      const int i = kv.first.first, j = kv.first.second;
      const std::vector<int>& ts = kv.second;
      bool hit;
      if (ts.size() >= 2) {
        const Tri &t1 = R[ts[0]], &t2 = R[ts[1]];
        hit = meets_window(t1.cx, t1.cy, t2.cx - t1.cx, t2.cy - t1.cy, 1.0, rw);
      } else {
        // hull edge: the ray leaves the circumcentre perpendicular to the edge, away from the triangle's third vertex
        const Tri& t = R[ts[0]];
        const int k3 = (t.a != i && t.a != j) ? t.a : ((t.b != i && t.b != j) ? t.b : t.c);
        double nx = Y[j] - Y[i], ny = -(X[j] - X[i]);
        const double mxe = 0.5 * (X[i] + X[j]), mye = 0.5 * (Y[i] + Y[j]);
        if (nx * (X[k3] - mxe) + ny * (Y[k3] - mye) > 0) { nx = -nx; ny = -ny; }
        hit = meets_window(t.cx, t.cy, nx, ny, INFINITY, rw);
      }
      if (hit) emit(i, j);
    }
  }
  *npairs = (int64_t)adj.size();
  if (pairs) {
    if ((int64_t)adj.size() > cap_pairs) return IAK3D_ERR_ARG;
    int64_t k = 0;
    for (auto& e : adj) { pairs[2 * k] = e.first; pairs[2 * k + 1] = e.second; k++; }
  }
  return IAK3D_OK;
}
