// normals.cu — pcl::NormalEstimation<PointXYZ, Normal>::computeFeature with setKSearch(k)
// ([PCL] features/impl/normal_3d.hpp, features/normal_3d.h, features/impl/feature.hpp;
//  SURVEY.md 8a-8) on the device.
//
// The k nearest neighbours (the point itself included, ascending squared distance, exactly
// what KdTreeFLANN::nearestKSearch returns) come from a uniform grid over the cloud whose cell
// edge is chosen so that the 3x3x3 block around a point usually already holds them.  One thread
// per point; its candidate list lives in shared memory, column per thread, kept sorted by insertion
// (64-bit (distance, position) keys: one compare / load / store per shift).  Points are processed in grid (cell) order, so the threads of a warp scan
// the same rows and hit L1.  The covariance is then accumulated over the neighbours IN LIST
// ORDER, sequentially, in float, without FMA — PCL 1.10's single-pass
// computeMeanAndCovarianceMatrix — followed by the closed-form eigen33 (core_math.cuh).
// Algorithmic HBM bytes: N * (16 read + 16 k gather + 32 write).
#include "nn_search.cuh"

namespace peb {

namespace {

// The candidate list of one thread: k entries, ascending.  An entry is the 64-bit key
// (squared distance bits << 32) | sorted position: squared distances are non-negative floats, so the integer order of
// the keys is the list's order (smaller distance first, then smaller position) and one 64-bit compare / load / store
// does the work of two.
struct KnnList {
  unsigned long long* e;  // [k][T], this thread's column
  int stride;             // T
  int k;
  int count;
  unsigned long long worst;  // key of the last entry once the list is full
};

__device__ __forceinline__ unsigned long long knn_key(float d2, int j) {
  return (static_cast<unsigned long long>(__float_as_uint(d2)) << 32) | static_cast<unsigned>(j);
}
__device__ __forceinline__ int knn_key_pos(unsigned long long key) { return static_cast<int>(key & 0xFFFFFFFFull); }

__device__ __forceinline__ void knn_insert(KnnList& l, float d2, int j) {
  const unsigned long long key = knn_key(d2, j);
  if (l.count == l.k && !(key < l.worst)) return;
  int pos = (l.count < l.k) ? l.count : l.k - 1;
  while (pos > 0) {
    const unsigned long long prev = l.e[(pos - 1) * l.stride];
    if (prev > key) {
      l.e[pos * l.stride] = prev;
      --pos;
    } else {
      break;
    }
  }
  l.e[pos * l.stride] = key;
  if (l.count < l.k) ++l.count;
  if (l.count == l.k) l.worst = l.e[(l.k - 1) * l.stride];
}

__device__ __forceinline__ void knn_scan_range(const GridView& g, uint32_t s, uint32_t e, float qx, float qy, float qz,
                                               KnnList& l) {
  for (uint32_t j = s; j < e; ++j) {
    const float4 p = g.pts[j];
    const float d2 = l2_simple(qx, qy, qz, p.x, p.y, p.z);
    knn_insert(l, d2, static_cast<int>(j));
  }
}

__device__ __forceinline__ void knn_scan_ring(const GridView& g, float qx, float qy, float qz, int cx, int cy, int cz,
                                              int r, bool full, KnnList& l) {
  const int w = 2 * r + 1;
  const int x0 = max(cx - r, 0), x1 = min(cx + r, g.dx - 1);
  for (int row = 0; row < w * w; ++row) {
    const int oy = row % w - r, oz = row / w - r;
    const int y = cy + oy, z = cz + oz;
    if (y < 0 || y >= g.dy || z < 0 || z >= g.dz) continue;
    const long long base = (static_cast<long long>(z) * g.dy + y) * g.dx;
    const bool outer = full || oy == -r || oy == r || oz == -r || oz == r;
    if (outer) {
      knn_scan_range(g, g.cell_start[base + x0], g.cell_start[base + x1 + 1], qx, qy, qz, l);
    } else {
      if (cx - r >= 0) knn_scan_range(g, g.cell_start[base + cx - r], g.cell_start[base + cx - r + 1], qx, qy, qz, l);
      if (cx + r < g.dx) knn_scan_range(g, g.cell_start[base + cx + r], g.cell_start[base + cx + r + 1], qx, qy, qz, l);
    }
  }
}

__global__ void normals_fill_nan_kernel(float* __restrict__ out8, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float qnan = __int_as_float(0x7fc00000);
  float4* o = reinterpret_cast<float4*>(out8 + 8 * static_cast<size_t>(i));
  o[0] = make_float4(qnan, qnan, qnan, 0.0f);
  o[1] = make_float4(qnan, 0.0f, 0.0f, 0.0f);
}

__global__ void normals_knn_kernel(const GridView g, int k, float vx, float vy, float vz, float* __restrict__ out8,
                                   int32_t* __restrict__ out_nn) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int T = blockDim.x;
  unsigned long long* se = reinterpret_cast<unsigned long long*>(smem_raw);
  const int q = blockIdx.x * T + threadIdx.x;
  if (q >= g.n) return;
  const float4 p = g.pts[q];
  const int orig = __float_as_int(p.w);
  KnnList l;
  l.e = se + threadIdx.x;
  l.stride = T;
  l.k = min(k, g.n);
  l.count = 0;
  l.worst = ~0ull;
  const int cx = grid_coord(p.x, g.ox, g.inv_h, g.dx);
  const int cy = grid_coord(p.y, g.oy, g.inv_h, g.dy);
  const int cz = grid_coord(p.z, g.oz, g.inv_h, g.dz);
  int r = 1;
  bool full = true;
  for (;;) {
    knn_scan_ring(g, p.x, p.y, p.z, cx, cy, cz, r, full, l);
    bool covers_all;
    const float b2 = grid_ring_bound2(g, p.x, p.y, p.z, cx, cy, cz, r, covers_all);
    if (covers_all || (l.count == l.k && __uint_as_float(static_cast<unsigned>(l.worst >> 32)) <= b2)) break;
    ++r;
    full = false;
  }
  float* o = out8 + 8 * static_cast<size_t>(orig);
  if (out_nn) {
    for (int i = 0; i < k; ++i)
      out_nn[static_cast<size_t>(orig) * k + i] = i < l.count ? __float_as_int(g.pts[knn_key_pos(l.e[i * T])].w) : -1;
  }
  if (l.count < 3) return;  // stays NaN ([PCL] normal_3d.h: computePointNormal fails below 3 points)
  float accu[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int i = 0; i < l.count; ++i) {
    const float4 c = g.pts[knn_key_pos(l.e[i * T])];
    accu[0] += c.x * c.x;
    accu[1] += c.x * c.y;
    accu[2] += c.x * c.z;
    accu[3] += c.y * c.y;
    accu[4] += c.y * c.z;
    accu[5] += c.z * c.z;
    accu[6] += c.x;
    accu[7] += c.y;
    accu[8] += c.z;
  }
  float res[8];
  normal_from_accu(accu, l.count, p.x, p.y, p.z, vx, vy, vz, res);
  reinterpret_cast<float4*>(o)[0] = make_float4(res[0], res[1], res[2], res[3]);
  reinterpret_cast<float4*>(o)[1] = make_float4(res[4], res[5], res[6], res[7]);
}

}  // namespace

int normals_knn_device(peb_ctx* ctx, const float4* d_in, int n, int k, const float vp[3], float* d_out8,
                       int32_t* d_out_nn) {
  if (k < 1) return fail(ctx, PEB_E_INVALID_ARG, "normals_knn: k must be >= 1 (got %d)", k);
  if (k > 192) return fail(ctx, PEB_E_UNSUPPORTED, "normals_knn: k = %d > 192 has no CUDA path (no CPU fallback)", k);
  if (n == 0) return PEB_OK;
  PEB_LAUNCH(ctx, normals_fill_nan_kernel, ceil_div(n, 256), 256, 0, d_out8, n);
  if (d_out_nn) PEB_CUDA(ctx, cudaMemsetAsync(d_out_nn, 0xFF, static_cast<size_t>(n) * k * sizeof(int32_t), ctx->stream));
  const float occupancy = fmaxf(2.0f, static_cast<float>(k) / 3.0f);
  PEB_TRY(grid_build(ctx, &ctx->aux_grid, d_in, nullptr, n, occupancy));
  const GridView& g = ctx->aux_grid.view;
  if (g.n == 0) return PEB_OK;
  const int T = k <= 48 ? 128 : (k <= 96 ? 64 : 32);
  const size_t smem = static_cast<size_t>(k) * T * 8;
  PEB_LAUNCH(ctx, normals_knn_kernel, ceil_div(g.n, T), T, smem, g, k, vp[0], vp[1], vp[2], d_out8, d_out_nn);
  return PEB_OK;
}

}  // namespace peb
