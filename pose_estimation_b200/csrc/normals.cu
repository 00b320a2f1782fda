// normals.cu — pcl::NormalEstimation<PointXYZ, Normal>::computeFeature with setKSearch(k)
// ([PCL] features/impl/normal_3d.hpp, features/normal_3d.h, features/impl/feature.hpp;
//  SURVEY.md 8a-8) on the device.
//
// The k nearest neighbours (the point itself included, ascending squared distance, exactly
// what KdTreeFLANN::nearestKSearch returns) come from a uniform grid over the cloud whose cell
// edge is chosen so that the 3x3x3 block around a point usually already holds them.
//
// k <= 32 (normals_warp_kernel): a WARP per query, 32 consecutive queries of the cell-sorted cloud per warp.  The
// candidates of the 3 x 3 x 3 block are read 32 at a time (one per lane, consecutive lanes = consecutive points of a grid
// row), a chunk that holds anything better than the current k-th neighbour is sorted with a warp bitonic network
// (64-bit (distance, position) keys in registers, shuffles) and merged into the sorted list the warp keeps one entry
// per lane.  The covariance is then accumulated over the neighbours IN LIST ORDER, sequentially, in float, without
// FMA — PCL 1.10's single-pass computeMeanAndCovarianceMatrix: its nine sums are nine independent chains, one per
// lane, fed by broadcasting neighbour after neighbour.  The closed-form eigen33 (core_math.cuh) of the warp's 32
// queries finally runs one query per lane.  A query whose k-th neighbour is not proven by the 3 x 3 x 3 block (sparse
// borders) goes to a list that the general kernel below finishes.
// k > 32, or leftovers (normals_knn_kernel): one thread per query; its candidate list lives in shared memory, column
// per thread, kept sorted by insertion, rings until the bound proves the list.
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

// list == nullptr: every sorted point is a query; else the queries are list[0 .. *list_n)
__global__ void normals_knn_kernel(const GridView g, int k, float vx, float vy, float vz, float* __restrict__ out8,
                                   int32_t* __restrict__ out_nn, const int* __restrict__ list,
                                   const int* __restrict__ list_n) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int T = blockDim.x;
  unsigned long long* se = reinterpret_cast<unsigned long long*>(smem_raw);
  int q = blockIdx.x * T + threadIdx.x;
  if (list) {
    if (q >= *list_n) return;
    q = list[q];
  }
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

// ---- k <= 32: a warp per query --------------------------------------------------------------------------------
constexpr int kNrmWarps = 4;  // warps per block

__device__ __forceinline__ unsigned long long warp_sort32(unsigned long long key, int lane) {
#pragma unroll
  for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
    for (int j = size >> 1; j > 0; j >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xFFFFFFFFu, key, j);
      const bool up = (lane & size) == 0;  // (size == 32: always ascending)
      const bool low = (lane & j) == 0;
      key = (low == up) ? (other < key ? other : key) : (other > key ? other : key);
    }
  }
  return key;
}

// best (ascending over the lanes) <- the 32 smallest of best and chunk (ascending over the lanes), ascending
__device__ __forceinline__ unsigned long long warp_merge32(unsigned long long best, unsigned long long chunk, int lane) {
  const unsigned long long rev = __shfl_sync(0xFFFFFFFFu, chunk, 31 - lane);
  unsigned long long m = rev < best ? rev : best;  // a bitonic sequence holding the 32 smallest
#pragma unroll
  for (int j = 16; j > 0; j >>= 1) {
    const unsigned long long other = __shfl_xor_sync(0xFFFFFFFFu, m, j);
    m = ((lane & j) == 0) ? (other < m ? other : m) : (other > m ? other : m);
  }
  return m;
}

__global__ void __launch_bounds__(32 * kNrmWarps) normals_warp_kernel(const GridView g, int k, float vx, float vy, float vz,
                                                                    float* __restrict__ out8, int32_t* __restrict__ out_nn,
                                                                    int* __restrict__ left, int* __restrict__ left_n) {
  __shared__ float s_accu[kNrmWarps][32][9];
  __shared__ uint32_t s_row_start[kNrmWarps][9];
  __shared__ uint32_t s_row_prefix[kNrmWarps][10];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int q0 = (blockIdx.x * kNrmWarps + warp) * 32;
  if (q0 >= g.n) return;
  const int nq = min(32, g.n - q0);
  // lane a < 9 owns covariance sum a: xx xy xz yy yz zz x y z = u * v with u, v picked from (x, y, z, 1)
  const int sel_u = lane < 3 ? 0 : (lane < 5 ? 1 : (lane == 5 ? 2 : lane - 6));
  const int sel_v = lane < 3 ? lane : (lane < 5 ? lane - 2 : (lane == 5 ? 2 : 3));
  bool mine_ok = false;  // lane q: query q0 + q was finished here
  int last_cx = -1, last_cy = -1, last_cz = -1;
  uint32_t total = 0;
  for (int qi = 0; qi < nq; ++qi) {
    const int q = q0 + qi;
    const float4 p = g.pts[q];
    const int cx = grid_coord(p.x, g.ox, g.inv_h, g.dx);
    const int cy = grid_coord(p.y, g.oy, g.inv_h, g.dy);
    const int cz = grid_coord(p.z, g.oz, g.inv_h, g.dz);
    if (cx != last_cx || cy != last_cy || cz != last_cz) {  // (warp-uniform) the 9 grid rows of the 3 x 3 x 3 block
      last_cx = cx;
      last_cy = cy;
      last_cz = cz;
      uint32_t rs = 0, len = 0;
      if (lane < 9) {
        const int y = cy + lane % 3 - 1, z = cz + lane / 3 - 1;
        if (y >= 0 && y < g.dy && z >= 0 && z < g.dz) {
          const long long base = (static_cast<long long>(z) * g.dy + y) * g.dx;
          rs = g.cell_start[base + max(cx - 1, 0)];
          len = g.cell_start[base + min(cx + 1, g.dx - 1) + 1] - rs;
        }
      }
      uint32_t inc = len;
#pragma unroll
      for (int o = 1; o < 16; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
        if (lane >= o) inc += t;
      }
      __syncwarp();
      if (lane < 9) {
        s_row_start[warp][lane] = rs;
        s_row_prefix[warp][lane] = inc - len;
      }
      if (lane == 8) s_row_prefix[warp][9] = inc;
      total = __shfl_sync(0xFFFFFFFFu, inc, 8);
      __syncwarp();
    }
    unsigned long long best = ~0ull;  // lane i: the i-th nearest so far
    unsigned long long kth = ~0ull;   // the k-th nearest so far (lane k - 1), known to every lane
    for (uint32_t c0 = 0; c0 < total; c0 += 32) {
      const uint32_t c = c0 + lane;
      unsigned long long key = ~0ull;
      if (c < total) {
        int r = 0;
#pragma unroll
        for (int t = 1; t < 9; ++t) r += (c >= s_row_prefix[warp][t]) ? 1 : 0;
        const uint32_t j = s_row_start[warp][r] + (c - s_row_prefix[warp][r]);
        const float4 t4 = g.pts[j];
        key = knn_key(l2_simple(p.x, p.y, p.z, t4.x, t4.y, t4.z), static_cast<int>(j));
      }
      if (!__any_sync(0xFFFFFFFFu, key < kth)) continue;
      best = warp_merge32(best, warp_sort32(key, lane), lane);
      kth = __shfl_sync(0xFFFFFFFFu, best, k - 1);
    }
    bool covers_all;
    const float b2 = grid_ring_bound2(g, p.x, p.y, p.z, cx, cy, cz, 1, covers_all);
    const bool full = kth != ~0ull;
    if (!(full && (covers_all || __uint_as_float(static_cast<unsigned>(kth >> 32)) <= b2))) {
      // not proven by the 3 x 3 x 3 block (or fewer than k points in it): the general kernel finishes this query
      if (lane == 0) left[atomicAdd(left_n, 1)] = q;
      continue;
    }
    // the neighbours in list order: lane i < k holds neighbour i
    float nx = 0.f, ny = 0.f, nz = 0.f;
    if (lane < k) {
      const float4 c4 = g.pts[knn_key_pos(best)];
      nx = c4.x;
      ny = c4.y;
      nz = c4.z;
      if (out_nn) out_nn[static_cast<size_t>(__float_as_int(p.w)) * k + lane] = __float_as_int(c4.w);
    }
    float acc = 0.0f;
    for (int i = 0; i < k; ++i) {
      const float x = __shfl_sync(0xFFFFFFFFu, nx, i), y = __shfl_sync(0xFFFFFFFFu, ny, i), z = __shfl_sync(0xFFFFFFFFu, nz, i);
      const float u = sel_u == 0 ? x : (sel_u == 1 ? y : z);
      const float v = sel_v == 0 ? x : (sel_v == 1 ? y : (sel_v == 2 ? z : 1.0f));
      acc += u * v;  // (a product with 1.0f is exact: sums 6..8 add the coordinate itself, like accu[6] += x)
    }
    if (lane < 9) s_accu[warp][qi][lane] = acc;
    if (lane == qi) mine_ok = true;
  }
  __syncwarp();
  if (mine_ok) {  // one query per lane: covariance -> eigen33 -> flip -> pcl::Normal
    const float4 p = g.pts[q0 + lane];
    float accu[9];
#pragma unroll
    for (int a = 0; a < 9; ++a) accu[a] = s_accu[warp][lane][a];
    float res[8];
    normal_from_accu(accu, k, p.x, p.y, p.z, vx, vy, vz, res);
    float4* o = reinterpret_cast<float4*>(out8 + 8 * static_cast<size_t>(__float_as_int(p.w)));
    o[0] = make_float4(res[0], res[1], res[2], res[3]);
    o[1] = make_float4(res[4], res[5], res[6], res[7]);
  }
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
  if (k <= 32 && k >= 3 && g.n >= k) {
    // queries the warp kernel could not prove: sorted positions + their count (index 0)
    PEB_CUDA(ctx, ctx->nrm_left.ensure((static_cast<size_t>(g.n) + 1) * sizeof(int)));
    int* left_n = ctx->nrm_left.as<int>();
    int* left = left_n + 1;
    PEB_CUDA(ctx, cudaMemsetAsync(left_n, 0, sizeof(int), ctx->stream));
    PEB_LAUNCH(ctx, normals_warp_kernel, ceil_div(g.n, 32 * kNrmWarps), 32 * kNrmWarps, 0, g, k, vp[0], vp[1], vp[2],
               d_out8, d_out_nn, left, left_n);
    // (sized for the worst case; blocks beyond the count return at once)
    PEB_LAUNCH(ctx, normals_knn_kernel, ceil_div(g.n, T), T, smem, g, k, vp[0], vp[1], vp[2], d_out8, d_out_nn, left, left_n);
    return PEB_OK;
  }
  PEB_LAUNCH(ctx, normals_knn_kernel, ceil_div(g.n, T), T, smem, g, k, vp[0], vp[1], vp[2], d_out8, d_out_nn,
             static_cast<const int*>(nullptr), static_cast<const int*>(nullptr));
  return PEB_OK;
}

}  // namespace peb
