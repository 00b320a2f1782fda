// grid.cu — builds the uniform grid that replaces pcl::KdTreeFLANN on the device
// ([PCL] kdtree/impl/kdtree_flann.hpp : setInputCloud; SURVEY.md 8a-2 / 8a-2').
//
// Like KdTreeFLANN, non-finite points are left out and results are reported as ORIGINAL indices.
//   bbox        : finite min/max + count, one pass, float4 loads, ordered-int atomics
//   cell ids    : key = (z * dy + y) * dx + x (x fastest); non-finite points get key = n_cells; the same kernel
//                 counts the digits of every sort pass (sort_scan.cuh)
//   radix sort  : (key, original index), stable, one kernel per 8-bit digit (radix_sort.cu)
//   gather      : sorted float4 records with the original index in .w (and sorted normals)
//   cell_start  : every boundary between two sorted keys writes the table entries of the cells in between
//                 (warp-cooperative, coalesced): one streaming write of the table instead of a binary search per cell
// HBM traffic is N * (16 read + 16 write + 4 idx) + 4 * cells, plus the sort passes.
// Host round trips: the bounding box (the cell size and the table size depend on it) and, after everything has been
// queued, the measured occupancy (a rebuild with another cell size happens only for unusual clouds).
#include <algorithm>

#include "core_math.cuh"
#include "sort_scan.cuh"

namespace peb {

namespace {

__device__ __forceinline__ unsigned ordered_from_float(float f) {
  unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float float_from_ordered(unsigned u) {
  unsigned v = (u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u;
#ifdef __CUDA_ARCH__
  return __uint_as_float(v);
#else
  float f;
  memcpy(&f, &v, 4);
  return f;
#endif
}

// out[0..2] = min (ordered uint), out[3..5] = max, out[6] = finite count
__global__ void __launch_bounds__(256) bbox_kernel(const float4* __restrict__ pts, int n, unsigned* __restrict__ out) {
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  unsigned cnt = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float4 p = pts[i];
    if (finite3(p.x, p.y, p.z)) {
      mn[0] = fminf(mn[0], p.x);
      mn[1] = fminf(mn[1], p.y);
      mn[2] = fminf(mn[2], p.z);
      mx[0] = fmaxf(mx[0], p.x);
      mx[1] = fmaxf(mx[1], p.y);
      mx[2] = fmaxf(mx[2], p.z);
      ++cnt;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      mn[a] = fminf(mn[a], __shfl_xor_sync(0xFFFFFFFFu, mn[a], o));
      mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xFFFFFFFFu, mx[a], o));
    }
    cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, o);
  }
  __shared__ float smn[8][3], smx[8][3];
  __shared__ unsigned scnt[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    for (int a = 0; a < 3; ++a) {
      smn[warp][a] = mn[a];
      smx[warp][a] = mx[a];
    }
    scnt[warp] = cnt;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    const int a = threadIdx.x;
    float lo = smn[0][a], hi = smx[0][a];
    for (int w = 1; w < 8; ++w) {
      lo = fminf(lo, smn[w][a]);
      hi = fmaxf(hi, smx[w][a]);
    }
    atomicMin(&out[a], ordered_from_float(lo));
    atomicMax(&out[3 + a], ordered_from_float(hi));
  }
  if (threadIdx.x == 3) {
    unsigned c = 0;
    for (int w = 0; w < 8; ++w) c += scnt[w];
    atomicAdd(&out[6], c);
  }
}

__global__ void __launch_bounds__(256) cell_key_kernel(const float4* __restrict__ pts, int n, GridView g,
                                                       uint32_t n_cells, uint32_t* __restrict__ keys,
                                                       uint32_t* __restrict__ vals, int passes,
                                                       uint32_t* __restrict__ hist) {
  __shared__ uint32_t sh[kSortMaxPasses][kSortRadix];
  for (int i = threadIdx.x; i < kSortMaxPasses * kSortRadix; i += 256) (&sh[0][0])[i] = 0;
  __syncthreads();
  const int stride = gridDim.x * 256;
  const int rounds = (n + stride - 1) / stride;  // every lane runs every round (warp votes in sort_hist_add)
  for (int r = 0; r < rounds; ++r) {
    const int i = r * stride + blockIdx.x * 256 + threadIdx.x;
    const bool in = i < n;
    uint32_t key = n_cells;
    if (in) {
      const float4 p = pts[i];
      if (finite3(p.x, p.y, p.z)) {
        const int cx = grid_coord(p.x, g.ox, g.inv_h, g.dx);
        const int cy = grid_coord(p.y, g.oy, g.inv_h, g.dy);
        const int cz = grid_coord(p.z, g.oz, g.inv_h, g.dz);
        key = static_cast<uint32_t>((static_cast<long long>(cz) * g.dy + cy) * g.dx + cx);
      }
      keys[i] = key;
      vals[i] = static_cast<uint32_t>(i);
    }
    sort_hist_add(sh, key, in, passes);
  }
  __syncthreads();
  sort_hist_flush(sh, hist, passes);
}

__global__ void __launch_bounds__(256) gather_sorted_kernel(const float4* __restrict__ pts,
                                                            const float4* __restrict__ normals,
                                                            const uint32_t* __restrict__ vals, int n_finite,
                                                            float4* __restrict__ out_pts,
                                                            float4* __restrict__ out_normals) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_finite) return;
  const uint32_t src = vals[j];
  float4 p = pts[src];
  p.w = __int_as_float(static_cast<int>(src));
  out_pts[j] = p;
  if (normals) out_normals[j] = normals[src];
}

// cell_start[c] = first sorted position whose key >= c, for c in [0, n_cells].  Thread j looks at the boundary between
// sorted positions j - 1 and j (j = n_finite: the end): every cell id in (key[j-1], key[j]] starts at j.  Gaps of more
// than two cells are filled by the whole warp (consecutive cells -> consecutive addresses).  Also counts the occupied
// cells (*occupied += number of boundaries with key[j] != key[j-1]).
__global__ void __launch_bounds__(256) cell_start_fill_kernel(const uint32_t* __restrict__ sorted_keys, int n_finite,
                                                              uint32_t n_cells, uint32_t* __restrict__ cell_start,
                                                              unsigned* __restrict__ occupied) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const bool in = j <= n_finite;
  long long prev = -1, cur = -1;
  if (in) {
    prev = j > 0 ? static_cast<long long>(sorted_keys[j - 1]) : -1;
    cur = j < n_finite ? static_cast<long long>(sorted_keys[j]) : static_cast<long long>(n_cells);
  }
  const long long gap = cur - prev;
  const unsigned heads = __ballot_sync(0xFFFFFFFFu, in && j < n_finite && gap > 0);
  if (lane == 0 && heads) atomicAdd(occupied, static_cast<unsigned>(__popc(heads)));
  if (in && gap >= 1 && gap <= 2)
    for (long long c = prev + 1; c <= cur; ++c) cell_start[c] = static_cast<uint32_t>(j);
  unsigned big = __ballot_sync(0xFFFFFFFFu, in && gap > 2);
  while (big) {
    const int src = __ffs(big) - 1;
    big &= big - 1;
    const long long p = __shfl_sync(0xFFFFFFFFu, prev, src);
    const long long e = __shfl_sync(0xFFFFFFFFu, cur, src);
    const uint32_t jj = static_cast<uint32_t>(__shfl_sync(0xFFFFFFFFu, j, src));
    for (long long c = p + 1 + lane; c <= e; c += 32) cell_start[c] = jj;
  }
}

}  // namespace

int bbox_finite(peb_ctx* ctx, const float4* d_pts, int n, float mn[3], float mx[3], int* n_finite) {
  unsigned* d = ctx->d_small.as<unsigned>();
  unsigned* h = ctx->h_small.as<unsigned>();
  for (int a = 0; a < 3; ++a) {
    h[a] = 0xFFFFFFFFu;
    h[3 + a] = 0u;
  }
  h[6] = 0;
  PEB_CUDA(ctx, cudaMemcpyAsync(d, h, 7 * sizeof(unsigned), cudaMemcpyHostToDevice, ctx->stream));
  if (n > 0) {
    const int blocks = min(ceil_div(n, 256), kSmCount * 8);
    PEB_LAUNCH(ctx, bbox_kernel, blocks, 256, 0, d_pts, n, d);
  }
  PEB_CUDA(ctx, cudaMemcpyAsync(h, d, 7 * sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
  PEB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (int a = 0; a < 3; ++a) {
    mn[a] = float_from_ordered(h[a]);
    mx[a] = float_from_ordered(h[3 + a]);
  }
  *n_finite = static_cast<int>(h[6]);
  return PEB_OK;
}

// occupancy = wanted mean number of points per OCCUPIED cell.  The first guess assumes a
// surface-like cloud (area ~ product of the two largest extents); the build is repeated once or
// twice with the measured occupancy if the guess was off by more than 2x (volumetric or
// strongly folded clouds).
int grid_build(peb_ctx* ctx, Grid* g, const float4* d_pts, const float4* d_normals, int n, float occupancy) {
  g->valid = false;
  g->n_input = n;
  float mn[3], mx[3];
  int n_finite = 0;
  PEB_TRY(bbox_finite(ctx, d_pts, n, mn, mx, &n_finite));
  GridView& v = g->view;
  v = GridView{};
  v.n = n_finite;
  if (n_finite == 0) {
    v.ox = v.oy = v.oz = 0.0f;
    v.h = 1.0f;
    v.inv_h = 1.0f;
    v.dx = v.dy = v.dz = 1;
    g->n_cells = 1;
    PEB_CUDA(ctx, g->cell_start.ensure(2 * sizeof(uint32_t)));
    PEB_CUDA(ctx, cudaMemsetAsync(g->cell_start.p, 0, 2 * sizeof(uint32_t), ctx->stream));
    PEB_CUDA(ctx, g->pts.ensure(sizeof(float4)));
    v.pts = g->pts.as<float4>();
    v.normals = nullptr;
    v.cell_start = g->cell_start.as<uint32_t>();
    g->valid = true;
    return PEB_OK;
  }
  float ext[3] = {mx[0] - mn[0], mx[1] - mn[1], mx[2] - mn[2]};
  float e[3] = {ext[0], ext[1], ext[2]};
  if (e[0] < e[1]) { float t = e[0]; e[0] = e[1]; e[1] = t; }
  if (e[1] < e[2]) { float t = e[1]; e[1] = e[2]; e[2] = t; }
  if (e[0] < e[1]) { float t = e[0]; e[0] = e[1]; e[1] = t; }
  float maxabs = 0.0f;
  for (int a = 0; a < 3; ++a) maxabs = fmaxf(maxabs, fmaxf(fabsf(mn[a]), fabsf(mx[a])));
  float h_floor;
  float h = grid_initial_cell(e, n_finite, occupancy, maxabs, &h_floor);
  const long long kMaxCells = 1ll << 25;  // 32 M cells = 128 MB of cell_start at most
  PEB_CUDA(ctx, g->keys.ensure(static_cast<size_t>(n) * 4));
  PEB_CUDA(ctx, g->vals.ensure(static_cast<size_t>(n) * 4));
  PEB_CUDA(ctx, g->keys_tmp.ensure(static_cast<size_t>(n) * 4));
  PEB_CUDA(ctx, g->vals_tmp.ensure(static_cast<size_t>(n) * 4));
  PEB_CUDA(ctx, g->pts.ensure(static_cast<size_t>(n_finite) * sizeof(float4)));
  if (d_normals) PEB_CUDA(ctx, g->normals.ensure(static_cast<size_t>(n_finite) * sizeof(float4)));

  uint32_t *sk = nullptr, *sv = nullptr;
  for (int attempt = 0; attempt < 3; ++attempt) {
    long long dx, dy, dz, cells;
    for (;;) {
      dx = static_cast<long long>(floorf(ext[0] / h)) + 1;
      dy = static_cast<long long>(floorf(ext[1] / h)) + 1;
      dz = static_cast<long long>(floorf(ext[2] / h)) + 1;
      cells = dx * dy * dz;
      if (cells <= kMaxCells) break;
      h *= 1.26f;  // halves the cell count
    }
    v.ox = mn[0];
    v.oy = mn[1];
    v.oz = mn[2];
    v.h = h;
    v.inv_h = 1.0f / h;
    v.dx = static_cast<int>(dx);
    v.dy = static_cast<int>(dy);
    v.dz = static_cast<int>(dz);
    g->n_cells = cells;
    int key_bits = 1;
    while ((1ll << key_bits) <= cells) ++key_bits;  // keys run 0..cells (cells = non-finite)
    SortPlan plan;
    PEB_TRY(sort_prepare(ctx, n, key_bits, &plan));
    const int key_blocks = std::min(ceil_div(n, 256 * 8), kSmCount * 8);
    PEB_LAUNCH(ctx, cell_key_kernel, key_blocks, 256, 0, d_pts, n, v, static_cast<uint32_t>(cells),
               g->keys.as<uint32_t>(), g->vals.as<uint32_t>(), plan.passes, plan.hist);
    PEB_TRY(sort_pairs_counted(ctx, plan, g->keys.as<uint32_t>(), g->vals.as<uint32_t>(), g->keys_tmp.as<uint32_t>(),
                               g->vals_tmp.as<uint32_t>(), n, &sk, &sv));
    // the rest of the build is queued before the occupancy comes back: the usual cloud needs no second attempt
    PEB_CUDA(ctx, g->cell_start.ensure(static_cast<size_t>(g->n_cells + 1) * sizeof(uint32_t)));
    unsigned* d_occ = ctx->d_small.as<unsigned>() + 16;
    unsigned* h_occ = ctx->h_small.as<unsigned>() + 16;
    PEB_CUDA(ctx, cudaMemsetAsync(d_occ, 0, sizeof(unsigned), ctx->stream));
    PEB_LAUNCH(ctx, gather_sorted_kernel, ceil_div(n_finite, 256), 256, 0, d_pts, d_normals, sv, n_finite,
               g->pts.as<float4>(), d_normals ? g->normals.as<float4>() : nullptr);
    PEB_LAUNCH(ctx, cell_start_fill_kernel, ceil_div(n_finite + 1, 256), 256, 0, sk, n_finite,
               static_cast<uint32_t>(g->n_cells), g->cell_start.as<uint32_t>(), d_occ);
    PEB_CUDA(ctx, cudaMemcpyAsync(h_occ, d_occ, sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
    PEB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const float occ = static_cast<float>(n_finite) / static_cast<float>(*h_occ > 0 ? *h_occ : 1);
    if (attempt == 2 || (occ <= occupancy * 2.0f && occ >= occupancy * 0.5f) || h <= h_floor * 1.0001f && occ < occupancy)
      break;
    // occupancy ~ h^2 on a surface, ~ h^3 in a volume: the geometric mean exponent converges fast enough
    float scale = powf(occupancy / occ, 0.4f);
    float h_new = fmaxf(h * scale, h_floor);
    if (h_new < h && cells * 1.0 / (scale * scale * scale) > static_cast<double>(kMaxCells)) break;
    h = h_new;
  }
  v.pts = g->pts.as<float4>();
  v.normals = d_normals ? g->normals.as<float4>() : nullptr;
  v.cell_start = g->cell_start.as<uint32_t>();
  g->valid = true;
  return PEB_OK;
}

}  // namespace peb
