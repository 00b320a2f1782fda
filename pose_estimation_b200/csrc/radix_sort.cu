// radix_sort.cu — stable LSD radix sort of (uint32 key, uint32 value) pairs and a device-wide
// exclusive scan.  These feed VoxelGrid (sort by voxel id, [PCL] filters/impl/voxel_grid.hpp
// "second pass") and the uniform-grid build that replaces the kd-tree (SURVEY.md 8a-2').
//
// Layout: 8-bit digits, only ceil(key_bits / 8) passes.  One pass = three launches:
//   digit_histogram : each CTA histograms its 4096-key tile                     (16 B / lane loads)
//   exclusive scan  : over the digit-major (digit, CTA) table -> global offsets
//   scatter         : each CTA re-reads its tile, ranks keys stably with warp match_any and a
//                     per-warp digit counter table in shared memory, writes to out[offset+rank]
// Stability keeps equal keys in ascending original index, which is what fixes the float
// summation order of the voxel centroids.  Sort traffic is overhead on top of the
// algorithmic bytes of its callers; at 2.3 M pairs every pass stays inside the 126 MB L2.
#include "common.cuh"

namespace peb {

namespace {

constexpr int kSortThreads = 256;
constexpr int kSortItems = 16;
constexpr int kSortTile = kSortThreads * kSortItems;  // 4096 keys per CTA
constexpr int kWarps = kSortThreads / 32;
constexpr int kRadix = 256;

__global__ void __launch_bounds__(kSortThreads) digit_histogram_kernel(const uint32_t* __restrict__ keys, int n,
                                                                       int shift, uint32_t* __restrict__ hist,
                                                                       int n_tiles) {
  __shared__ uint32_t sh[kRadix];
  sh[threadIdx.x] = 0;
  __syncthreads();
  const int tile = blockIdx.x;
  const int base = tile * kSortTile;
  // 4 x uint4 per thread, coalesced 16-byte loads
  const int end = min(base + kSortTile, n);
  for (int i = base + threadIdx.x * 4; i < end; i += kSortThreads * 4) {
    if (i + 3 < end) {
      uint4 k = *reinterpret_cast<const uint4*>(keys + i);
      atomicAdd(&sh[(k.x >> shift) & 0xFF], 1u);
      atomicAdd(&sh[(k.y >> shift) & 0xFF], 1u);
      atomicAdd(&sh[(k.z >> shift) & 0xFF], 1u);
      atomicAdd(&sh[(k.w >> shift) & 0xFF], 1u);
    } else {
      for (int j = i; j < end; ++j) atomicAdd(&sh[(keys[j] >> shift) & 0xFF], 1u);
    }
  }
  __syncthreads();
  hist[static_cast<size_t>(threadIdx.x) * n_tiles + tile] = sh[threadIdx.x];
}

__global__ void __launch_bounds__(kSortThreads) scatter_kernel(const uint32_t* __restrict__ keys_in,
                                                               const uint32_t* __restrict__ vals_in,
                                                               uint32_t* __restrict__ keys_out,
                                                               uint32_t* __restrict__ vals_out, int n, int shift,
                                                               const uint32_t* __restrict__ offsets, int n_tiles) {
  __shared__ uint32_t warp_cnt[kWarps][kRadix];  // per-warp running digit counters
  __shared__ uint32_t digit_base[kRadix];        // global offset of (digit, this tile)
  const int tile = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < kWarps * kRadix; i += kSortThreads) (&warp_cnt[0][0])[i] = 0;
  digit_base[threadIdx.x] = offsets[static_cast<size_t>(threadIdx.x) * n_tiles + tile];
  __syncthreads();

  // warp w owns the contiguous chunk [base + w*512, +512): item r of lane l is element r*32 + l,
  // so consecutive rounds and lanes walk the chunk in index order (stability).
  const int chunk = tile * kSortTile + warp * (32 * kSortItems);
  uint32_t key[kSortItems], val[kSortItems], rank[kSortItems];
#pragma unroll
  for (int r = 0; r < kSortItems; ++r) {
    const int i = chunk + r * 32 + lane;
    const bool ok = i < n;
    key[r] = ok ? keys_in[i] : 0xFFFFFFFFu;
    val[r] = ok ? vals_in[i] : 0u;
    const uint32_t digit = (key[r] >> shift) & 0xFF;
    // lanes past the end take part in the ballot with a digit nobody else can rank against
    const uint32_t active = __ballot_sync(0xFFFFFFFFu, ok);
    uint32_t peers = __match_any_sync(0xFFFFFFFFu, ok ? digit : (0x100u + lane)) & active;
    const uint32_t lower = peers & ((1u << lane) - 1u);
    uint32_t before = 0;
    if (ok) before = warp_cnt[warp][digit];
    __syncwarp();
    if (ok && lower == 0) warp_cnt[warp][digit] = before + __popc(peers);  // first peer bumps the counter
    __syncwarp();
    rank[r] = before + __popc(lower);
  }
  __syncthreads();
  // exclusive scan over warps, per digit: warp_cnt[w][d] becomes the number of keys with digit d in warps < w
  {
    const int d = threadIdx.x;
    uint32_t run = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
      uint32_t c = warp_cnt[w][d];
      warp_cnt[w][d] = run;
      run += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < kSortItems; ++r) {
    const int i = chunk + r * 32 + lane;
    if (i < n) {
      const uint32_t digit = (key[r] >> shift) & 0xFF;
      const uint32_t dst = digit_base[digit] + warp_cnt[warp][digit] + rank[r];
      keys_out[dst] = key[r];
      vals_out[dst] = val[r];
    }
  }
}

// ---- exclusive scan (three phases) ------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* total) {
  __shared__ uint32_t warp_sums[kScanThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_sums[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = lane < kScanThreads / 32 ? warp_sums[lane] : 0;
    uint32_t winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t t = __shfl_up_sync(0xFFFFFFFFu, winc, o);
      if (lane >= o) winc += t;
    }
    if (lane < kScanThreads / 32) warp_sums[lane] = winc - w;
    if (lane == 31 && total) *total = winc;
  }
  __syncthreads();
  uint32_t res = warp_sums[warp] + inc - v;
  __syncthreads();
  return res;
}

__global__ void __launch_bounds__(kScanThreads) scan_tile_sums_kernel(const uint32_t* __restrict__ in, int n,
                                                                      uint32_t* __restrict__ tile_sums) {
  __shared__ uint32_t total;
  const int base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i)
    if (base + i < n) s += in[base + i];
  block_exclusive_scan(s, &total);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// single CTA: exclusive scan of the tile sums in place; writes the grand total
__global__ void __launch_bounds__(kScanThreads) scan_spine_kernel(uint32_t* __restrict__ tile_sums, int n_tiles,
                                                                  uint32_t* __restrict__ d_total) {
  __shared__ uint32_t total;
  uint32_t carry = 0;
  for (int base = 0; base < n_tiles; base += kScanThreads) {
    const int i = base + threadIdx.x;
    uint32_t v = i < n_tiles ? tile_sums[i] : 0;
    uint32_t ex = block_exclusive_scan(v, &total);
    if (i < n_tiles) tile_sums[i] = carry + ex;
    carry += total;
    __syncthreads();
  }
  if (threadIdx.x == 0 && d_total) *d_total = carry;
}

__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(const uint32_t* __restrict__ in,
                                                                  uint32_t* __restrict__ out, int n,
                                                                  const uint32_t* __restrict__ tile_sums) {
  const int base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
  uint32_t v[kScanItems];
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    v[i] = (base + i < n) ? in[base + i] : 0;
    s += v[i];
  }
  uint32_t ex = block_exclusive_scan(s, nullptr) + tile_sums[blockIdx.x];
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    if (base + i < n) out[base + i] = ex;
    ex += v[i];
  }
}

}  // namespace

// out may alias in.  d_total (nullable) receives the sum of all elements.
int exclusive_scan_u32(peb_ctx* ctx, const uint32_t* in, uint32_t* out, int n, uint32_t* d_total) {
  if (n <= 0) {
    if (d_total) PEB_CUDA(ctx, cudaMemsetAsync(d_total, 0, sizeof(uint32_t), ctx->stream));
    return PEB_OK;
  }
  const int n_tiles = ceil_div(n, kScanTile);
  // tile sums live at the tail of d_scratch so that callers may keep using its head
  static_assert(sizeof(uint32_t) == 4, "");
  PEB_CUDA(ctx, ctx->d_small.ensure(1 << 20));
  uint32_t* tile_sums = ctx->d_small.as<uint32_t>() + (1 << 16);  // d_small: [0,256KB) results, [256KB, ...) spine
  if (static_cast<size_t>(n_tiles) * 4 + (1 << 18) > ctx->d_small.cap)
    return fail(ctx, PEB_E_INVALID_ARG, "exclusive_scan_u32: %d elements exceed the scan spine", n);
  PEB_LAUNCH(ctx, scan_tile_sums_kernel, n_tiles, kScanThreads, 0, in, n, tile_sums);
  PEB_LAUNCH(ctx, scan_spine_kernel, 1, kScanThreads, 0, tile_sums, n_tiles, d_total);
  PEB_LAUNCH(ctx, scan_apply_kernel, n_tiles, kScanThreads, 0, in, out, n, tile_sums);
  return PEB_OK;
}

// Sorts n pairs by the low key_bits bits of the key; ping-pongs between the two buffer pairs and
// reports where the result landed.
int sort_pairs(peb_ctx* ctx, uint32_t* keys, uint32_t* vals, uint32_t* keys_tmp, uint32_t* vals_tmp, int n,
               int key_bits, uint32_t** keys_out, uint32_t** vals_out) {
  *keys_out = keys;
  *vals_out = vals;
  if (n <= 1) return PEB_OK;
  const int passes = (key_bits + 7) / 8;
  const int n_tiles = ceil_div(n, kSortTile);
  const size_t hist_bytes = static_cast<size_t>(kRadix) * n_tiles * sizeof(uint32_t);
  PEB_CUDA(ctx, ctx->d_scratch.ensure(hist_bytes));
  uint32_t* hist = ctx->d_scratch.as<uint32_t>();
  uint32_t *ki = keys, *vi = vals, *ko = keys_tmp, *vo = vals_tmp;
  for (int p = 0; p < passes; ++p) {
    const int shift = 8 * p;
    PEB_LAUNCH(ctx, digit_histogram_kernel, n_tiles, kSortThreads, 0, ki, n, shift, hist, n_tiles);
    PEB_TRY(exclusive_scan_u32(ctx, hist, hist, kRadix * n_tiles, nullptr));
    PEB_LAUNCH(ctx, scatter_kernel, n_tiles, kSortThreads, 0, ki, vi, ko, vo, n, shift, hist, n_tiles);
    uint32_t* t = ki;
    ki = ko;
    ko = t;
    t = vi;
    vi = vo;
    vo = t;
  }
  *keys_out = ki;
  *vals_out = vi;
  return PEB_OK;
}

}  // namespace peb
