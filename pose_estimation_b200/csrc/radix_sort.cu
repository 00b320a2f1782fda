// radix_sort.cu — stable LSD radix sort of (uint32 key, uint32 value) pairs ("one sweep": ONE kernel per 8-bit digit)
// and a single-pass device-wide exclusive scan.  They feed VoxelGrid (sort by voxel id, [PCL]
// filters/impl/voxel_grid.hpp "second pass") and the uniform-grid build that replaces the kd-tree (SURVEY.md 8a-2').
//
//   digit counts   the histograms of ALL passes are taken in one pass over the keys — by the kernel that produces the
//                  keys (sort_scan.cuh : sort_hist_add) or, for keys that already exist, by sort_hist_kernel
//   one pass       onesweep_kernel: a CTA takes a tile (ticket order), ranks its keys stably (warp match_any + per-warp
//                  digit counters), publishes the tile's digit counts, obtains the counts of all earlier tiles by
//                  decoupled look-back (256 chains, one per digit and thread), reorders the tile through shared memory
//                  and writes runs of equal digits to consecutive addresses — coalesced stores, one read of the input
// Only ceil(key_bits / 8) passes run.  Stability keeps equal keys in ascending original index, which fixes the float
// summation order of the voxel centroids.
// Round 1 ran five launches per pass (histogram, three-kernel scan, scatter from registers): 20 launches for the 25-bit
// voxel ids of the 2.33 M-point scene; this version: memset + 4.
#include "sort_scan.cuh"

namespace peb {

namespace {

constexpr int kSortThreads = 256;
constexpr int kWarps = kSortThreads / 32;

// digit histograms of all passes for keys that already exist (the debug entry point; the library's own producers count
// while they write the keys)
__global__ void __launch_bounds__(kSortThreads) sort_hist_kernel(const uint32_t* __restrict__ keys, int n, int passes,
                                                                 uint32_t* __restrict__ hist) {
  __shared__ uint32_t sh[kSortMaxPasses][kSortRadix];
  for (int i = threadIdx.x; i < kSortMaxPasses * kSortRadix; i += kSortThreads) (&sh[0][0])[i] = 0;
  __syncthreads();
  const int stride = gridDim.x * kSortThreads;
  const int rounds = (n + stride - 1) / stride;  // every lane runs every round: the warp votes need all 32
  for (int r = 0; r < rounds; ++r) {
    const int i = r * stride + blockIdx.x * kSortThreads + threadIdx.x;
    const bool ok = i < n;
    sort_hist_add(sh, ok ? keys[i] : 0u, ok, passes);
  }
  __syncthreads();
  sort_hist_flush(sh, hist, passes);
}

template <int ITEMS>
__global__ void __launch_bounds__(kSortThreads) onesweep_kernel(const uint32_t* __restrict__ keys_in,
                                                                const uint32_t* __restrict__ vals_in,
                                                                uint32_t* __restrict__ keys_out,
                                                                uint32_t* __restrict__ vals_out, int n, int shift,
                                                                const uint32_t* __restrict__ hist,  // this pass: [256]
                                                                uint32_t* __restrict__ tile_state,  // this pass: [n_tiles][256]
                                                                uint32_t* __restrict__ ticket) {
  constexpr int kTile = kSortThreads * ITEMS;
  __shared__ uint32_t warp_cnt[kWarps][kSortRadix];  // per-warp running digit counters -> exclusive over the warps
  __shared__ uint32_t local_start[kSortRadix];       // first position of a digit inside the reordered tile
  __shared__ uint32_t dst_base[kSortRadix];          // global address of position 0 of a digit's run, minus local_start
  __shared__ uint32_t sk[kTile], sv[kTile];          // the reordered tile
  __shared__ uint32_t s_scan[kWarps];
  __shared__ int s_tile;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) s_tile = static_cast<int>(atomicAdd(ticket, 1u));
  for (int i = threadIdx.x; i < kWarps * kSortRadix; i += kSortThreads) (&warp_cnt[0][0])[i] = 0;
  __syncthreads();
  const int tile = s_tile;
  const int tile_n = min(kTile, n - tile * kTile);

  // warp w owns the contiguous chunk [w * 32 * ITEMS, + 32 * ITEMS) of the tile: item r of lane l is element r * 32 + l,
  // so consecutive rounds and lanes walk the chunk in index order (stability)
  const int chunk = tile * kTile + warp * (32 * ITEMS);
  uint32_t key[ITEMS], val[ITEMS], rank[ITEMS];
#pragma unroll
  for (int r = 0; r < ITEMS; ++r) {
    const int i = chunk + r * 32 + lane;
    const bool ok = i < n;
    key[r] = ok ? keys_in[i] : 0xFFFFFFFFu;
    val[r] = ok ? vals_in[i] : 0u;
  }
#pragma unroll
  for (int r = 0; r < ITEMS; ++r) {
    const bool ok = chunk + r * 32 + lane < n;
    const uint32_t digit = (key[r] >> shift) & 0xFFu;
    // lanes past the end take part in the vote with a digit nobody else can rank against
    const uint32_t peers = __match_any_sync(0xFFFFFFFFu, ok ? digit : (0x100u + lane));
    const uint32_t lower = peers & ((1u << lane) - 1u);
    uint32_t before = 0;
    if (ok) before = warp_cnt[warp][digit];
    __syncwarp();
    if (ok && lower == 0) warp_cnt[warp][digit] = before + __popc(peers);  // the first peer bumps the counter
    __syncwarp();
    rank[r] = before + __popc(lower);
  }
  __syncthreads();

  // thread d: digit d.  Counters -> exclusive over the warps; the tile's count of d goes to the look-back chain at once
  const int d = threadIdx.x;
  uint32_t count = 0;
#pragma unroll
  for (int w = 0; w < kWarps; ++w) {
    const uint32_t c = warp_cnt[w][d];
    warp_cnt[w][d] = count;
    count += c;
  }
  uint32_t* word = tile_state + static_cast<size_t>(tile) * kSortRadix + d;
  st_relaxed_u32(word, (tile == 0 ? kFlagInclusive : kFlagAggregate) | count);

  // exclusive scans over the digits: of the tile's counts (positions inside the tile) and of the global histogram
  // (where a digit's run starts in the output)
  uint32_t inc_local = count, inc_global = hist[d];
  const uint32_t hist_d = inc_global;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t a = __shfl_up_sync(0xFFFFFFFFu, inc_local, o);
    const uint32_t b = __shfl_up_sync(0xFFFFFFFFu, inc_global, o);
    if (lane >= o) {
      inc_local += a;
      inc_global += b;
    }
  }
  __shared__ uint32_t s_scan_g[kWarps];
  if (lane == 31) {
    s_scan[warp] = inc_local;
    s_scan_g[warp] = inc_global;
  }
  __syncthreads();
  uint32_t off_local = 0, off_global = 0;
#pragma unroll
  for (int w = 0; w < kWarps; ++w) {
    if (w < warp) {
      off_local += s_scan[w];
      off_global += s_scan_g[w];
    }
  }
  const uint32_t start_local = off_local + inc_local - count;
  const uint32_t start_global = off_global + inc_global - hist_d;

  uint32_t before_tiles = 0;
  if (tile > 0) {
    before_tiles = lookback_exclusive(tile_state + d, tile, kSortRadix);
    st_relaxed_u32(word, kFlagInclusive | (before_tiles + count));
  }
  local_start[d] = start_local;
  dst_base[d] = start_global + before_tiles - start_local;
  __syncthreads();

  // reorder through shared memory: position = start of the digit + keys of the digit in earlier warps + rank in the warp
#pragma unroll
  for (int r = 0; r < ITEMS; ++r) {
    if (chunk + r * 32 + lane < n) {
      const uint32_t digit = (key[r] >> shift) & 0xFFu;
      const uint32_t pos = local_start[digit] + warp_cnt[warp][digit] + rank[r];
      sk[pos] = key[r];
      sv[pos] = val[r];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < tile_n; i += kSortThreads) {
    const uint32_t k = sk[i];
    const uint32_t dst = dst_base[(k >> shift) & 0xFFu] + static_cast<uint32_t>(i);
    keys_out[dst] = k;
    vals_out[dst] = sv[i];
  }
}

// ---- exclusive scan of an array: one launch (block scan + decoupled look-back) ------------------------------------
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanBlock * kScanItems;

__global__ void __launch_bounds__(kScanBlock) scan_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int n,
                                                          ScanState st, int n_tiles, uint32_t* __restrict__ d_total) {
  const int tile = scan_take_ticket(st);
  const int base = tile * kScanTile + threadIdx.x * kScanItems;
  uint32_t v[kScanItems];
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    v[i] = (base + i < n) ? in[base + i] : 0u;
    s += v[i];
  }
  uint32_t ex = scan_exclusive(s, st, tile, n_tiles);
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    if (base + i < n) out[base + i] = ex;
    ex += v[i];
  }
  if (d_total && tile == n_tiles - 1 && threadIdx.x == kScanBlock - 1) *d_total = ex;
}

}  // namespace

// ctx->scan_scratch holds kScanSlots areas of look-back words; each prepare zeroes its area with one memset
int scan_state_prepare(peb_ctx* ctx, int n_blocks, ScanState* st, int slot) {
  constexpr int kScanSlots = 4;
  if (slot < 0 || slot >= kScanSlots) return fail(ctx, PEB_E_INVALID_ARG, "scan_state_prepare: slot %d", slot);
  const size_t words = static_cast<size_t>(n_blocks) + 2;
  DevBuf& buf = ctx->scan_scratch[slot];
  PEB_CUDA(ctx, buf.ensure(words * sizeof(uint32_t)));
  PEB_CUDA(ctx, cudaMemsetAsync(buf.p, 0, words * sizeof(uint32_t), ctx->stream));
  st->words = buf.as<uint32_t>();
  st->ticket = st->words + n_blocks;
  st->total = st->words + n_blocks + 1;
  return PEB_OK;
}

// out may alias in.  d_total (nullable) receives the sum of all elements.
int exclusive_scan_u32(peb_ctx* ctx, const uint32_t* in, uint32_t* out, int n, uint32_t* d_total) {
  if (n <= 0) {
    if (d_total) PEB_CUDA(ctx, cudaMemsetAsync(d_total, 0, sizeof(uint32_t), ctx->stream));
    return PEB_OK;
  }
  const int n_tiles = ceil_div(n, kScanTile);
  ScanState st;
  PEB_TRY(scan_state_prepare(ctx, n_tiles, &st, 3));
  PEB_LAUNCH(ctx, scan_kernel, n_tiles, kScanBlock, 0, in, out, n, st, n_tiles, d_total);
  return PEB_OK;
}

int sort_prepare(peb_ctx* ctx, int n, int key_bits, SortPlan* plan) {
  plan->passes = std::min(kSortMaxPasses, std::max(1, (key_bits + 7) / 8));
  plan->items = n >= (1 << 20) ? 16 : 4;  // small inputs: more, smaller tiles so that the launch still fills the SMs
  plan->n_tiles = std::max(1, ceil_div(n, kSortThreads * plan->items));
  const size_t head = static_cast<size_t>(kSortMaxPasses) * kSortRadix + kSortMaxPasses;  // histograms + tickets
  const size_t words = head + static_cast<size_t>(plan->passes) * plan->n_tiles * kSortRadix;
  PEB_CUDA(ctx, ctx->sort_scratch.ensure(words * sizeof(uint32_t)));
  PEB_CUDA(ctx, cudaMemsetAsync(ctx->sort_scratch.p, 0, words * sizeof(uint32_t), ctx->stream));
  plan->hist = ctx->sort_scratch.as<uint32_t>();
  plan->tickets = plan->hist + static_cast<size_t>(kSortMaxPasses) * kSortRadix;
  plan->tile_state = plan->hist + head;
  return PEB_OK;
}

int sort_pairs_counted(peb_ctx* ctx, const SortPlan& plan, uint32_t* keys, uint32_t* vals, uint32_t* keys_tmp,
                       uint32_t* vals_tmp, int n, uint32_t** keys_out, uint32_t** vals_out) {
  uint32_t *ki = keys, *vi = vals, *ko = keys_tmp, *vo = vals_tmp;
  for (int p = 0; p < plan.passes && n > 1; ++p) {
    const uint32_t* hist = plan.hist + static_cast<size_t>(p) * kSortRadix;
    uint32_t* state = plan.tile_state + static_cast<size_t>(p) * plan.n_tiles * kSortRadix;
    if (plan.items == 16)
      PEB_LAUNCH(ctx, onesweep_kernel<16>, plan.n_tiles, kSortThreads, 0, ki, vi, ko, vo, n, 8 * p, hist, state, plan.tickets + p);
    else
      PEB_LAUNCH(ctx, onesweep_kernel<4>, plan.n_tiles, kSortThreads, 0, ki, vi, ko, vo, n, 8 * p, hist, state, plan.tickets + p);
    std::swap(ki, ko);
    std::swap(vi, vo);
  }
  *keys_out = ki;
  *vals_out = vi;
  return PEB_OK;
}

// Sorts n pairs by the low key_bits bits of the key; ping-pongs between the two buffer pairs and
// reports where the result landed.
int sort_pairs(peb_ctx* ctx, uint32_t* keys, uint32_t* vals, uint32_t* keys_tmp, uint32_t* vals_tmp, int n,
               int key_bits, uint32_t** keys_out, uint32_t** vals_out) {
  *keys_out = keys;
  *vals_out = vals;
  if (n <= 1) return PEB_OK;
  SortPlan plan;
  PEB_TRY(sort_prepare(ctx, n, key_bits, &plan));
  const int blocks = std::min(ceil_div(n, kSortThreads * 8), kSmCount * 4);
  PEB_LAUNCH(ctx, sort_hist_kernel, blocks, kSortThreads, 0, keys, n, plan.passes, plan.hist);
  return sort_pairs_counted(ctx, plan, keys, vals, keys_tmp, vals_tmp, n, keys_out, vals_out);
}

}  // namespace peb
