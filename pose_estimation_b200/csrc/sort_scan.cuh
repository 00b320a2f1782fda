// sort_scan.cuh — device-side pieces of the one-sweep radix sort and of the single-pass scan (radix_sort.cu) that the
// producers of the keys share: a kernel that computes a key (voxel id, grid cell id) adds it to the digit histograms of
// ALL passes in the same pass over the data, so the sort itself never re-reads the keys for counting.
#pragma once

#include "common.cuh"

namespace peb {

constexpr int kSortRadix = 256;     // 8-bit digits
constexpr int kSortMaxPasses = 4;   // 32-bit keys

// What a sort needs on the device; laid out in ctx->d_scratch by sort_prepare() and zeroed by ONE memset.
struct SortPlan {
  int passes = 0;
  int n_tiles = 0;
  int items = 0;            // keys per thread of a sort tile (16: large inputs, 4: small ones)
  uint32_t* hist = nullptr;        // [kSortMaxPasses][256] digit counts of the whole input
  uint32_t* tickets = nullptr;     // [kSortMaxPasses] tile tickets (a tile index is handed out in launch order)
  uint32_t* tile_state = nullptr;  // [passes][n_tiles][256] decoupled look-back words: flag << 30 | count
};

// Look-back word: the two top bits say what the 30-bit count is.
constexpr uint32_t kFlagAggregate = 1u << 30;  // this tile's own count
constexpr uint32_t kFlagInclusive = 2u << 30;  // the sum over this tile and all tiles before it
constexpr uint32_t kFlagMask = 3u << 30;
constexpr uint32_t kCountMask = ~kFlagMask;

__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Sum of the counts of all tiles before `tile` (decoupled look-back, one chain of words per caller): walks back over
// the predecessors' words, adding aggregates until it meets an inclusive sum.  Forward progress: tile indices are
// handed out by an atomic ticket, so every predecessor has started and publishes its aggregate without waiting.
__device__ __forceinline__ uint32_t lookback_exclusive(const uint32_t* state, int tile, int stride) {
  uint32_t excl = 0;
  for (int t = tile - 1; t >= 0; --t) {
    uint32_t s;
    while (((s = ld_relaxed_u32(state + static_cast<size_t>(t) * stride)) & kFlagMask) == 0u) __nanosleep(20);
    excl += s & kCountMask;
    if (s & kFlagInclusive) break;
  }
  return excl;
}

// The same for ONE chain walked by a whole warp (all 32 lanes call it): 32 predecessors per step — lane l looks at
// tile - 1 - l — instead of one dependent L2 round trip per predecessor.  The words of the step are usable up to the
// nearest inclusive sum; if a nearer word is not published yet the step is read again.
__device__ __forceinline__ uint32_t lookback_exclusive_warp(const uint32_t* state, int tile, int lane) {
  uint32_t excl = 0;
  for (int t = tile - 1;; t -= 32) {
    const int idx = t - lane;
    uint32_t s, upto, incl;
    for (;;) {
      s = idx >= 0 ? ld_relaxed_u32(state + idx) : kFlagInclusive;  // (before tile 0: an inclusive sum of zero)
      incl = __ballot_sync(0xFFFFFFFFu, (s & kFlagMask) == kFlagInclusive);
      const uint32_t none = __ballot_sync(0xFFFFFFFFu, (s & kFlagMask) == 0u);
      upto = incl ? ((2u << (__ffs(incl) - 1)) - 1u) : 0xFFFFFFFFu;  // the lanes up to the nearest inclusive word
      if ((none & upto) == 0u) break;
      __nanosleep(20);
    }
    uint32_t v = ((upto >> lane) & 1u) ? (s & kCountMask) : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    excl += v;
    if (incl) break;
  }
  return excl;
}

// ---- digit histograms of all passes, accumulated by the kernel that produces the keys ---------------------------
// sh: [passes][256] in shared memory, zeroed by the caller.  Warp-aggregated: one shared-memory atomic per distinct
// digit and warp (the high digits of cell ids are nearly constant: 32 lanes on one counter otherwise).
// All 32 lanes must call it (inactive lanes pass valid = false).
__device__ __forceinline__ void sort_hist_add(uint32_t (*sh)[kSortRadix], uint32_t key, bool valid, int passes) {
  const unsigned lane = threadIdx.x & 31u;
  for (int p = 0; p < passes; ++p) {
    const uint32_t digit = (key >> (8 * p)) & 0xFFu;
    const unsigned peers = __match_any_sync(0xFFFFFFFFu, valid ? digit : (0x100u + lane));
    if (valid && (peers & ((1u << lane) - 1u)) == 0u) atomicAdd(&sh[p][digit], static_cast<uint32_t>(__popc(peers)));
  }
}
// after a __syncthreads(): the block's counts go to the global histograms (blockDim.x >= 256 not required)
__device__ __forceinline__ void sort_hist_flush(uint32_t (*sh)[kSortRadix], uint32_t* hist, int passes) {
  for (int i = threadIdx.x; i < passes * kSortRadix; i += blockDim.x) {
    const uint32_t c = (&sh[0][0])[i];
    if (c) atomicAdd(hist + i, c);
  }
}

// radix_sort.cu
int sort_prepare(peb_ctx* ctx, int n, int key_bits, SortPlan* plan);
// the passes only: plan->hist already holds the digit counts of `keys` (sort_hist_add / sort_hist_flush)
int sort_pairs_counted(peb_ctx* ctx, const SortPlan& plan, uint32_t* keys, uint32_t* vals, uint32_t* keys_tmp,
                       uint32_t* vals_tmp, int n, uint32_t** keys_out, uint32_t** vals_out);

// ---- single-pass exclusive scan of one flag / count per thread ----------------------------------------------------
// For kernels that compute a per-element count and need its exclusive prefix in the same launch (stream compaction):
// block scan + decoupled look-back over the blocks.  state: n_blocks + 2 zeroed words (the last two: ticket, total).
struct ScanState {
  uint32_t* words = nullptr;  // [n_blocks] look-back words
  uint32_t* ticket = nullptr;
  uint32_t* total = nullptr;
};
// slot: which of the context's scan areas (a chain of kernels may keep several scans in flight)
int scan_state_prepare(peb_ctx* ctx, int n_blocks, ScanState* st, int slot);

constexpr int kScanBlock = 256;

// Tiles are processed in TICKET order (not blockIdx order): a kernel takes its ticket first and derives the elements
// it owns from it, so that a tile only ever waits for tiles that have already started.
__device__ __forceinline__ int scan_take_ticket(const ScanState& st) {
  __shared__ int s_tile;
  if (threadIdx.x == 0) s_tile = static_cast<int>(atomicAdd(st.ticket, 1u));
  __syncthreads();
  return s_tile;
}

// The two halves of a single-pass scan, for kernels whose tile is several block-sized rounds:
// block_exclusive_scan: prefix of v inside the block (shared memory only); *block_total = the block's sum.
// tile_lookback: the sum of the totals of every earlier tile (publishes this tile's total first; decoupled look-back by
// warp 0; the last tile leaves the grand total in *st.total).  blockDim.x == kScanBlock, every thread calls them.
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* block_total) {
  __shared__ uint32_t s_warp[kScanBlock / 32];
  __shared__ uint32_t s_total;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    const uint32_t w = lane < kScanBlock / 32 ? s_warp[lane] : 0u;
    uint32_t winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, winc, o);
      if (lane >= o) winc += t;
    }
    if (lane < kScanBlock / 32) s_warp[lane] = winc - w;
    if (lane == 31) s_total = winc;
  }
  __syncthreads();
  const uint32_t res = s_warp[warp] + inc - v;
  *block_total = s_total;
  __syncthreads();  // the shared words are reused by the next call
  return res;
}
__device__ __forceinline__ uint32_t tile_lookback(uint32_t total, const ScanState& st, int tile, int n_tiles) {
  __shared__ uint32_t s_base;
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    uint32_t* word = st.words + tile;
    uint32_t excl = 0;
    if (tile > 0) {
      if (lane == 0) st_relaxed_u32(word, kFlagAggregate | total);
      excl = lookback_exclusive_warp(st.words, tile, lane);
    }
    if (lane == 0) {
      st_relaxed_u32(word, kFlagInclusive | (excl + total));
      if (tile == n_tiles - 1) *st.total = excl + total;
      s_base = excl;
    }
  }
  __syncthreads();
  const uint32_t base = s_base;
  __syncthreads();
  return base;
}

// Exclusive prefix of v over every element before this thread's (all earlier tiles + the earlier threads of this
// tile); blockDim.x == kScanBlock, every thread of the block calls it.  The last tile leaves the grand total in
// *st.total.  block_total (nullable) receives this tile's own sum.
__device__ __forceinline__ uint32_t scan_exclusive(uint32_t v, const ScanState& st, int tile, int n_tiles,
                                                   uint32_t* block_total = nullptr) {
  uint32_t total;
  const uint32_t local = block_exclusive_scan(v, &total);
  const uint32_t base = tile_lookback(total, st, tile, n_tiles);
  if (block_total) *block_total = total;
  return base + local;
}

}  // namespace peb
