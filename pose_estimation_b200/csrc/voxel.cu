// voxel.cu — pcl::VoxelGrid<PointXYZ>::applyFilter on the device
// ([PCL] filters/impl/voxel_grid.hpp; getMinMax3D [PCL] common/impl/common.hpp; centroid
// [PCL] common/impl/centroid.hpp : CentroidPoint).
//
//   bbox            finite min/max                                            (grid.cu)
//   host            inv_leaf, the int64 overflow guard, min_b / div_b / divb_mul — PCL's scalars
//   voxel_key       idx = ijk . divb_mul per finite point (float floor arithmetic of PCL); non-finite points get a key
//                   above every voxel id; the same kernel counts the digits of every sort pass (sort_scan.cuh)
//   radix sort      (idx, point index), stable, one kernel per 8-bit digit => within a voxel the points stay in
//                   ascending original index, the summation order the oracle defines
//   centroid        ONE kernel: the thread at the head of a run of equal keys adds the run SEQUENTIALLY in float
//                   (bit-exact with CentroidPoint's Vector3f += ... / n), runs with >= min_pts points get their output
//                   slot from a single-pass scan (block scan + decoupled look-back) in the same launch
// Output order = ascending voxel index (x fastest) like PCL.  Algorithmic HBM bytes:
// 16 * N_in + 16 * M_out; the sort passes are overhead on top (L2 resident at 2.3 M points).
// Launches for the 2.33 M-point scene: 1 + 2 memsets + 1 + 4 + 1 = 9 (round 1: 31) and two host round trips (bounding
// box for PCL's scalars, the output count).
#include <algorithm>

#include "core_math.cuh"
#include "sort_scan.cuh"

namespace peb {

namespace {

struct VoxelParams {
  float inv[3];
  float min_b[3];   // (float)min_b, as PCL subtracts it
  uint32_t mul[3];  // divb_mul, applied in wrapping 32-bit arithmetic like PCL's int
  uint32_t sentinel;
};

__global__ void __launch_bounds__(256) voxel_key_kernel(const float4* __restrict__ pts, int n, VoxelParams vp,
                                                        uint32_t* __restrict__ keys, uint32_t* __restrict__ vals,
                                                        int passes, uint32_t* __restrict__ hist) {
  __shared__ uint32_t sh[kSortMaxPasses][kSortRadix];
  for (int i = threadIdx.x; i < kSortMaxPasses * kSortRadix; i += 256) (&sh[0][0])[i] = 0;
  __syncthreads();
  const int stride = gridDim.x * 256;
  const int rounds = (n + stride - 1) / stride;  // every lane runs every round (warp votes in sort_hist_add)
  for (int r = 0; r < rounds; ++r) {
    const int i = r * stride + blockIdx.x * 256 + threadIdx.x;
    const bool in = i < n;
    uint32_t key = vp.sentinel;
    if (in) {
      const float4 p = pts[i];
      if (finite3(p.x, p.y, p.z)) {
        const int ijk0 = static_cast<int>(floorf(p.x * vp.inv[0]) - vp.min_b[0]);
        const int ijk1 = static_cast<int>(floorf(p.y * vp.inv[1]) - vp.min_b[1]);
        const int ijk2 = static_cast<int>(floorf(p.z * vp.inv[2]) - vp.min_b[2]);
        key = static_cast<uint32_t>(ijk0) * vp.mul[0] + static_cast<uint32_t>(ijk1) * vp.mul[1] +
              static_cast<uint32_t>(ijk2) * vp.mul[2];
      }
      keys[i] = key;
      vals[i] = static_cast<uint32_t>(i);
    }
    sort_hist_add(sh, key, in, passes);
  }
  __syncthreads();
  sort_hist_flush(sh, hist, passes);
}

// A tile = kCentroidRounds x 256 consecutive sorted positions (tiles in ticket order), one thread per position and
// round.  Every thread fetches ITS point (a gather with all lanes busy) into shared memory; the thread at the head of a
// run of equal voxel ids then walks the run there: count, and the SEQUENTIAL float sum of its points in sorted order =
// ascending original index (a run that leaves the round continues from global memory).  A run with at least min_pts
// points is an output point; its slot is the number of such runs before it: block scans inside the tile, ONE
// decoupled look-back per tile (2.33 M points: 1 140 tiles, one wave of blocks — with one round per tile the 8 750
// tiles queued behind each other's look-back).
constexpr int kCentroidRounds = 8;
__global__ void __launch_bounds__(kScanBlock) voxel_centroid_kernel(const float4* __restrict__ pts,
                                                                    const uint32_t* __restrict__ sorted_keys,
                                                                    const uint32_t* __restrict__ sorted_vals, int n_finite,
                                                                    unsigned min_pts, float4* __restrict__ out,
                                                                    ScanState st, int n_tiles) {
  __shared__ float s_x[kScanBlock], s_y[kScanBlock], s_z[kScanBlock];
  __shared__ uint32_t s_key[kScanBlock + 1];
  const int tile = scan_take_ticket(st);
  const int t = threadIdx.x;
  float cx[kCentroidRounds], cy[kCentroidRounds], cz[kCentroidRounds];
  uint32_t slot[kCentroidRounds];  // offset inside the tile, 0xFFFFFFFF = not an output point
  uint32_t running = 0;
#pragma unroll
  for (int r = 0; r < kCentroidRounds; ++r) {
    const int j0 = (tile * kCentroidRounds + r) * kScanBlock;
    const int j = j0 + t;
    uint32_t key = 0xFFFFFFFFu;
    if (j < n_finite) {
      key = sorted_keys[j];
      const float4 p = pts[sorted_vals[j]];
      s_x[t] = p.x;
      s_y[t] = p.y;
      s_z[t] = p.z;
    }
    s_key[t] = key;  // (finite points have keys below the sentinel, so 0xFFFFFFFF ends every run)
    if (t == 0) s_key[kScanBlock] = 0xFFFFFFFFu;
    uint32_t prev_key = 0xFFFFFFFFu;
    if (t == 0 && j > 0 && j < n_finite) prev_key = sorted_keys[j - 1];
    __syncthreads();
    if (t > 0) prev_key = s_key[t - 1];
    float sx = 0.0f, sy = 0.0f, sz = 0.0f;
    uint32_t len = 0;
    if (j < n_finite && (j == 0 || prev_key != key)) {
      int e = t;
      do {
        sx += s_x[e];
        sy += s_y[e];
        sz += s_z[e];
        ++e;
      } while (s_key[e] == key);
      if (e == kScanBlock) {  // the run goes on beyond this round
        int g = j0 + e;
        while (g < n_finite && sorted_keys[g] == key) {
          const float4 p = pts[sorted_vals[g]];
          sx += p.x;
          sy += p.y;
          sz += p.z;
          ++g;
        }
        len = static_cast<uint32_t>(g - j);
      } else {
        len = static_cast<uint32_t>(e - t);
      }
    }
    const uint32_t keep = (len > 0 && len >= min_pts) ? 1u : 0u;
    uint32_t round_total;
    const uint32_t off = block_exclusive_scan(keep, &round_total);  // (its barriers also release the staging arrays)
    slot[r] = keep ? running + off : 0xFFFFFFFFu;
    running += round_total;
    const float cnt = static_cast<float>(len);
    cx[r] = sx / cnt;
    cy[r] = sy / cnt;
    cz[r] = sz / cnt;
  }
  const uint32_t base = tile_lookback(running, st, tile, n_tiles);
#pragma unroll
  for (int r = 0; r < kCentroidRounds; ++r)
    if (slot[r] != 0xFFFFFFFFu) out[base + slot[r]] = make_float4(cx[r], cy[r], cz[r], 1.0f);
}

__global__ void __launch_bounds__(256) copy_xyz1_kernel(const float4* __restrict__ in, int n, float4* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 p = in[i];
  p.w = 1.0f;
  out[i] = p;
}

}  // namespace

int voxel_grid_device(peb_ctx* ctx, const float4* d_in, int n, float lx, float ly, float lz, unsigned min_pts,
                      float4* d_out, size_t* out_n) {
  *out_n = 0;
  if (!(lx > 0.0f) || !(ly > 0.0f) || !(lz > 0.0f))
    return fail(ctx, PEB_E_INVALID_ARG, "voxel_grid: leaf size must be positive (got %g %g %g)", lx, ly, lz);
  if (n == 0) return PEB_OK;
  float mn[3], mx[3];
  int n_finite = 0;
  PEB_TRY(bbox_finite(ctx, d_in, n, mn, mx, &n_finite));
  const float leaf[3] = {lx, ly, lz};
  const float inv[3] = {1.0f / leaf[0], 1.0f / leaf[1], 1.0f / leaf[2]};
  if (n_finite == 0) return PEB_OK;  // nothing finite: PCL's output is empty
  // [PCL] voxel_grid.hpp: "Leaf size is too small for the input dataset. Integer indices would overflow."
  const int64_t gx = static_cast<int64_t>((mx[0] - mn[0]) * inv[0]) + 1;
  const int64_t gy = static_cast<int64_t>((mx[1] - mn[1]) * inv[1]) + 1;
  const int64_t gz = static_cast<int64_t>((mx[2] - mn[2]) * inv[2]) + 1;
  if (gx * gy * gz > static_cast<int64_t>(INT32_MAX)) {
    PEB_LAUNCH(ctx, copy_xyz1_kernel, ceil_div(n, 256), 256, 0, d_in, n, d_out);
    *out_n = static_cast<size_t>(n);
    return PEB_OK;
  }
  VoxelParams vp;
  int64_t div_b[3];
  for (int d = 0; d < 3; ++d) {
    const int min_b = static_cast<int>(floorf(mn[d] * inv[d]));
    const int max_b = static_cast<int>(floorf(mx[d] * inv[d]));
    div_b[d] = static_cast<int64_t>(max_b) - min_b + 1;
    vp.inv[d] = inv[d];
    vp.min_b[d] = static_cast<float>(min_b);
  }
  vp.mul[0] = 1u;
  vp.mul[1] = static_cast<uint32_t>(div_b[0]);
  vp.mul[2] = static_cast<uint32_t>(div_b[0] * div_b[1]);
  const int64_t cells = div_b[0] * div_b[1] * div_b[2];
  if (cells >= (1ll << 32) - 1) return fail(ctx, PEB_E_UNSUPPORTED, "voxel_grid: %lld voxels exceed 32-bit ids", (long long)cells);
  vp.sentinel = static_cast<uint32_t>(cells);
  int key_bits = 1;
  while ((1ll << key_bits) <= cells) ++key_bits;

  Grid& g = ctx->aux_grid;  // reuse its sort buffers
  PEB_CUDA(ctx, g.keys.ensure(static_cast<size_t>(n) * 4));
  PEB_CUDA(ctx, g.vals.ensure(static_cast<size_t>(n) * 4));
  PEB_CUDA(ctx, g.keys_tmp.ensure(static_cast<size_t>(n) * 4));
  PEB_CUDA(ctx, g.vals_tmp.ensure(static_cast<size_t>(n) * 4));
  SortPlan plan;
  PEB_TRY(sort_prepare(ctx, n, key_bits, &plan));
  const int key_blocks = std::min(ceil_div(n, 256 * 8), kSmCount * 8);
  PEB_LAUNCH(ctx, voxel_key_kernel, key_blocks, 256, 0, d_in, n, vp, g.keys.as<uint32_t>(), g.vals.as<uint32_t>(),
             plan.passes, plan.hist);
  uint32_t *sk = nullptr, *sv = nullptr;
  PEB_TRY(sort_pairs_counted(ctx, plan, g.keys.as<uint32_t>(), g.vals.as<uint32_t>(), g.keys_tmp.as<uint32_t>(),
                             g.vals_tmp.as<uint32_t>(), n, &sk, &sv));
  const int n_tiles = ceil_div(n_finite, kScanBlock * kCentroidRounds);
  ScanState st;
  PEB_TRY(scan_state_prepare(ctx, n_tiles, &st, 0));
  PEB_LAUNCH(ctx, voxel_centroid_kernel, n_tiles, kScanBlock, 0, d_in, sk, sv, n_finite, min_pts, d_out, st, n_tiles);
  uint32_t* h_total = ctx->h_small.as<uint32_t>() + 32;
  PEB_CUDA(ctx, cudaMemcpyAsync(h_total, st.total, 4, cudaMemcpyDeviceToHost, ctx->stream));
  PEB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  *out_n = static_cast<size_t>(*h_total);
  return PEB_OK;
}

}  // namespace peb
