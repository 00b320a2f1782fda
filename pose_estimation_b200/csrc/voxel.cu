// voxel.cu — pcl::VoxelGrid<PointXYZ>::applyFilter on the device
// ([PCL] filters/impl/voxel_grid.hpp; getMinMax3D [PCL] common/impl/common.hpp; centroid
// [PCL] common/impl/centroid.hpp : CentroidPoint).
//
//   bbox            finite min/max                                            (grid.cu)
//   host            inv_leaf, the int64 overflow guard, min_b / div_b / divb_mul — PCL's scalars
//   voxel_key       idx = ijk . divb_mul per finite point (float floor arithmetic of PCL);
//                   non-finite points get a key above every voxel id
//   radix sort      (idx, point index), stable => within a voxel the points stay in ascending
//                   original index, the summation order the oracle defines
//   run heads       flag + exclusive scan -> voxel ordinal -> run start table
//   (min_pts >= 2)  keep flag + scan -> output slot
//   centroid        one thread per voxel adds its run SEQUENTIALLY in float (bit-exact with
//                   CentroidPoint's Vector3f += ... / n), writes (cx, cy, cz, 1)
// Output order = ascending voxel index (x fastest) like PCL.  Algorithmic HBM bytes:
// 16 * N_in + 16 * M_out; the sort passes are overhead on top (L2 resident at 2.3 M points).
#include "core_math.cuh"

namespace peb {

namespace {

struct VoxelParams {
  float inv[3];
  float min_b[3];   // (float)min_b, as PCL subtracts it
  uint32_t mul[3];  // divb_mul, applied in wrapping 32-bit arithmetic like PCL's int
  uint32_t sentinel;
};

__global__ void __launch_bounds__(256) voxel_key_kernel(const float4* __restrict__ pts, int n, VoxelParams vp,
                                                        uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = pts[i];
  uint32_t key = vp.sentinel;
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

__global__ void __launch_bounds__(256) run_head_kernel(const uint32_t* __restrict__ sorted_keys, int n_finite,
                                                       uint32_t* __restrict__ flags) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_finite) return;
  flags[j] = (j == 0 || sorted_keys[j] != sorted_keys[j - 1]) ? 1u : 0u;
}

// starts[ordinal] = j for every run head; the last element also writes the end sentinel starts[n_runs]
__global__ void __launch_bounds__(256) run_start_kernel(const uint32_t* __restrict__ flags,
                                                        const uint32_t* __restrict__ ordinal, int n_finite,
                                                        uint32_t* __restrict__ starts) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_finite) return;
  if (flags[j]) starts[ordinal[j]] = static_cast<uint32_t>(j);
  if (j == n_finite - 1) starts[ordinal[j] + flags[j]] = static_cast<uint32_t>(n_finite);  // = starts[n_runs]
}

__global__ void __launch_bounds__(256) run_keep_kernel(const uint32_t* __restrict__ starts, int n_runs, unsigned min_pts,
                                                       uint32_t* __restrict__ keep) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n_runs) return;
  keep[v] = (starts[v + 1] - starts[v] >= min_pts) ? 1u : 0u;
}

// slot == nullptr: every run is kept and slot = v
__global__ void __launch_bounds__(128) centroid_kernel(const float4* __restrict__ pts, const uint32_t* __restrict__ sorted_vals,
                                                       const uint32_t* __restrict__ starts, int n_runs,
                                                       const uint32_t* __restrict__ keep, const uint32_t* __restrict__ slot,
                                                       float4* __restrict__ out) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n_runs) return;
  if (keep && !keep[v]) return;
  const uint32_t s = starts[v], e = starts[v + 1];
  float sx = 0.0f, sy = 0.0f, sz = 0.0f;
  for (uint32_t j = s; j < e; ++j) {
    const float4 p = pts[sorted_vals[j]];
    sx += p.x;
    sy += p.y;
    sz += p.z;
  }
  const float cnt = static_cast<float>(e - s);
  out[slot ? slot[v] : static_cast<uint32_t>(v)] = make_float4(sx / cnt, sy / cnt, sz / cnt, 1.0f);
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
  PEB_LAUNCH(ctx, voxel_key_kernel, ceil_div(n, 256), 256, 0, d_in, n, vp, g.keys.as<uint32_t>(), g.vals.as<uint32_t>());
  uint32_t *sk = nullptr, *sv = nullptr;
  PEB_TRY(sort_pairs(ctx, g.keys.as<uint32_t>(), g.vals.as<uint32_t>(), g.keys_tmp.as<uint32_t>(),
                     g.vals_tmp.as<uint32_t>(), n, key_bits, &sk, &sv));

  PEB_CUDA(ctx, ctx->vg_flags.ensure(static_cast<size_t>(n_finite) * 4));
  PEB_CUDA(ctx, ctx->vg_scan.ensure(static_cast<size_t>(n_finite) * 4));
  PEB_CUDA(ctx, ctx->vg_starts.ensure((static_cast<size_t>(n_finite) + 1) * 4));
  uint32_t* flags = ctx->vg_flags.as<uint32_t>();
  uint32_t* ordinal = ctx->vg_scan.as<uint32_t>();
  uint32_t* starts = ctx->vg_starts.as<uint32_t>();
  uint32_t* d_total = ctx->d_small.as<uint32_t>() + 32;
  uint32_t* h_total = ctx->h_small.as<uint32_t>() + 32;
  PEB_LAUNCH(ctx, run_head_kernel, ceil_div(n_finite, 256), 256, 0, sk, n_finite, flags);
  PEB_TRY(exclusive_scan_u32(ctx, flags, ordinal, n_finite, d_total));
  PEB_LAUNCH(ctx, run_start_kernel, ceil_div(n_finite, 256), 256, 0, flags, ordinal, n_finite, starts);
  PEB_CUDA(ctx, cudaMemcpyAsync(h_total, d_total, 4, cudaMemcpyDeviceToHost, ctx->stream));
  PEB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  const int n_runs = static_cast<int>(*h_total);
  size_t n_out = static_cast<size_t>(n_runs);
  const uint32_t* keep = nullptr;
  const uint32_t* slot = nullptr;
  if (min_pts >= 2) {
    // flags / ordinal are free again: reuse them as keep / slot
    PEB_LAUNCH(ctx, run_keep_kernel, ceil_div(n_runs, 256), 256, 0, starts, n_runs, min_pts, flags);
    PEB_TRY(exclusive_scan_u32(ctx, flags, ordinal, n_runs, d_total));
    PEB_CUDA(ctx, cudaMemcpyAsync(h_total, d_total, 4, cudaMemcpyDeviceToHost, ctx->stream));
    PEB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    n_out = static_cast<size_t>(*h_total);
    keep = flags;
    slot = ordinal;
  }
  PEB_LAUNCH(ctx, centroid_kernel, ceil_div(n_runs, 128), 128, 0, d_in, sv, starts, n_runs, keep, slot, d_out);
  *out_n = n_out;
  return PEB_OK;
}

}  // namespace peb
