// prefilter.cu — the deterministic part of the reference's scene preparation, fused into one pass
// in front of VoxelGrid:
//   pcl::removeNaNFromPointCloud                  pose_estimation/src/pose_estimation.cpp:246-248
//   PoseEstimation::filter_points                 pose_estimation/src/pose_estimation.cpp:347-372
//   remove_planes' band test                      pose_estimation/src/pose_estimation.cpp:309-333
// One flag per point (the three tests are independent per point, so "one after the other" equals
// "all must pass"), exclusive scan, order-preserving scatter.  The float expressions keep the
// reference's operation order (-fmad=false), including its quirk of dividing the plane residual by
// the norm of the POINT instead of the norm of the plane normal.
// Algorithmic HBM bytes: 16 * N_in + 16 * N_out.
#include "core_math.cuh"

namespace peb {

namespace {

__device__ __forceinline__ bool prefilter_keep(const float4& p, const peb_prefilter_params& f) {
  if (!finite3(p.x, p.y, p.z)) return false;  // removeNaNFromPointCloud: std::isfinite on x, y, z
  if (f.use_sphere) {
    const float dx = f.sphere_center[0] - p.x;
    const float dy = f.sphere_center[1] - p.y;
    const float dz = f.sphere_center[2] - p.z;
    const float d = sqrtf(dx * dx + dy * dy + dz * dz);
    const bool inside = d <= f.sphere_radius;
    // ExtractIndices::setNegative(remove_inliers): negative -> everything BUT the listed (inside) points
    if (f.remove_inliers ? inside : !inside) return false;
  }
  for (int k = 0; k < f.n_planes; ++k) {
    const float* c = f.planes + 4 * k;
    const float d = (p.x * c[0] + p.y * c[1] + p.z * c[2] + c[3]) / sqrtf(p.x * p.x + p.y * p.y + p.z * p.z);
    if (fabsf(d) <= f.plane_band) return false;  // a NaN residual (point at the origin) compares false: kept
  }
  return true;
}

__global__ void __launch_bounds__(256) prefilter_flag_kernel(const float4* __restrict__ in, int n,
                                                             const peb_prefilter_params f, uint32_t* __restrict__ flags) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  flags[i] = prefilter_keep(in[i], f) ? 1u : 0u;
}

__global__ void __launch_bounds__(256) prefilter_scatter_kernel(const float4* __restrict__ in, int n,
                                                                const uint32_t* __restrict__ flags,
                                                                const uint32_t* __restrict__ slot, float4* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !flags[i]) return;
  float4 p = in[i];
  p.w = 1.0f;
  out[slot[i]] = p;
}

}  // namespace

int scene_prefilter_device(peb_ctx* ctx, const float4* d_in, int n, const peb_prefilter_params* prm, float4* d_out,
                           size_t* out_n) {
  *out_n = 0;
  if (prm->n_planes < 0 || prm->n_planes > PEB_PREFILTER_MAX_PLANES)
    return fail(ctx, PEB_E_INVALID_ARG, "scene_prefilter: n_planes %d out of [0, %d]", prm->n_planes, PEB_PREFILTER_MAX_PLANES);
  if (n == 0) return PEB_OK;
  PEB_CUDA(ctx, ctx->vg_flags.ensure(static_cast<size_t>(n) * 4));
  PEB_CUDA(ctx, ctx->vg_scan.ensure(static_cast<size_t>(n) * 4));
  uint32_t* flags = ctx->vg_flags.as<uint32_t>();
  uint32_t* slot = ctx->vg_scan.as<uint32_t>();
  uint32_t* d_total = ctx->d_small.as<uint32_t>() + 40;
  uint32_t* h_total = ctx->h_small.as<uint32_t>() + 40;
  PEB_LAUNCH(ctx, prefilter_flag_kernel, ceil_div(n, 256), 256, 0, d_in, n, *prm, flags);
  PEB_TRY(exclusive_scan_u32(ctx, flags, slot, n, d_total));
  PEB_LAUNCH(ctx, prefilter_scatter_kernel, ceil_div(n, 256), 256, 0, d_in, n, flags, slot, d_out);
  PEB_CUDA(ctx, cudaMemcpyAsync(h_total, d_total, 4, cudaMemcpyDeviceToHost, ctx->stream));
  PEB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  *out_n = static_cast<size_t>(*h_total);
  return PEB_OK;
}

}  // namespace peb
