// sac.cu — pcl::SACSegmentation<PointXYZ>::segment for SACMODEL_PLANE + SAC_RANSAC (SURVEY.md 8f rank 1:
// the plane fit of the reference's remove_planes, pose_estimation/src/pose_estimation.cpp:285-297).
//
// PCL's loop is sequential (draw a sample, fit a plane, count its inliers over the whole cloud, keep
// the best, shrink the iteration bound k), but the sample sequence comes from a seeded Mersenne
// twister and a persistent partial shuffle: it does not depend on the data.  So
//   host   : draws the index triples of a whole run up front (sample_draw),
//   device : gathers their coordinates (a few hundred points back to the host),
//   host   : applies isSampleGood / computeModelCoefficients in PCL's float arithmetic,
//   device : counts the inliers of ALL candidate planes in ONE pass over the cloud
//            (planes in shared memory, 16 B per point of HBM traffic instead of 16 B x models),
//   host   : replays RandomSampleConsensus::computeModel on the counts — same decisions,
//   device : selects the inliers, accumulates their moments (double), the host solves the 3x3
//            eigen problem (eigen33, the same code as the normals), the device re-selects.
// If the replay needs more candidates than were drawn (many degenerate samples) another batch follows.
#include <algorithm>
#include <cmath>
#include <limits>
#include <unordered_map>
#include <vector>

#include "core_math.cuh"

namespace peb {

namespace {

constexpr int kSacMaxModels = 128;  // candidate planes per counting pass (4 floats each in shared memory)

// boost::mt19937 (== std::mt19937), restated: the sample sequence must not depend on a library version
struct Mt19937 {
  uint32_t mt[624];
  int idx;
  explicit Mt19937(uint32_t seed) {
    mt[0] = seed;
    for (int i = 1; i < 624; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + static_cast<uint32_t>(i);
    idx = 624;
  }
  uint32_t next() {
    if (idx >= 624) {
      for (int i = 0; i < 624; ++i) {
        const uint32_t y = (mt[i] & 0x80000000u) | (mt[(i + 1) % 624] & 0x7fffffffu);
        mt[i] = mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
      }
      idx = 0;
    }
    uint32_t y = mt[idx++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
  }
};

// [PCL] sac_model.h : drawIndexSample on the persistent shuffle.  Only the entries that were ever
// swapped are stored (the identity elsewhere), so a run over a 2.3 M-point cloud touches a few hundred.
struct SparseShuffle {
  size_t n;
  std::unordered_map<size_t, int> moved;  // position -> value
  explicit SparseShuffle(size_t n_) : n(n_) {}
  int get(size_t pos) const {
    const auto it = moved.find(pos);
    return it == moved.end() ? static_cast<int>(pos) : it->second;
  }
  void swap(size_t a, size_t b) {
    const int va = get(a), vb = get(b);
    moved[a] = vb;
    moved[b] = va;
  }
};

__global__ void sac_gather_kernel(const float4* __restrict__ pts, const int* __restrict__ idx, int m, float4* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m) out[i] = pts[idx[i]];
}

// [PCL] sac_model_plane.hpp : countWithinDistance for up to kSacMaxModels planes at once.
// |((a x + b y) + c z) + d| < threshold, float products and sums in that order, compared in double.
__global__ void __launch_bounds__(256) sac_count_kernel(const float4* __restrict__ pts, int n, const float4* __restrict__ models,
                                                        int m, double threshold, unsigned* __restrict__ counts) {
  __shared__ float4 s_model[kSacMaxModels];
  __shared__ unsigned s_count[kSacMaxModels];
  for (int k = threadIdx.x; k < m; k += blockDim.x) {
    s_model[k] = models[k];
    s_count[k] = 0u;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i - lane < n; i += gridDim.x * blockDim.x) {
    const bool in = i < n;
    float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
    if (in) p = pts[i];
    for (int k = 0; k < m; ++k) {
      const float4 c = s_model[k];
      const float d = ((c.x * p.x + c.y * p.y) + c.z * p.z) + c.w * 1.0f;
      const bool hit = in && static_cast<double>(fabsf(d)) < threshold;
      const unsigned b = __ballot_sync(0xFFFFFFFFu, hit);
      if (lane == 0 && b) atomicAdd(&s_count[k], static_cast<unsigned>(__popc(b)));
    }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < m; k += blockDim.x)
    if (s_count[k]) atomicAdd(&counts[k], s_count[k]);
}

// [PCL] selectWithinDistance: flag per point (the compaction keeps index order)
__global__ void __launch_bounds__(256) sac_flag_kernel(const float4* __restrict__ pts, int n, float4 c, double threshold,
                                                       uint32_t* __restrict__ flags) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = pts[i];
  const float d = ((c.x * p.x + c.y * p.y) + c.z * p.z) + c.w * 1.0f;
  flags[i] = static_cast<double>(fabsf(d)) < threshold ? 1u : 0u;
}

__global__ void __launch_bounds__(256) sac_indices_kernel(int n, const uint32_t* __restrict__ flags,
                                                          const uint32_t* __restrict__ slot, int32_t* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && flags[i]) out[slot[i]] = i;
}

// Moments of the flagged points for computeMeanAndCovarianceMatrix: xx xy xz yy yz zz x y z, in double.
// One record of 9 doubles per block; the host adds the records in block order (deterministic).
constexpr int kSacMomentBlocks = 592;  // 4 per SM
__global__ void __launch_bounds__(256) sac_moments_kernel(const float4* __restrict__ pts, int n, const uint32_t* __restrict__ flags,
                                                          double* __restrict__ records) {
  double a[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (!flags[i]) continue;
    const float4 p = pts[i];
    const double x = p.x, y = p.y, z = p.z;
    a[0] += x * x;
    a[1] += x * y;
    a[2] += x * z;
    a[3] += y * y;
    a[4] += y * z;
    a[5] += z * z;
    a[6] += x;
    a[7] += y;
    a[8] += z;
  }
  __shared__ double sm[8][9];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    double v = a[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    if (lane == 0) sm[warp][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < 9) {
    double r = 0.0;
    for (int w = 0; w < 8; ++w) r += sm[w][threadIdx.x];
    records[blockIdx.x * 9 + threadIdx.x] = r;
  }
}

// ---- host arithmetic of sac_model_plane.hpp (plain float expressions; this file is compiled with
//      -fmad=false, and the host compiler does not contract across statements at -O3 without -ffast-math;
//      the products below are rounded to float explicitly through volatile-free temporaries) ----------
bool plane_sample_collinear(const float4& p0, const float4& p1, const float4& p2) {
  // Eigen::Array4f dy1dy2 = (p1 - p0) / (p2 - p0): x, y, z lanes
  const float r0 = (p1.x - p0.x) / (p2.x - p0.x);
  const float r1 = (p1.y - p0.y) / (p2.y - p0.y);
  const float r2 = (p1.z - p0.z) / (p2.z - p0.z);
  return (r0 == r1) && (r2 == r1);
}

bool plane_from_sample(const float4& p0, const float4& p1, const float4& p2, float mc[4]) {
  if (plane_sample_collinear(p0, p1, p2)) return false;
  const float ax = p1.x - p0.x, ay = p1.y - p0.y, az = p1.z - p0.z;
  const float bx = p2.x - p0.x, by = p2.y - p0.y, bz = p2.z - p0.z;
  const float m0a = ay * bz, m0b = az * by, m1a = az * bx, m1b = ax * bz, m2a = ax * by, m2b = ay * bx;
  mc[0] = m0a - m0b;
  mc[1] = m1a - m1b;
  mc[2] = m2a - m2b;
  mc[3] = 0.0f;
  const float s0 = mc[0] * mc[0], s1 = mc[1] * mc[1], s2 = mc[2] * mc[2], s3 = mc[3] * mc[3];
  const float nn = std::sqrt(((s0 + s1) + s2) + s3);
  for (int i = 0; i < 4; ++i) mc[i] /= nn;
  const float t0 = mc[0] * p0.x, t1 = mc[1] * p0.y, t2 = mc[2] * p0.z, t3 = mc[3] * 1.0f;
  mc[3] = -1.0f * (((t0 + t1) + t2) + t3);
  return true;
}

}  // namespace

int sac_plane_device(peb_ctx* ctx, const float4* d_pts, int n, const peb_sac_params* prm, float out_coeff[4],
                     int32_t* d_out_inliers, size_t* out_n_inliers, int32_t* out_iterations) {
  for (int i = 0; i < 4; ++i) out_coeff[i] = 0.0f;
  *out_n_inliers = 0;
  if (out_iterations) *out_iterations = 0;
  if (prm->max_iterations < 0) return fail(ctx, PEB_E_INVALID_ARG, "sac_plane: max_iterations < 0");
  if (!(prm->probability > 0.0 && prm->probability < 1.0))
    return fail(ctx, PEB_E_INVALID_ARG, "sac_plane: probability must lie in (0, 1)");
  if (n < 3) return PEB_OK;  // PCL: "Can not select 3 unique points out of n" -> no model

  // ---- replay state of RandomSampleConsensus::computeModel ----
  Mt19937 rng(prm->seed);
  SparseShuffle shuffled(static_cast<size_t>(n));
  int iterations = 0;
  int n_best = -std::numeric_limits<int>::max();
  double k = 1.0;
  const double log_probability = std::log(1.0 - prm->probability);
  const double one_over_indices = 1.0 / static_cast<double>(n);
  unsigned skipped = 0;
  const unsigned max_skip = static_cast<unsigned>(prm->max_iterations) * 10u;
  float best[4] = {0, 0, 0, 0};
  bool have = false, stop = false;

  // scratch: indices (3 ints per drawn triple), gathered points, models, counts
  const int kDraw = 3 * kSacMaxModels;  // triples per gather
  PEB_CUDA(ctx, ctx->d_scratch.ensure(static_cast<size_t>(kDraw) * 3 * (sizeof(int) + sizeof(float4)) +
                                      kSacMaxModels * (sizeof(float4) + sizeof(unsigned)) + kSacMomentBlocks * 9 * sizeof(double) + 256));
  PEB_CUDA(ctx, ctx->h_sac.ensure(static_cast<size_t>(kDraw) * 3 * (sizeof(int) + sizeof(float4)) +
                                    kSacMaxModels * (sizeof(float4) + sizeof(unsigned)) + kSacMomentBlocks * 9 * sizeof(double) + 256));
  char* dbase = ctx->d_scratch.as<char>();
  char* hbase = ctx->h_sac.as<char>();
  size_t off = 0;
  auto carve = [&](size_t bytes) {
    const size_t o = off;
    off += (bytes + 255) & ~static_cast<size_t>(255);
    return o;
  };
  const size_t o_idx = carve(kDraw * 3 * sizeof(int)), o_pts = carve(kDraw * 3 * sizeof(float4));
  const size_t o_models = carve(kSacMaxModels * sizeof(float4)), o_counts = carve(kSacMaxModels * sizeof(unsigned));
  const size_t o_mom = carve(kSacMomentBlocks * 9 * sizeof(double));
  PEB_CUDA(ctx, ctx->d_scratch.ensure(off));
  PEB_CUDA(ctx, ctx->h_sac.ensure(off));
  dbase = ctx->d_scratch.as<char>();
  hbase = ctx->h_sac.as<char>();
  int* h_idx = reinterpret_cast<int*>(hbase + o_idx);
  float4* h_pts = reinterpret_cast<float4*>(hbase + o_pts);
  float4* h_models = reinterpret_cast<float4*>(hbase + o_models);
  unsigned* h_counts = reinterpret_cast<unsigned*>(hbase + o_counts);
  int* d_idx = reinterpret_cast<int*>(dbase + o_idx);
  float4* d_spts = reinterpret_cast<float4*>(dbase + o_pts);
  float4* d_models = reinterpret_cast<float4*>(dbase + o_models);
  unsigned* d_counts = reinterpret_cast<unsigned*>(dbase + o_counts);
  double* d_mom = reinterpret_cast<double*>(dbase + o_mom);
  double* h_mom = reinterpret_cast<double*>(hbase + o_mom);

  // Triples drawn but not consumed yet (coordinates gathered), in draw order.
  std::vector<float4> pending;  // 3 points per triple
  size_t pending_pos = 0;
  auto refill = [&]() -> int {
    // keep what is left, draw kDraw more triples, gather their coordinates
    std::vector<float4> rest(pending.begin() + static_cast<long>(pending_pos), pending.end());
    for (int t = 0; t < kDraw; ++t) {
      for (size_t i = 0; i < 3; ++i) {
        const int r = static_cast<int>(rng.next() >> 1);  // boost::uniform_int<>(0, INT_MAX) over a 32-bit engine
        shuffled.swap(i, i + static_cast<size_t>(r) % (static_cast<size_t>(n) - i));
      }
      for (size_t i = 0; i < 3; ++i) h_idx[3 * t + i] = shuffled.get(i);
    }
    PEB_CUDA(ctx, cudaMemcpyAsync(d_idx, h_idx, kDraw * 3 * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    PEB_LAUNCH(ctx, sac_gather_kernel, ceil_div(kDraw * 3, 256), 256, 0, d_pts, d_idx, kDraw * 3, d_spts);
    PEB_CUDA(ctx, cudaMemcpyAsync(h_pts, d_spts, kDraw * 3 * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
    PEB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    rest.insert(rest.end(), h_pts, h_pts + kDraw * 3);
    pending.swap(rest);
    pending_pos = 0;
    return PEB_OK;
  };

  while (!stop) {
    // ---- collect the next batch of candidate planes exactly as the loop would meet them, assuming it keeps running ----
    struct Cand {
      float mc[4];
      unsigned skipped_before;  // computeModelCoefficients failures right before this candidate
    };
    std::vector<Cand> cands;
    bool no_sample = false;
    unsigned pending_skips = 0;
    while (static_cast<int>(cands.size()) < kSacMaxModels && !no_sample) {
      // getSamples: up to max_sample_checks_ (1000) draws until isSampleGood
      bool good = false;
      float4 s0{}, s1{}, s2{};
      for (unsigned it = 0; it < 1000 && !good; ++it) {
        if (pending_pos + 3 > pending.size()) PEB_TRY(refill());
        s0 = pending[pending_pos];
        s1 = pending[pending_pos + 1];
        s2 = pending[pending_pos + 2];
        pending_pos += 3;
        good = !plane_sample_collinear(s0, s1, s2);
      }
      if (!good) {
        no_sample = true;  // "No samples could be selected!": the loop breaks when it gets here
        break;
      }
      Cand c;
      if (!plane_from_sample(s0, s1, s2, c.mc)) {
        ++pending_skips;  // ++skipped_count; continue
        if (skipped + pending_skips >= max_skip + 1000u) break;  // (bounded; the replay below applies the real limit)
        continue;
      }
      c.skipped_before = pending_skips;
      pending_skips = 0;
      cands.push_back(c);
      // the loop cannot run more iterations than max_iterations + 1 in total
      if (iterations + static_cast<int>(cands.size()) > prm->max_iterations + 1) break;
    }
    // ---- count the inliers of all candidates in one pass ----
    const int m = static_cast<int>(cands.size());
    if (m > 0) {
      for (int c = 0; c < m; ++c) h_models[c] = make_float4(cands[c].mc[0], cands[c].mc[1], cands[c].mc[2], cands[c].mc[3]);
      PEB_CUDA(ctx, cudaMemcpyAsync(d_models, h_models, m * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
      PEB_CUDA(ctx, cudaMemsetAsync(d_counts, 0, m * sizeof(unsigned), ctx->stream));
      const int blocks = std::min(ceil_div(n, 256), kSmCount * 8);
      PEB_LAUNCH(ctx, sac_count_kernel, blocks, 256, 0, d_pts, n, d_models, m, prm->distance_threshold, d_counts);
      PEB_CUDA(ctx, cudaMemcpyAsync(h_counts, d_counts, m * sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
      PEB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    // ---- replay: while (iterations_ < k && skipped_count < max_skip) ----
    int c = 0;
    for (;;) {
      if (!(iterations < k && skipped < max_skip)) {
        stop = true;
        break;
      }
      if (c == m) {
        if (no_sample) stop = true;  // getSamples came back empty: break
        break;                       // otherwise: next batch
      }
      // the failed fits in front of this candidate: each one re-tests the loop condition
      bool ended = false;
      for (unsigned s = 0; s < cands[c].skipped_before; ++s) {
        ++skipped;
        if (!(iterations < k && skipped < max_skip)) {
          ended = true;
          break;
        }
      }
      if (ended) {
        stop = true;
        break;
      }
      const int cnt = static_cast<int>(h_counts[c]);
      if (cnt > n_best) {
        n_best = cnt;
        have = true;
        for (int i = 0; i < 4; ++i) best[i] = cands[c].mc[i];
        const double w = static_cast<double>(n_best) * one_over_indices;
        double p_no_outliers = 1.0 - std::pow(w, 3.0);
        p_no_outliers = std::max(std::numeric_limits<double>::epsilon(), p_no_outliers);
        p_no_outliers = std::min(1.0 - std::numeric_limits<double>::epsilon(), p_no_outliers);
        k = log_probability / std::log(p_no_outliers);
      }
      ++iterations;
      ++c;
      if (iterations > prm->max_iterations) {
        stop = true;
        break;
      }
    }
    if (m == 0 && !stop) stop = true;  // nothing could be drawn
  }
  if (out_iterations) *out_iterations = iterations;
  if (!have) return PEB_OK;

  // ---- getInliers / optimizeModelCoefficients / refine ----
  PEB_CUDA(ctx, ctx->vg_flags.ensure(static_cast<size_t>(n) * 4));
  PEB_CUDA(ctx, ctx->vg_scan.ensure(static_cast<size_t>(n) * 4));
  uint32_t* flags = ctx->vg_flags.as<uint32_t>();
  uint32_t* slot = ctx->vg_scan.as<uint32_t>();
  uint32_t* d_total = ctx->d_small.as<uint32_t>() + 40;
  uint32_t* h_total = ctx->h_small.as<uint32_t>() + 40;
  auto select = [&](const float c4[4], bool want_indices, size_t* count) -> int {
    PEB_LAUNCH(ctx, sac_flag_kernel, ceil_div(n, 256), 256, 0, d_pts, n, make_float4(c4[0], c4[1], c4[2], c4[3]),
               prm->distance_threshold, flags);
    PEB_TRY(exclusive_scan_u32(ctx, flags, slot, n, d_total));
    if (want_indices && d_out_inliers) PEB_LAUNCH(ctx, sac_indices_kernel, ceil_div(n, 256), 256, 0, n, flags, slot, d_out_inliers);
    PEB_CUDA(ctx, cudaMemcpyAsync(h_total, d_total, 4, cudaMemcpyDeviceToHost, ctx->stream));
    PEB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *count = *h_total;
    return PEB_OK;
  };
  float final_c[4] = {best[0], best[1], best[2], best[3]};
  size_t n_inl = 0;
  PEB_TRY(select(best, !prm->optimize_coefficients, &n_inl));
  if (prm->optimize_coefficients) {
    if (n_inl > 3) {  // "Not enough inliers found to optimize model coefficients" otherwise: same coefficients
      PEB_LAUNCH(ctx, sac_moments_kernel, kSacMomentBlocks, 256, 0, d_pts, n, flags, d_mom);
      PEB_CUDA(ctx, cudaMemcpyAsync(h_mom, d_mom, kSacMomentBlocks * 9 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
      PEB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      double a[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
      for (int b = 0; b < kSacMomentBlocks; ++b)
        for (int q = 0; q < 9; ++q) a[q] += h_mom[b * 9 + q];
      for (int q = 0; q < 9; ++q) a[q] /= static_cast<double>(n_inl);
      float cov[9];
      cov[0] = static_cast<float>(a[0] - a[6] * a[6]);
      cov[1] = static_cast<float>(a[1] - a[6] * a[7]);
      cov[2] = static_cast<float>(a[2] - a[6] * a[8]);
      cov[4] = static_cast<float>(a[3] - a[7] * a[7]);
      cov[5] = static_cast<float>(a[4] - a[7] * a[8]);
      cov[8] = static_cast<float>(a[5] - a[8] * a[8]);
      cov[3] = cov[1];
      cov[6] = cov[2];
      cov[7] = cov[5];
      const float cen[3] = {static_cast<float>(a[6]), static_cast<float>(a[7]), static_cast<float>(a[8])};
      float ev, vec[3];
      eigen33_smallest(cov, ev, vec);
      final_c[0] = vec[0];
      final_c[1] = vec[1];
      final_c[2] = vec[2];
      final_c[3] = 0.0f;
      const float t0 = final_c[0] * cen[0], t1 = final_c[1] * cen[1], t2 = final_c[2] * cen[2], t3 = final_c[3] * 1.0f;
      final_c[3] = -1.0f * (((t0 + t1) + t2) + t3);
    }
    PEB_TRY(select(final_c, true, &n_inl));  // segment(): "Refine inliers"
  }
  for (int i = 0; i < 4; ++i) out_coeff[i] = final_c[i];
  *out_n_inliers = n_inl;
  return PEB_OK;
}

}  // namespace peb
