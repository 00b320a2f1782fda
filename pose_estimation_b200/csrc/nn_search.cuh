// nn_search.cuh — exact 1-NN over the uniform grid, G lanes per query (device only).
// Replaces pcl::KdTreeFLANN::nearestKSearch(p, 1, ...) ([PCL] kdtree/impl/kdtree_flann.hpp,
// [FLANN] algorithms/kdtree_single_index.h) with identical results: same float L2_Simple
// distance, original target index; exact float ties go to the lowest index (FLANN: first
// visited; SURVEY.md 8a-2 allows ties within 1e-6 m).
#pragma once

#include "core_math.cuh"

namespace peb {

__device__ __forceinline__ float pos_inf() { return __int_as_float(0x7f800000); }

template <int G>
__device__ __forceinline__ unsigned group_mask() {
  if (G == 32) return 0xFFFFFFFFu;
  const unsigned lane = threadIdx.x & 31u;
  return ((1u << G) - 1u) << (lane & ~static_cast<unsigned>(G - 1));
}

// stop_d2: correspondences farther than this are rejected by the caller anyway, so rings whose
// lower bound exceeds it need not be searched (pass +inf for plain nearestKSearch semantics).
template <int G>
__device__ __forceinline__ NnBest grid_nn(const GridView& g, float qx, float qy, float qz, float stop_d2) {
  NnBest best;
  best.d2 = pos_inf();
  best.idx = -1;
  best.j = -1;
  if (g.n == 0) return best;
  const int lane_in_group = (G == 1) ? 0 : static_cast<int>(threadIdx.x & (G - 1));
  const unsigned mask = group_mask<G>();
  const int cx = grid_coord(qx, g.ox, g.inv_h, g.dx);
  const int cy = grid_coord(qy, g.oy, g.inv_h, g.dy);
  const int cz = grid_coord(qz, g.oz, g.inv_h, g.dz);
  int r = 1;
  bool full = true;
  for (;;) {
    grid_scan_ring(g, qx, qy, qz, cx, cy, cz, r, full, lane_in_group, G, best);
    if (G > 1) {
#pragma unroll
      for (int o = 1; o < G; o <<= 1) {
        const float od2 = __shfl_xor_sync(mask, best.d2, o);
        const int oidx = __shfl_xor_sync(mask, best.idx, o);
        const int oj = __shfl_xor_sync(mask, best.j, o);
        nn_consider(best, od2, oidx, oj);
      }
    }
    bool covers_all;
    const float b2 = grid_ring_bound2(g, qx, qy, qz, cx, cy, cz, r, covers_all);
    if (covers_all || best.d2 <= b2 || b2 > stop_d2) break;
    if (r >= kMaxRings) {
      // far query: every point, lanes striding over the sorted array (exact, no bound needed)
      for (int j = lane_in_group; j < g.n; j += G) {
        const float4 p = g.pts[j];
        nn_consider(best, l2_simple(qx, qy, qz, p.x, p.y, p.z), __float_as_int(p.w), j);
      }
      if (G > 1) {
#pragma unroll
        for (int o = 1; o < G; o <<= 1) {
          const float od2 = __shfl_xor_sync(mask, best.d2, o);
          const int oidx = __shfl_xor_sync(mask, best.idx, o);
          const int oj = __shfl_xor_sync(mask, best.j, o);
          nn_consider(best, od2, oidx, oj);
        }
      }
      break;
    }
    ++r;
    full = false;
  }
  return best;
}

}  // namespace peb
