// nn_cache.cuh — the candidate cache of the warm ICP iterations (icp.cu : icp_iteration_cached_kernel).
//
// From the second or third iteration on a query moves by a fraction of a cell per iteration, so its nearest neighbour
// is almost always one of the few target points that were nearest the last time somebody looked.  A cache entry keeps
// the kCacheK nearest points found around a position q0 together with a certificate: every target point that is NOT in
// the entry is at least `bound` away from q0 (it was either examined and ranked behind the kept ones, or it lies
// outside the ball of radius R that the collecting search covered completely).  For a query now at q, any such point x
// obeys |q - x| >= |q0 - x| - |q - q0| >= bound - drift (triangle inequality); if the nearest KEPT point is strictly
// closer than that, it is the exact nearest neighbour — found with kCacheK gathers and no grid walk, the same for every
// lane of the warp.  Otherwise the query is collected again (grid_ball_collect), by a dense warp of such queries.
// Exactness does not depend on R, on kCacheK or on how often entries are rebuilt: they only move work between the two
// paths, so results are bit-identical to the plain warm search (tests/test_gpu_parity.py :
// test_speed_options_never_change_results).
//
// Float safety: coordinate differences of nearby points are exact (Sterbenz), the squared sums carry a few ulps; the
// bound is shrunk and the test padded by 4e-6 relative + 1e-6 cell, far above that and far below anything that matters
// for the hit rate.
#pragma once

#include <cmath>

#include "core_math.cuh"

namespace peb {

constexpr int kCacheK = 4;

struct alignas(16) NnCache {  // 32 bytes per (hypothesis, source point)
  float qx, qy, qz;  // where the candidates were collected
  float bound;       // every target point not among c[] is at least this far from (qx, qy, qz); <= 0: no certificate
  int c[kCacheK];    // sorted positions of the nearest points found there, nearest first (-1: none)
};

// The kCacheK + 1 nearest points by (squared distance, original index), ascending.
struct NnTop {
  float d2[kCacheK + 1];
  int idx[kCacheK + 1];
  int j[kCacheK + 1];
};

PEB_HD void nn_top_insert(NnTop& t, float d2, int idx, int j) {
  auto less = [](float a, int ai, float b, int bi) { return a < b || (a == b && ai < bi); };
  if (!less(d2, idx, t.d2[kCacheK], t.idx[kCacheK])) return;
#pragma unroll
  for (int i = kCacheK; i >= 1; --i) {
    const bool shift = less(d2, idx, t.d2[i - 1], t.idx[i - 1]);
    const bool here = less(d2, idx, t.d2[i], t.idx[i]);
    t.j[i] = shift ? t.j[i - 1] : (here ? j : t.j[i]);
    t.idx[i] = shift ? t.idx[i - 1] : (here ? idx : t.idx[i]);
    t.d2[i] = shift ? t.d2[i - 1] : (here ? d2 : t.d2[i]);
  }
  const bool here = less(d2, idx, t.d2[0], t.idx[0]);
  t.j[0] = here ? j : t.j[0];
  t.idx[0] = here ? idx : t.idx[0];
  t.d2[0] = here ? d2 : t.d2[0];
}

// Examines EVERY target point within r_cells (cell units) of q — the rows of the ball's bounding box, each cut to the
// ball's chord, with the conservative margins of grid_ball_search — and ranks what it sees.  Points beyond the radius
// that happen to lie in an examined cell are ranked too (harmless: the bound below never exceeds the radius).
PEB_HD void grid_ball_collect(const GridView& g, float qx, float qy, float qz, float r_cells, NnTop& top) {
#pragma unroll
  for (int i = 0; i <= kCacheK; ++i) {
    top.d2[i] = HUGE_VALF;
    top.idx[i] = 0x7FFFFFFF;
    top.j[i] = -1;
  }
  const float fx = (qx - g.ox) * g.inv_h, fy = (qy - g.oy) * g.inv_h, fz = (qz - g.oz) * g.inv_h;
  const float pad = 0.001f + 4.8e-7f * static_cast<float>(max(g.dx, max(g.dy, g.dz)));  // as in grid_ball_search
  const float R = r_cells * 1.0001f + pad;
  const float R2 = R * R;
  const int y0 = grid_clamp_cell(fy - R, g.dy), y1 = grid_clamp_cell(fy + R, g.dy);
  const int z0 = grid_clamp_cell(fz - R, g.dz), z1 = grid_clamp_cell(fz + R, g.dz);
  for (int z = z0; z <= z1; ++z) {
    const float dz = grid_slab_dist_cells(fz, z);
    for (int y = y0; y <= y1; ++y) {
      const float dy = grid_slab_dist_cells(fy, y);
      const float dyz2 = dy * dy + dz * dz;
      if (dyz2 > R2) continue;
      const float rx = sqrtf(R2 - dyz2) * 1.0001f + pad;
      const int x0 = grid_clamp_cell(fx - rx, g.dx), x1 = grid_clamp_cell(fx + rx, g.dx);
      const int base = (z * g.dy + y) * g.dx;
      const uint32_t s = g.cell_start[base + x0], e = g.cell_start[base + x1 + 1];
      for (uint32_t j = s; j < e; ++j) {
        const float4 p = g.pts[j];
        nn_top_insert(top, l2_simple(qx, qy, qz, p.x, p.y, p.z), point_index(p), static_cast<int>(j));
      }
    }
  }
}

// The cache entry of a finished collection around q with radius r_cells: bound = min(radius, distance of the first
// point that was NOT kept), shrunk by the float margin.  An empty collection still certifies "nothing within R".
PEB_HD NnCache nn_cache_from_top(const GridView& g, float qx, float qy, float qz, float r_cells, const NnTop& top) {
  NnCache ce;
  ce.qx = qx;
  ce.qy = qy;
  ce.qz = qz;
#pragma unroll
  for (int i = 0; i < kCacheK; ++i) ce.c[i] = top.j[i];
  float b = r_cells * g.h;
  if (top.j[kCacheK] >= 0) b = fminf(b, sqrtf(top.d2[kCacheK]));
  ce.bound = b * (1.0f - 4e-6f) - 1e-6f * g.h;
  return ce;
}

// The certificate test: the exact nearest neighbour of q from the entry alone, or false.
PEB_HD bool nn_cache_lookup(const GridView& g, const NnCache& ce, float qx, float qy, float qz, NnBest& best) {
  best.d2 = HUGE_VALF;
  best.idx = -1;
  best.j = -1;
  if (!(ce.bound > 0.0f) || ce.c[0] < 0) return false;
#pragma unroll
  for (int i = 0; i < kCacheK; ++i) {
    const int j = ce.c[i];
    if (j >= 0) {
      const float4 p = g.pts[j];
      nn_consider(best, l2_simple(qx, qy, qz, p.x, p.y, p.z), point_index(p), j);
    }
  }
  const float dx = qx - ce.qx, dy = qy - ce.qy, dz = qz - ce.qz;
  const float drift = sqrtf(dx * dx + dy * dy + dz * dz);
  return (sqrtf(best.d2) + drift) * (1.0f + 4e-6f) < ce.bound;
}

}  // namespace peb
