// nn_upfront.cuh — the warm ball search with every row bound fetched before the first scan.
//
// STATUS: experimental.  icp.cu instantiates the warm iteration kernels a second time with it (template flag UPF), chosen
// by peb_ctx_set_int(ctx, "warm_upfront", 1); the default stays core_math.cuh : grid_nn_warm, and the default kernels are
// byte-identical in SASS with and without this file.  Never run on a GPU yet (written after the round's GPU time was spent).
// Exactness is checked on the CPU (tests/host/host_check.cu : hc_grid_nn_warm_upfront, tests/test_host_fuzz.py);
// whether it is faster has to be measured on the B200 before it replaces anything.
//
// Why: a warm query is a chain of DEPENDENT L2 round trips — work[i] -> pts[j_prev] -> for every row of the ball:
// (two cell_start loads -> the points of the row), and row k + 1 cannot start before row k is scanned because its chord
// is cut with the distance found so far.  On the C4 clouds (tests/debug/warm_search_anatomy.py, full scale) a query scans
// 3.1 rows on average (1: 19 %, 2: 35 %, 3: 6 %, 4: 32 %, more: 8 %) and 10 points, and its match changes in 18 % of the
// searches only: 8.2 dependent load levels, of which the shrinking chord saves almost nothing.  Here the chords of the
// up to 2 x 2 rows of a small ball are cut with the INITIAL radius (a superset of what the row-after-row walk examines,
// so the result is the same: the comparison (smaller distance, then lower index) does not depend on the order or on
// extra candidates farther than the winner), all eight bounds are loaded at once, and the four ranges are scanned back
// to back: 4 dependent levels, and the same instruction sequence for every lane of a warp (predicated rows instead of
// data-dependent loop trip counts).  Larger balls take the general walk: at full scale 9 % of the warm queries exceed a
// 2 x 2 box (64 % in iteration 1, 28 % in iteration 2, 9 % in iteration 10, 1 % in iteration 29), 2.4 % a 3 x 3 box
// (template parameter RW).
// Checked offline (nvcc 12.9, sm_100a, a kernel that does nothing but this search, __launch_bounds__(128, 8)): 45
// registers and no spills for this variant and for grid_nn_warm alike; in the SASS the four pairs of cell_start loads
// are issued in four predicated regions with no use of their results in between (all eight in flight together), the
// scans follow.
#pragma once

#include "core_math.cuh"

namespace peb {

// RW: the ball's bounding box may span up to RW x RW grid rows (2: 91 % of the warm queries of the C4 clouds; 3: 97.6 %,
// at 18 instead of 8 bound registers)
template <int RW = 2>
PEB_HD void grid_ball_search_upfront(const GridView& g, float qx, float qy, float qz, float limit_d2, NnBest& best) {
  const float fx = (qx - g.ox) * g.inv_h, fy = (qy - g.oy) * g.inv_h, fz = (qz - g.oz) * g.inv_h;
  const float inv_h2 = g.inv_h * g.inv_h;
  const float pad = 0.001f + 4.8e-7f * static_cast<float>(max(g.dx, max(g.dy, g.dz)));  // as in grid_ball_search
  const float cur = fminf(best.d2, limit_d2) * inv_h2;                                   // (cells^2) the initial ball
  const float R = sqrtf(cur) * 1.0001f + pad;
  const int y0 = grid_clamp_cell(fy - R, g.dy), y1 = grid_clamp_cell(fy + R, g.dy);
  const int z0 = grid_clamp_cell(fz - R, g.dz), z1 = grid_clamp_cell(fz + R, g.dz);
  if (y1 - y0 > RW - 1 || z1 - z0 > RW - 1) {  // a large ball: the general walk, which narrows as it goes
    grid_ball_search(g, qx, qy, qz, limit_d2, best);
    return;
  }
  uint32_t s[RW * RW], e[RW * RW];
#pragma unroll
  for (int k = 0; k < RW * RW; ++k) {
    const int y = y0 + (k % RW), z = z0 + (k / RW);
    s[k] = 0u;
    e[k] = 0u;
    if (y <= y1 && z <= z1) {
      const float dy = grid_slab_dist_cells(fy, y), dz = grid_slab_dist_cells(fz, z);
      const float dyz2 = dy * dy + dz * dz;
      if (dyz2 <= cur) {
        const float rx = sqrtf(cur - dyz2) * 1.0001f + pad;
        const int x0 = grid_clamp_cell(fx - rx, g.dx), x1 = grid_clamp_cell(fx + rx, g.dx);
        const int base = (z * g.dy + y) * g.dx;
        s[k] = g.cell_start[base + x0];
        e[k] = g.cell_start[base + x1 + 1];
      }
    }
  }
#pragma unroll
  for (int k = 0; k < RW * RW; ++k) grid_scan_range(g, s[k], e[k], qx, qy, qz, best);
}

// grid_nn_warm with the search above
template <int RW = 2>
PEB_HD NnBest grid_nn_warm_upfront(const GridView& g, float qx, float qy, float qz, int j_prev, float limit_d2) {
  NnBest best;
  const float4 p = g.pts[j_prev];
  best.d2 = l2_simple(qx, qy, qz, p.x, p.y, p.z);
  best.idx = point_index(p);
  best.j = j_prev;
  grid_ball_search_upfront<RW>(g, qx, qy, qz, limit_d2, best);
  return best;
}

// ---- staged, not wired into any kernel: the warm search from a BOUND instead of a candidate ---------------------
// The remaining dependent level in front of the row bounds is the load of the previous match, pts[j_prev], needed only
// to know how large the ball is.  The triangle inequality gives that without touching the target: if the previous
// iteration found the match at distance d_old and the query has moved by m since (both known in the thread: d_old
// would travel in the working record where j_prev travels today, m = |T p - p| is what the certificate path already
// computes), SOME target point lies within d_old + m.  The ball of that radius (a few per cent larger: a warm query
// moves by tens of micrometres against distances of 0.3-1 mm) is scanned from scratch — the previous match is inside
// it and is found by the scan itself.  Chain: work[i] -> row bounds -> points, 3 dependent levels (6.8 today).
// bound_d2 must be a true upper bound of the nearest neighbour's squared distance AS l2_simple COMPUTES IT; the caller
// inflates it (warm_bound_d2 below).  Returns idx = -1 if nothing lies within min(bound_d2, limit_d2) — with a true
// bound that means the match is beyond the rejection limit.
PEB_HD float warm_bound_d2(float d2_old, float moved, float h) {
  // (sqrt(d2_old) + moved)^2, rounded up generously: 1e-5 relative on the radius covers the few ulps of the two
  // computed distances, 1e-6 cell the ulp of the coordinates at the grid's scale (h >= 1e-4 * max |coordinate|)
  const float r = (sqrtf(d2_old) + moved) * 1.00001f + 1e-6f * h;
  return r * r;
}

template <int RW = 2>
PEB_HD NnBest grid_nn_bounded_upfront(const GridView& g, float qx, float qy, float qz, float bound_d2, float limit_d2) {
  NnBest best;
  best.d2 = fminf(bound_d2, limit_d2);
  best.idx = 0x7fffffff;  // not a point: any real point at exactly the bound still wins the (distance, index) comparison
  best.j = -1;
  grid_ball_search_upfront<RW>(g, qx, qy, qz, limit_d2, best);
  if (best.j < 0) best.idx = -1;
  return best;
}

}  // namespace peb
