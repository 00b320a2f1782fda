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
// is cut with the distance found so far.  On the C4 geometry (tests/debug/warm_search_anatomy.py) a query scans 2.4 rows
// on average (1: 27 %, 2: 41 %, 3: 5 %, 4: 25 %, more: 2.6 %) and 9 points, and its match changes in 12 % of the
// searches only: 6.8 dependent load levels, of which the shrinking chord saves almost nothing.  Here the chords of the
// up to 2 x 2 rows of a small ball are cut with the INITIAL radius (a superset of what the row-after-row walk examines,
// so the result is the same: the comparison (smaller distance, then lower index) does not depend on the order or on
// extra candidates farther than the winner), all eight bounds are loaded at once, and the four ranges are scanned back
// to back: 4 dependent levels, and the same instruction sequence for every lane of a warp (predicated rows instead of
// data-dependent loop trip counts).  Larger balls (2.6 % of the warm queries) take the general walk.
// Checked offline (nvcc 12.9, sm_100a, a kernel that does nothing but this search, __launch_bounds__(128, 8)): 45
// registers and no spills for this variant and for grid_nn_warm alike; in the SASS the four pairs of cell_start loads
// are issued in four predicated regions with no use of their results in between (all eight in flight together), the
// scans follow.
#pragma once

#include "core_math.cuh"

namespace peb {

PEB_HD void grid_ball_search_upfront(const GridView& g, float qx, float qy, float qz, float limit_d2, NnBest& best) {
  const float fx = (qx - g.ox) * g.inv_h, fy = (qy - g.oy) * g.inv_h, fz = (qz - g.oz) * g.inv_h;
  const float inv_h2 = g.inv_h * g.inv_h;
  const float pad = 0.001f + 4.8e-7f * static_cast<float>(max(g.dx, max(g.dy, g.dz)));  // as in grid_ball_search
  const float cur = fminf(best.d2, limit_d2) * inv_h2;                                   // (cells^2) the initial ball
  const float R = sqrtf(cur) * 1.0001f + pad;
  const int y0 = grid_clamp_cell(fy - R, g.dy), y1 = grid_clamp_cell(fy + R, g.dy);
  const int z0 = grid_clamp_cell(fz - R, g.dz), z1 = grid_clamp_cell(fz + R, g.dz);
  if (y1 - y0 > 1 || z1 - z0 > 1) {  // a large ball: the general walk, which narrows as it goes
    grid_ball_search(g, qx, qy, qz, limit_d2, best);
    return;
  }
  uint32_t s[4], e[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int y = y0 + (k & 1), z = z0 + (k >> 1);
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
  for (int k = 0; k < 4; ++k) grid_scan_range(g, s[k], e[k], qx, qy, qz, best);
}

// grid_nn_warm with the search above
PEB_HD NnBest grid_nn_warm_upfront(const GridView& g, float qx, float qy, float qz, int j_prev, float limit_d2) {
  NnBest best;
  const float4 p = g.pts[j_prev];
  best.d2 = l2_simple(qx, qy, qz, p.x, p.y, p.z);
  best.idx = point_index(p);
  best.j = j_prev;
  grid_ball_search_upfront(g, qx, qy, qz, limit_d2, best);
  return best;
}

}  // namespace peb
