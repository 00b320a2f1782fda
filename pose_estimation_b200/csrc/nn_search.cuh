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
// FIRST_HIT: return the best point of the first ring that holds any (a SEED for other searches, not a result:
// the rings needed to prove it nearest are the most expensive ones)
template <int G, bool FIRST_HIT = false>
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
    if (FIRST_HIT && best.idx >= 0) break;
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

// ---- warp-cooperative verification of 32 nearby queries (first ICP iteration of a batch) ---------
// In the first iteration the queries are millimetres away from the target surface: each one has a
// candidate (from its patch's anchor, core_math.cuh : grid_nn_seed_probe) and must verify that the
// ball of the candidate's distance holds nothing closer — hundreds of grid rows per query, and the
// 32 balls of a 32-point patch overlap almost entirely.  Here the warp walks the rows of ONE region
// that contains all 32 balls (the bounding ball of their union) with one row per lane, stages the
// points found in a shared-memory tile (coalesced copies), and every lane scans the tile: each row is
// looked up once per warp instead of once per lane, with uniform control flow.  The result is the same
// as 32 independent ball searches (same distance arithmetic, same tie rule, which does not depend on
// the order of examination).
constexpr int kCoopTile = 128;
#ifdef PEB_COOP_STATS
// development build only (make EXTRA=-DPEB_COOP_STATS): patches, fallbacks, rows, non-empty rows, staged points
__device__ unsigned long long g_coop_stats[8];
#define PEB_COOP_COUNT(k, v) do { const unsigned long long v_ = static_cast<unsigned long long>(v); if ((threadIdx.x & 31) == 0) atomicAdd(&g_coop_stats[k], v_); } while (0)
#define PEB_COOP_COUNT_LANE(k) atomicAdd(&g_coop_stats[k], 1ull)
#else
#define PEB_COOP_COUNT(k, v) do { } while (0)
#define PEB_COOP_COUNT_LANE(k) do { } while (0)
#endif
struct alignas(16) CoopTile {
  float xs[kCoopTile], ys[kCoopTile], zs[kCoopTile];  // structure of arrays: two points per packed f32x2 instruction
  int ids[kCoopTile];                                 // original index
  int pos[kCoopTile];                                 // sorted position of the staged point
};

// Packed single precision (sm_100: FADD2 / FMUL2 / FFMA2 work on two floats in a 64-bit register pair with
// one issue slot).  ptxas contracts mul.f32x2 + add.f32x2 into FFMA2 even under --fmad=false, so the packed
// distance below is NOT PCL's arithmetic: it is used as a filter only (see CoopTile scan), and every candidate
// that passes is re-evaluated with l2_simple.
__device__ __forceinline__ unsigned long long f32x2_splat(float v) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %1};" : "=l"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ unsigned long long f32x2_sub(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ unsigned long long f32x2_mul(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ unsigned long long f32x2_fma(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

__device__ __forceinline__ float warp_min_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xFFFFFFFFu, v, o));
  return v;
}
__device__ __forceinline__ float warp_max_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xFFFFFFFFu, v, o));
  return v;
}

// Must be called by all 32 lanes.  `need`: this lane holds a finite query and a candidate in `best`
// that still has to be verified.  Returns false (nothing done) if the region is larger than max_rows
// grid rows or unbounded; the caller then verifies every lane on its own (grid_ball_search).
__device__ __forceinline__ bool grid_nn_coop_verify(const GridView& g, CoopTile* __restrict__ tile, bool need, float qx,
                                                    float qy, float qz, float limit_d2, int max_rows, NnBest& best) {
  constexpr unsigned kFull = 0xFFFFFFFFu;
  const int lane = threadIdx.x & 31;
  if (__ballot_sync(kFull, need) == 0u) return true;
  const float big = 3.0e38f;
  const float lox = warp_min_f(need ? qx : big), hix = warp_max_f(need ? qx : -big);
  const float loy = warp_min_f(need ? qy : big), hiy = warp_max_f(need ? qy : -big);
  const float loz = warp_min_f(need ? qz : big), hiz = warp_max_f(need ? qz : -big);
  const float cx = 0.5f * (lox + hix), cy = 0.5f * (loy + hiy), cz = 0.5f * (loz + hiz);
  const float pad = 0.001f * g.h;  // >> one ulp of any coordinate (h >= 1e-4 * max |coordinate|)
  float reach = 0.0f;
  if (need) {
    const float r = sqrtf(fminf(best.d2, limit_d2)) * 1.0001f + pad;          // this lane's ball (grid_ball_search)
    reach = sqrtf(l2_simple(qx, qy, qz, cx, cy, cz)) * 1.0001f + pad + r;     // triangle inequality, rounded up
  }
  const float Ru = warp_max_f(reach);
  PEB_COOP_COUNT(0, 1);
  if (!(Ru < big)) return false;
  const int y0 = grid_coord(cy - Ru, g.oy, g.inv_h, g.dy), y1 = grid_coord(cy + Ru, g.oy, g.inv_h, g.dy);
  const int z0 = grid_coord(cz - Ru, g.oz, g.inv_h, g.dz), z1 = grid_coord(cz + Ru, g.oz, g.inv_h, g.dz);
  const int ny = y1 - y0 + 1;
  const int nrows = ny * (z1 - z0 + 1);
  if (nrows > max_rows) {
    PEB_COOP_COUNT(1, 1);
    return false;
  }
  PEB_COOP_COUNT(2, nrows);
  PEB_COOP_COUNT(5, __popc(__ballot_sync(kFull, need)));
  const float Ru2 = Ru * Ru;
  // lanes with nothing to verify scan along, but can never accept a point
  const NnBest own = best;
  if (!need) {
    qx = cx;
    qy = cy;
    qz = cz;
    best.d2 = -1.0f;
  }
  int fill = 0;   // points staged in the tile (uniform)
  int slot = -1;  // tile slot of this lane's current winner, if it came from the tile being scanned
  // Scan of the staged points, two per iteration in packed arithmetic.  The packed squared distance uses
  // fused multiply-adds, so it may differ from l2_simple by a few ulps (each is within 2 ulps of the exact
  // value): it only FILTERS — a pair that comes within 1e-6 relative of the lane's best is re-evaluated with
  // l2_simple and goes through the usual comparison (smaller distance, then lower index).
  const unsigned long long qx2 = f32x2_splat(qx), qy2 = f32x2_splat(qy), qz2 = f32x2_splat(qz);
  auto flush = [&]() {
    if (lane < ((4 - (fill & 3)) & 3)) {  // pad to a multiple of four with copies of the last point (a copy never wins a comparison)
      tile->xs[fill + lane] = tile->xs[fill - 1];
      tile->ys[fill + lane] = tile->ys[fill - 1];
      tile->zs[fill + lane] = tile->zs[fill - 1];
      tile->ids[fill + lane] = tile->ids[fill - 1];
      tile->pos[fill + lane] = tile->pos[fill - 1];
    }
    __syncwarp();
    float thresh = best.d2 * 1.000001f + 1e-37f;
    for (int j = 0; j < fill; j += 4) {
      // four points per trip: two independent packed chains (one LDS.128 per coordinate)
      const ulonglong2 x4 = *reinterpret_cast<const ulonglong2*>(tile->xs + j);
      const ulonglong2 y4 = *reinterpret_cast<const ulonglong2*>(tile->ys + j);
      const ulonglong2 z4 = *reinterpret_cast<const ulonglong2*>(tile->zs + j);
      const unsigned long long dxa = f32x2_sub(qx2, x4.x), dxb = f32x2_sub(qx2, x4.y);
      const unsigned long long dya = f32x2_sub(qy2, y4.x), dyb = f32x2_sub(qy2, y4.y);
      const unsigned long long dza = f32x2_sub(qz2, z4.x), dzb = f32x2_sub(qz2, z4.y);
      const unsigned long long da = f32x2_fma(dza, dza, f32x2_fma(dya, dya, f32x2_mul(dxa, dxa)));
      const unsigned long long db = f32x2_fma(dzb, dzb, f32x2_fma(dyb, dyb, f32x2_mul(dxb, dxb)));
      float d0, d1, d2_, d3;
      asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(da));
      asm("mov.b64 {%0, %1}, %2;" : "=f"(d2_), "=f"(d3) : "l"(db));
      if (fminf(fminf(d0, d1), fminf(d2_, d3)) <= thresh) {
        // rare per lane, but some lane of the warp passes on most trips: only the elements that pass are re-evaluated
        const float dk[4] = {d0, d1, d2_, d3};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (dk[k] <= thresh) {
            const float d2 = l2_simple(qx, qy, qz, tile->xs[j + k], tile->ys[j + k], tile->zs[j + k]);
            const int id = tile->ids[j + k];
            if (d2 < best.d2 || (d2 == best.d2 && id < best.idx)) {
              best.d2 = d2;
              best.idx = id;
              slot = j + k;
              thresh = d2 * 1.000001f + 1e-37f;
            }
          }
        }
      }
    }
    if (slot >= 0) best.j = tile->pos[slot];
    slot = -1;
    fill = 0;
    __syncwarp();
  };
  for (int row0 = 0; row0 < nrows; row0 += 32) {
    const int row = row0 + lane;
    uint32_t s = 0, e = 0;
    if (row < nrows) {
      const int zi = row / ny;
      const int y = y0 + (row - zi * ny), z = z0 + zi;
      const float dy = grid_slab_dist(cy, g.oy, g.h, y), dz = grid_slab_dist(cz, g.oz, g.h, z);
      const float dyz2 = dy * dy + dz * dz;
      if (dyz2 <= Ru2) {
        const float rx = sqrtf(Ru2 - dyz2) * 1.0001f + pad;
        const int x0 = grid_coord(cx - rx, g.ox, g.inv_h, g.dx), x1 = grid_coord(cx + rx, g.ox, g.inv_h, g.dx);
        const int base = (z * g.dy + y) * g.dx;
        s = g.cell_start[base + x0];
        e = g.cell_start[base + x1 + 1];
      }
    }
    // The points of these 32 rows are staged as ONE list, 32 points per step whatever the rows' lengths (rows hold
    // 0-15 points: staging them one row per step kept a third of the lanes busy and made every row a load round trip
    // of its own): inclusive scan of the lengths, then every lane finds the row of its list element by a binary
    // search over the lanes' scan values.  The tile receives the points in the same order as before.
    const int len = e > s ? static_cast<int>(e - s) : 0;
    int end = len;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(kFull, end, o);
      if (lane >= o) end += v;
    }
    const int total = __shfl_sync(kFull, end, 31);
    PEB_COOP_COUNT(3, __popc(__ballot_sync(kFull, len > 0)));
    PEB_COOP_COUNT(4, total);
    for (int t0 = 0; t0 < total; t0 += 32) {
      if (fill + 32 > kCoopTile - 4) flush();  // (the scan pads to a multiple of four)
      const int t = t0 + lane;
      int lo = 0;  // the first row whose scan value exceeds t
#pragma unroll
      for (int step = 16; step > 0; step >>= 1) {
        const int ev = __shfl_sync(kFull, end, lo + step - 1);
        if (ev <= t) lo += step;
      }
      const int src_lane = min(lo, 31);
      const int end_r = __shfl_sync(kFull, end, src_lane);
      const int len_r = __shfl_sync(kFull, len, src_lane);
      const uint32_t s_r = __shfl_sync(kFull, s, src_lane);
      if (t < total) {
        const uint32_t j = s_r + static_cast<uint32_t>(t - (end_r - len_r));
        const float4 pt = g.pts[j];
        tile->xs[fill + lane] = pt.x;
        tile->ys[fill + lane] = pt.y;
        tile->zs[fill + lane] = pt.z;
        tile->ids[fill + lane] = __float_as_int(pt.w);
        tile->pos[fill + lane] = static_cast<int>(j);
      }
      fill += min(32, total - t0);
    }
  }
  flush();
  if (!need) best = own;
  return true;
}

}  // namespace peb
