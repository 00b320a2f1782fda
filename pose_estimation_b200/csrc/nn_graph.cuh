// nn_graph.cuh — exact warm nearest-neighbour search over a k-nearest-neighbour graph of the TARGET.
//
// The warm searches of a batched align start from last iteration's match s and have to PROVE that nothing is closer to
// the moved query q (or find what is).  The grid walk (core_math.cuh : grid_ball_search) proves it by scanning every
// cell the ball of radius |q - s| touches: ~3 grid rows and ~16 points per query on C4, per-lane loops of different
// lengths, 8 dependent load levels.  The scene does not change during the 30 x 1024 x 50 000 searches of a batch, so
// the neighbourhood of every target point is computed ONCE (normals.cu : knn_graph_kernel):
//
//   row(s) = the sorted positions of the kGraphK nearest other target points of s in ascending (distance, position)
//            order, and the squared distance from s to the first point that a scan of 4, 8 or all kGraphK of them has
//            NOT examined (+inf where there is none).
//
// Certificate.  With d = |q - s|, a point p closer to q than s satisfies |s - p| <= |s - q| + |q - p| < 2 d.  Rows are
// sorted by |s - p|, so once the first unexamined point of a row is farther than 2 d from s, every such p has been
// compared: the best of s and the examined points IS the nearest neighbour.  On C4 the first four neighbours settle
// most late-iteration queries: one 64-byte row and four gathers that hit L1.  Rows that cannot give the certificate
// (2 d beyond the kGraphK-th neighbour) are still scanned completely — a greedy step on the graph that usually lands
// next to the true match, whose own (smaller) ball is then tried — and if a full row does not improve the candidate
// either, the grid walk finishes the search from the best point found (exact for any candidate).  Every comparison
// uses the library's tie rule (smaller squared distance, then lower original index; nn_consider), which does not
// depend on the order candidates are met in, so the result is bit-identical to the grid search's
// (tests/test_gpu_warm_options.py; on the CPU against brute force: tests/test_host_fuzz.py).
//
// Rounding: the row order and the stored distances are the float values l2_simple(s, p) of the kernel that built the
// rows; the triangle inequality is applied with a relative margin of 1e-5 on 4 d^2, three orders of magnitude above
// the rounding of the squared distances involved.
#pragma once

#include "core_math.cuh"

namespace peb {

PEB_HD float graph_inf() { return 3.0e38f * 10.0f; }  // +inf without device intrinsics (host-compiled checks share this file)

constexpr int kGraphHalves = 1;              // 64-byte half rows per row
constexpr int kGraphK = 12 * kGraphHalves;   // neighbours per row

// half h of a row: neighbours 12 h .. 12 h + 11 and the distances that end a scan after 4, 8 and 12 of them.  Most
// queries are settled by the first half; the second one is only touched by queries whose ball reaches past it.
struct __align__(16) KnnHalf {
  uint32_t pos[12];  // sorted positions, ascending (distance, position); a cloud with fewer points fills up with the row's
                     // own position (comparing s with itself changes nothing)
  float next2[3];    // |s - p|^2 of the first neighbour BEHIND chunk 0, 1, 2 of this half (+inf: there is none)
  float spare;
};
struct __align__(16) KnnRow {
  KnnHalf half[kGraphHalves];
};
static_assert(sizeof(KnnRow) == 64 * kGraphHalves, "a half row = two 32-byte sectors");
// (Two halves = 24 neighbours per row, certificates up to d = 1.3 mm on C4 instead of 0.95: measured 14 390 against
//  14 400 hypotheses/s.  Launches 5-12 get 5 % faster, the late ones 12 % slower — rows at a stride of 128 bytes fill
//  L1 / L2 lines with second halves nobody reads — and the graph takes twice as long to build.)
// (Rows that also carry the coordinates of the first four neighbours — no second gather for most queries — were
//  measured: slower, 1.77 against 1.56 ms per late C4 launch.  The copies are private to a row, while gathers from the
//  8 MB point array are shared by neighbouring queries and hit L1.)

// ---- the flatness certificate ------------------------------------------------------------------------------------------
// The triangle inequality proves a match only while 2 d stays inside the row (d < 0.95 mm on C4), although an ICP residual
// mostly points ALONG the surface normal, where nothing else is: the scene is locally a sheet.  Per target point s the
// graph therefore also keeps a unit direction n (the normal of the plane through its row) and
//     H >= |(p - s) . n|  for EVERY target point p within R_s = sqrt(kFlatR2 * outer bound of the row) of s
// (knn_aux_of: an exact ball walk over the grid, once per target).  Write q - s = a n + b (b perpendicular to n, |b| = beta,
// d^2 = a^2 + beta^2) and p - s = eta n + r for a point p the scan has NOT examined, so |p - s| >= r_out (the rows are
// sorted) and, while |p - s| <= R_s, |eta| <= H.  Then
//     |q - p|^2 - d^2 = |p - s|^2 - 2 a eta - 2 b . r  >=  |p - s|^2 - 2 |a| H - 2 beta |p - s|,
// which grows with |p - s| beyond beta: if  r_out^2 - 2 |a| H > 2 beta r_out  no unexamined point within R_s is as close
// to q as s, and beyond R_s the triangle inequality takes over as long as R_s > 2 d.  On a smooth patch (H of the order of
// the sensor noise) that holds for a = 2 mm and beta up to 0.8 mm, where the plain certificate stops at d = 0.95 mm.
// Rounding: |a| and beta are inflated by 0.1 % + 1e-4 cell, H likewise when it is stored, the comparison keeps 0.1 %.
#ifndef PEB_FLAT_R2
#define PEB_FLAT_R2 4.0f
#endif
constexpr float kFlatR2 = PEB_FLAT_R2;  // R_s = 2 x the distance from s to the first point behind its row (C4: 9.0 16 490, 6.25 16 720, 4.0 16 900, 3.0 16 920, 2.25 16 870 hypotheses/s: a smaller ball has a smaller H, a larger one certifies farther queries)

// aux[j] = (n, H) of sorted position j; H = +inf: no certificate (fewer than three points, degenerate row)
PEB_HD float4 knn_aux_of(const GridView& g, const KnnRow* __restrict__ rows, int j) {
  const float4 s = g.pts[j];
  const KnnRow& row = rows[j];
  const float outer = row.half[kGraphHalves - 1].next2[2];
  float4 none = make_float4(0.0f, 0.0f, 1.0f, graph_inf());
  if (!(outer < graph_inf()) || !(outer > 0.0f)) return none;
  // plane through the row: covariance of the offsets of s and its neighbours (s itself contributes the zero offset)
  float m[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  int cnt = 1;
  for (int k = 0; k < 12; ++k) {
    const uint32_t pk = row.half[0].pos[k];
    if (static_cast<int>(pk) == j) continue;
    const float4 p = g.pts[pk];
    const float ox = p.x - s.x, oy = p.y - s.y, oz = p.z - s.z;
    m[0] += ox * ox;
    m[1] += ox * oy;
    m[2] += ox * oz;
    m[3] += oy * oy;
    m[4] += oy * oz;
    m[5] += oz * oz;
    m[6] += ox;
    m[7] += oy;
    m[8] += oz;
    ++cnt;
  }
  if (cnt < 4) return none;
  const float fc = static_cast<float>(cnt);
  for (int i = 0; i < 9; ++i) m[i] = m[i] / fc;
  float cov[9];
  cov[0] = m[0] - m[6] * m[6];
  cov[1] = m[1] - m[6] * m[7];
  cov[2] = m[2] - m[6] * m[8];
  cov[4] = m[3] - m[7] * m[7];
  cov[5] = m[4] - m[7] * m[8];
  cov[8] = m[5] - m[8] * m[8];
  cov[3] = cov[1];
  cov[6] = cov[2];
  cov[7] = cov[5];
  // (in units of the row's extent: eigen33 scales by the largest entry anyway, this keeps denormals out)
  const float sc = 1.0f / outer;
  for (int i = 0; i < 9; ++i) cov[i] *= sc;
  float ev, n[3];
  eigen33_smallest(cov, ev, n);
  const float len2 = n[0] * n[0] + n[1] * n[1] + n[2] * n[2];
  if (!(len2 > 0.25f) || !(len2 < 4.0f)) return none;  // (eigen33 returns unit vectors; anything else: no certificate)
  const float il = 1.0f / sqrtf(len2);
  n[0] *= il;
  n[1] *= il;
  n[2] *= il;
  // H: every target point within R_s of s (exact ball walk; the ball is padded, a point just outside only raises H)
  const float R2 = kFlatR2 * outer * 1.001f;
  const float fx = (s.x - g.ox) * g.inv_h, fy = (s.y - g.oy) * g.inv_h, fz = (s.z - g.oz) * g.inv_h;
  const float pad = 0.001f + 4.8e-7f * static_cast<float>(max(g.dx, max(g.dy, g.dz)));
  const float R = sqrtf(R2) * g.inv_h * 1.0001f + pad;
  const int x0 = grid_clamp_cell(fx - R, g.dx), x1 = grid_clamp_cell(fx + R, g.dx);
  const int y0 = grid_clamp_cell(fy - R, g.dy), y1 = grid_clamp_cell(fy + R, g.dy);
  const int z0 = grid_clamp_cell(fz - R, g.dz), z1 = grid_clamp_cell(fz + R, g.dz);
  float H = 0.0f;
  for (int z = z0; z <= z1; ++z)
    for (int y = y0; y <= y1; ++y) {
      const long long base = (static_cast<long long>(z) * g.dy + y) * g.dx;
      const uint32_t b = g.cell_start[base + x0], e = g.cell_start[base + x1 + 1];
      for (uint32_t k = b; k < e; ++k) {
        const float4 p = g.pts[k];
        const float ox = p.x - s.x, oy = p.y - s.y, oz = p.z - s.z;
        if (ox * ox + oy * oy + oz * oz <= R2) H = fmaxf(H, fabsf(ox * n[0] + oy * n[1] + oz * n[2]));
      }
    }
  return make_float4(n[0], n[1], n[2], H * 1.001f + 1e-4f * g.h);
}

constexpr bool kGraphSkipHopeless = true;  // (measured on C4: +3.7 %; see grid_nn_graph_try)
constexpr int kGraphMaxRounds = 6;  // greedy steps before the grid walk takes over (late iterations need 1-2)

// The graph part of the search: starts from the candidate at sorted position j_prev (last iteration's match) and
// returns true when `best` is PROVEN to be the nearest neighbour of q.  false: `best` is the closest point met so
// far — a valid candidate for grid_ball_search, which is exact for any candidate.  (Split from the walk so that a
// kernel can collect the unproven queries of a tile and walk the grid for them with dense warps: icp.cu.)
// skip_hopeless: when the row of j_prev cannot give the certificate whatever its scan finds (4 d^2 beyond the row's
// outer bound) the scan is skipped and the walk starts from j_prev itself — the twelve gathers of such a row are
// only worth their three dependent rounds when they find a closer point, and late in an align the previous match
// still is the nearest point for four queries out of five.
// (Skipping only rows whose bound is missed by less than a factor — greedy steps first where the previous match is far
//  away, so that the walk covers a smaller ball — measured slower for every factor from 2 to 16: 13 870-15 480 against
//  15 830 hypotheses/s; the steps cost more than the rows they save, also in launch 1.)
PEB_HD bool grid_nn_graph_try(const GridView& g, const KnnRow* __restrict__ rows, float qx, float qy, float qz,
                              int j_prev, NnBest& best, bool skip_hopeless = false, bool peek = false,
                              const float4* __restrict__ aux = nullptr) {
  int js = j_prev;
  // the candidate, the first four positions of its row and the row's three distances: independent loads
  const KnnRow* row = rows + js;
  uint4 p0 = *reinterpret_cast<const uint4*>(row->half[0].pos);
  float4 nx = *reinterpret_cast<const float4*>(row->half[0].next2);
  {
    const float4 s = g.pts[j_prev];
    best.d2 = l2_simple(qx, qy, qz, s.x, s.y, s.z);
    best.idx = point_index(s);
    best.j = j_prev;
  }
  float sx, sy, sz;  // the owner of the current row (= best at the start of a round)
  {
    const float4 s = g.pts[j_prev];
    sx = s.x;
    sy = s.y;
    sz = s.z;
  }
  for (int round = 0; round < kGraphMaxRounds; ++round) {
    const float lim = 4.0f * best.d2 * 1.00001f;
    // the flatness certificate of this row's owner (see the top of the file): flat(r2) with r2 = the squared distance
    // from the owner to the first neighbour a scan has not examined
    float two_ah = graph_inf(), four_b2 = 0.0f;
    bool flat_far = false;
    if (aux) {
      const float4 ax = aux[js];
      const float padf = 1e-4f * g.h;
      const float a = (qx - sx) * ax.x + (qy - sy) * ax.y + (qz - sz) * ax.z;
      const float A = fabsf(a) * 1.001f + padf;
      const float B = sqrtf(fmaxf(best.d2 - a * a, 0.0f)) * 1.001f + padf;  // (best is the row's owner here)
      two_ah = 2.0f * A * ax.w;
      four_b2 = 4.0f * B * B * 1.001f;
      flat_far = lim < kFlatR2 * nx.z;  // R_s > 2 d
    }
    auto flat = [&](float r2) {
      const float t = r2 - two_ah;
      return flat_far && t > 0.0f && t * t > four_b2 * r2;
    };
    if (skip_hopeless && kGraphHalves == 1 && round == 0 && !(nx.z > lim) && !flat(nx.z)) {
      if (peek) {
        // the first launches after launch 0 move the queries by millimetres and the nearest point often has moved on to
        // a neighbour of s: a look at the four nearest ones shrinks the ball the walk has to cover (launch 1 of C4:
        // 5.0 -> 4.6 ms; from launch 3 on it costs more than it saves)
        const uint32_t pos[4] = {p0.x, p0.y, p0.z, p0.w};
        float4 n[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) n[k] = g.pts[pos[k]];
#pragma unroll
        for (int k = 0; k < 4; ++k)
          nn_consider(best, l2_simple(qx, qy, qz, n[k].x, n[k].y, n[k].z), point_index(n[k]), static_cast<int>(pos[k]));
      }
      return false;
    }
    bool proven = false;
#pragma unroll 1
    for (int hf = 0; hf < kGraphHalves && !proven; ++hf) {
      const KnnHalf* half = row->half + hf;
      const float4 nxh = hf == 0 ? nx : *reinterpret_cast<const float4*>(half->next2);
#pragma unroll 1
      for (int c = 0; c < 3 && !proven; ++c) {
        const uint4 pc = (hf == 0 && c == 0) ? p0 : *reinterpret_cast<const uint4*>(half->pos + 4 * c);
        const uint32_t pos[4] = {pc.x, pc.y, pc.z, pc.w};
        float4 n[4];
        float d2[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) n[k] = g.pts[pos[k]];  // (independent gathers)
#pragma unroll
        for (int k = 0; k < 4; ++k) d2[k] = l2_simple(qx, qy, qz, n[k].x, n[k].y, n[k].z);
        // (most chunks hold nothing that beats or ties the candidate: one test instead of four tie rules)
        if (fminf(fminf(d2[0], d2[1]), fminf(d2[2], d2[3])) <= best.d2) {
#pragma unroll
          for (int k = 0; k < 4; ++k) nn_consider(best, d2[k], point_index(n[k]), static_cast<int>(pos[k]));
        }
        // (+inf when nothing is left: everything has been compared)
        const float r2 = c == 0 ? nxh.x : c == 1 ? nxh.y : nxh.z;
        proven = r2 > lim || flat(r2);
      }
    }
    // proven: every point within 2 d of s has been compared, and nothing else can be closer to q than s is
    if (proven) return true;
    if (best.j == js) break;  // no certificate and no better point on the graph: the grid walk decides
    js = best.j;              // a closer point: its ball is smaller, try its row
    if (aux) {
      const float4 s = g.pts[js];
      sx = s.x;
      sy = s.y;
      sz = s.z;
    }
    row = rows + js;
    p0 = *reinterpret_cast<const uint4*>(row->half[0].pos);
    nx = *reinterpret_cast<const float4*>(row->half[0].next2);
  }
  return false;
}

// Greedy descent: from the point at sorted position j_start to a point none of whose kGraphK neighbours is closer to q.
// A CANDIDATE only (the first iteration of a batch verifies it: icp.cu : first_iteration_search) — on a surface scan
// the local minimum usually is the nearest neighbour, at a hole or a depth edge it may not be.
constexpr int kGraphMaxHops = 24;
PEB_HD NnBest grid_nn_graph_descend(const GridView& g, const KnnRow* __restrict__ rows, float qx, float qy, float qz, int j_start) {
  NnBest best;
  {
    const float4 s = g.pts[j_start];
    best.d2 = l2_simple(qx, qy, qz, s.x, s.y, s.z);
    best.idx = point_index(s);
    best.j = j_start;
  }
  int js = j_start;
#pragma unroll 1
  for (int hop = 0; hop < kGraphMaxHops; ++hop) {
    const KnnRow* row = rows + js;
#pragma unroll 1
    for (int hf = 0; hf < kGraphHalves; ++hf) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const uint4 pc = *reinterpret_cast<const uint4*>(row->half[hf].pos + 4 * c);
        const uint32_t pos[4] = {pc.x, pc.y, pc.z, pc.w};
        float4 n[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) n[k] = g.pts[pos[k]];
#pragma unroll
        for (int k = 0; k < 4; ++k)
          nn_consider(best, l2_simple(qx, qy, qz, n[k].x, n[k].y, n[k].z), point_index(n[k]), static_cast<int>(pos[k]));
      }
    }
    if (best.j == js) break;
    js = best.j;
  }
  return best;
}

// Exact 1-NN of q given a candidate at sorted position j_prev (last iteration's match).
PEB_HD NnBest grid_nn_warm_graph(const GridView& g, const KnnRow* __restrict__ rows, float qx, float qy, float qz,
                                 int j_prev, float limit_d2, bool skip_hopeless = false, bool peek = false,
                                 const float4* __restrict__ aux = nullptr) {
  NnBest best;
  if (!grid_nn_graph_try(g, rows, qx, qy, qz, j_prev, best, skip_hopeless, peek, aux)) grid_ball_search(g, qx, qy, qz, limit_d2, best);
  return best;
}

}  // namespace peb
