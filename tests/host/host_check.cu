// host_check.cu — TEST SHIM, not product code.  Compiles the __host__ __device__ arithmetic of
// pose_estimation_b200/csrc/core_math.cuh for the CPU (nvcc host pass, no CUDA API calls) so that
// the -m "not gpu" suite can check the device-side logic — grid ring search + termination bound,
// umeyama / LLS from moment sums, the convergence state machine, eigen33 normals — against the
// oracle without a GPU.  The product never loads this library.
#include <algorithm>
#include <cstring>
#include <vector>

#include "../../pose_estimation_b200/csrc/core_math.cuh"
#include "../../pose_estimation_b200/csrc/nn_graph.cuh"
#include "../../pose_estimation_b200/csrc/nn_upfront.cuh"

using namespace peb;

namespace {

struct HostGrid {
  std::vector<float4> pts;
  std::vector<uint32_t> cell_start;

  GridView v{};
};

void build_grid(const float* t, size_t n, size_t stride_f, float occupancy, float h_override, HostGrid& g) {
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  int nf = 0;
  for (size_t i = 0; i < n; ++i) {
    const float* p = t + i * stride_f;
    if (!finite3(p[0], p[1], p[2])) continue;
    for (int a = 0; a < 3; ++a) {
      mn[a] = fminf(mn[a], p[a]);
      mx[a] = fmaxf(mx[a], p[a]);
    }
    ++nf;
  }
  GridView& v = g.v;
  v.n = nf;
  if (nf == 0) {
    v.h = v.inv_h = 1.0f;
    v.dx = v.dy = v.dz = 1;
    g.cell_start.assign(2, 0);
    g.pts.resize(1);
    v.pts = g.pts.data();
    v.cell_start = g.cell_start.data();
    return;
  }
  float ext[3] = {mx[0] - mn[0], mx[1] - mn[1], mx[2] - mn[2]};
  float e[3] = {ext[0], ext[1], ext[2]};
  std::sort(e, e + 3, [](float a, float b) { return a > b; });
  float maxabs = 0.0f;
  for (int a = 0; a < 3; ++a) maxabs = fmaxf(maxabs, fmaxf(fabsf(mn[a]), fabsf(mx[a])));
  float h = h_override > 0 ? h_override : grid_initial_cell(e, nf, occupancy, maxabs, nullptr);
  v.ox = mn[0];
  v.oy = mn[1];
  v.oz = mn[2];
  v.h = h;
  v.inv_h = 1.0f / h;
  v.dx = (int)floorf(ext[0] / h) + 1;
  v.dy = (int)floorf(ext[1] / h) + 1;
  v.dz = (int)floorf(ext[2] / h) + 1;
  const long long cells = (long long)v.dx * v.dy * v.dz;
  std::vector<std::pair<uint32_t, uint32_t>> kv;
  for (size_t i = 0; i < n; ++i) {
    const float* p = t + i * stride_f;
    if (!finite3(p[0], p[1], p[2])) continue;
    int cx = grid_coord(p[0], v.ox, v.inv_h, v.dx), cy = grid_coord(p[1], v.oy, v.inv_h, v.dy),
        cz = grid_coord(p[2], v.oz, v.inv_h, v.dz);
    kv.push_back({(uint32_t)(((long long)cz * v.dy + cy) * v.dx + cx), (uint32_t)i});
  }
  std::stable_sort(kv.begin(), kv.end(), [](auto& a, auto& b) { return a.first < b.first; });
  g.pts.resize(kv.size());
  g.cell_start.assign(cells + 1, 0);
  for (size_t j = 0; j < kv.size(); ++j) {
    const float* p = t + kv[j].second * stride_f;
    int id = (int)kv[j].second;
    float w;
    memcpy(&w, &id, 4);
    g.pts[j] = make_float4(p[0], p[1], p[2], w);
    g.cell_start[kv[j].first + 1]++;
  }
  for (long long c = 0; c < cells; ++c) g.cell_start[c + 1] += g.cell_start[c];
  v.pts = g.pts.data();
  v.cell_start = g.cell_start.data();
}

// the G = 1 body of nn_search.cuh : grid_nn (that header is device-only because of its shuffles)
NnBest grid_nn_host(const GridView& g, float qx, float qy, float qz, float stop_d2, int* rings) {
  NnBest best;
  best.d2 = INFINITY;
  best.idx = -1;
  best.j = -1;
  *rings = 0;
  if (g.n == 0) return best;
  const int cx = grid_coord(qx, g.ox, g.inv_h, g.dx);
  const int cy = grid_coord(qy, g.oy, g.inv_h, g.dy);
  const int cz = grid_coord(qz, g.oz, g.inv_h, g.dz);
  int r = 1;
  bool full = true;
  for (;;) {
    grid_scan_ring(g, qx, qy, qz, cx, cy, cz, r, full, 0, 1, best);
    bool covers_all;
    const float b2 = grid_ring_bound2(g, qx, qy, qz, cx, cy, cz, r, covers_all);
    *rings = r;
    if (covers_all || best.d2 <= b2 || b2 > stop_d2) break;
    if (r >= kMaxRings) {
      grid_scan_range(g, 0u, static_cast<uint32_t>(g.n), qx, qy, qz, best);
      *rings = -1;
      break;
    }
    ++r;
    full = false;
  }
  return best;
}

}  // namespace

extern "C" {
#define HC_API __attribute__((visibility("default")))

// exact 1-NN of every query through the grid logic; out_rings (nullable) = rings searched
HC_API void hc_grid_nn(const float* tgt, size_t n, size_t tstride, const float* q, size_t nq, size_t qstride,
                       float occupancy, float h_override, float stop_d2, int32_t* out_idx, float* out_d2,
                       int32_t* out_rings) {
  HostGrid g;
  build_grid(tgt, n, tstride / 4, occupancy, h_override, g);
  for (size_t i = 0; i < nq; ++i) {
    const float* p = q + i * (qstride / 4);
    int rings = 0;
    NnBest b = grid_nn_host(g.v, p[0], p[1], p[2], stop_d2, &rings);
    out_idx[i] = b.idx;
    out_d2[i] = b.d2;
    if (out_rings) out_rings[i] = rings;
  }
}

// warm-started search: prev[i] = ORIGINAL index of the candidate handed to query i
HC_API void hc_grid_nn_warm(const float* tgt, size_t n, size_t tstride, const float* q, size_t nq, size_t qstride,
                            float occupancy, const int32_t* prev, float limit_d2, float margin, int32_t* out_idx,
                            float* out_d2, float* out_slack) {
  HostGrid g;
  build_grid(tgt, n, tstride / 4, occupancy, 0.0f, g);
  std::vector<int> pos(n, -1);  // original index -> sorted position
  for (int j = 0; j < g.v.n; ++j) {
    int id;
    memcpy(&id, &g.pts[j].w, 4);
    pos[id] = j;
  }
  for (size_t i = 0; i < nq; ++i) {
    const float* p = q + i * (qstride / 4);
    float slack = 0.0f;
    NnBest b = margin < 0.0f ? grid_nn_warm(g.v, p[0], p[1], p[2], pos[prev[i]], limit_d2)
                             : grid_nn_warm_cert(g.v, p[0], p[1], p[2], pos[prev[i]], limit_d2, margin, &slack);
    out_idx[i] = b.idx;
    out_d2[i] = b.d2;
    out_slack[i] = slack;
  }
}

// the staged variant of the warm search (csrc/nn_upfront.cuh): same interface as hc_grid_nn_warm without the certificate
HC_API void hc_grid_nn_warm_upfront(const float* tgt, size_t n, size_t tstride, const float* q, size_t nq, size_t qstride,
                                    float occupancy, const int32_t* prev, float limit_d2, int32_t* out_idx, float* out_d2,
                                    int rows3) {
  HostGrid g;
  build_grid(tgt, n, tstride / 4, occupancy, 0.0f, g);
  std::vector<int> pos(n, -1);
  for (int j = 0; j < g.v.n; ++j) pos[point_index(g.pts[j])] = j;
  for (size_t i = 0; i < nq; ++i) {
    const float* p = q + i * (qstride / 4);
    NnBest b = rows3 ? grid_nn_warm_upfront<3>(g.v, p[0], p[1], p[2], pos[prev[i]], limit_d2)
                     : grid_nn_warm_upfront<2>(g.v, p[0], p[1], p[2], pos[prev[i]], limit_d2);
    out_idx[i] = b.idx;
    out_d2[i] = b.d2;
  }
}

// the warm search over the target's k-NN graph (csrc/nn_graph.cuh): the rows are built here by brute force with the
// arithmetic and the (distance, position) order of normals.cu : knn_graph_kernel; same interface as hc_grid_nn_warm.
// out_rounds (nullable) is not filled by the search itself: the caller only checks results.
HC_API void hc_grid_nn_warm_graph(const float* tgt, size_t n, size_t tstride, const float* q, size_t nq, size_t qstride,
                                  float occupancy, const int32_t* prev, float limit_d2, int32_t* out_idx, float* out_d2, int mode) {
  // mode 0: every row is scanned; 1: rows that cannot certify are skipped (the library's warm launches); 2: launch 0's
  // candidate — greedy descent from prev — verified by the ball search
  HostGrid g;
  build_grid(tgt, n, tstride / 4, occupancy, 0.0f, g);
  std::vector<int> pos(n, -1);
  for (int j = 0; j < g.v.n; ++j) pos[point_index(g.pts[j])] = j;
  std::vector<KnnRow> rows(std::max(g.v.n, 1));
  std::vector<std::pair<float, int>> cand;
  for (int sj = 0; sj < g.v.n; ++sj) {
    const float4 s = g.pts[sj];
    cand.clear();
    for (int j = 0; j < g.v.n; ++j)
      if (j != sj) cand.push_back({l2_simple(s.x, s.y, s.z, g.pts[j].x, g.pts[j].y, g.pts[j].z), j});
    std::sort(cand.begin(), cand.end());
    KnnRow& r = rows[sj];
    memset(&r, 0, sizeof(r));
    auto d2_of = [&](int k) { return k < (int)cand.size() ? cand[k].first : INFINITY; };
    for (int h = 0; h < kGraphHalves; ++h) {
      for (int k = 0; k < 12; ++k) r.half[h].pos[k] = 12 * h + k < (int)cand.size() ? (uint32_t)cand[12 * h + k].second : (uint32_t)sj;
      for (int c = 0; c < 3; ++c) r.half[h].next2[c] = d2_of(12 * h + 4 * c + 4);
    }
  }
  // mode 3: mode 1 + the flatness certificate (records of knn_aux_of)
  std::vector<float4> aux;
  if (mode == 3) {
    aux.resize(std::max(g.v.n, 1));
    for (int sj = 0; sj < g.v.n; ++sj) aux[sj] = knn_aux_of(g.v, rows.data(), sj);
  }
  for (size_t i = 0; i < nq; ++i) {
    const float* p = q + i * (qstride / 4);
    NnBest b;
    if (mode == 3) {
      b = grid_nn_warm_graph(g.v, rows.data(), p[0], p[1], p[2], pos[prev[i]], limit_d2, true, (i & 1) != 0, aux.data());
    } else if (mode == 2) {
      b = grid_nn_graph_descend(g.v, rows.data(), p[0], p[1], p[2], pos[prev[i]]);
      grid_ball_search(g.v, p[0], p[1], p[2], limit_d2, b);
    } else {
      b = grid_nn_warm_graph(g.v, rows.data(), p[0], p[1], p[2], pos[prev[i]], limit_d2, mode == 1);
    }
    out_idx[i] = b.idx;
    out_d2[i] = b.d2;
  }
}

// the bound-only warm search (csrc/nn_upfront.cuh, staged): bound_d2[i] is handed in per query
HC_API void hc_grid_nn_bounded(const float* tgt, size_t n, size_t tstride, const float* q, size_t nq, size_t qstride,
                               float occupancy, const float* bound_d2, float limit_d2, int32_t* out_idx, float* out_d2) {
  HostGrid g;
  build_grid(tgt, n, tstride / 4, occupancy, 0.0f, g);
  for (size_t i = 0; i < nq; ++i) {
    const float* p = q + i * (qstride / 4);
    NnBest b = grid_nn_bounded_upfront(g.v, p[0], p[1], p[2], bound_d2[i], limit_d2);
    out_idx[i] = b.idx;
    out_d2[i] = b.d2;
  }
}

// An ICP-like sequence of warm searches as a kernel would run them: query set k is searched with the bound
// warm_bound_d2(d2 found for set k - 1, |q_k - q_(k-1)| in float, h).  q: steps x nq x 3 floats.  Set 0 is searched cold.
HC_API void hc_bounded_trajectory(const float* tgt, size_t n, size_t tstride, const float* q, size_t nq, size_t steps,
                                  float occupancy, float limit_d2, int32_t* out_idx, float* out_d2) {
  HostGrid g;
  build_grid(tgt, n, tstride / 4, occupancy, 0.0f, g);
  std::vector<float> d2_old(nq);
  for (size_t k = 0; k < steps; ++k) {
    for (size_t i = 0; i < nq; ++i) {
      const float* p = q + (k * nq + i) * 3;
      NnBest b;
      bool cold = k == 0 || !(d2_old[i] < INFINITY);
      if (!cold) {
        const float* o = q + ((k - 1) * nq + i) * 3;
        const float mx = p[0] - o[0], my = p[1] - o[1], mz = p[2] - o[2];
        const float moved = sqrtf(mx * mx + my * my + mz * mz);
        b = grid_nn_bounded_upfront(g.v, p[0], p[1], p[2], warm_bound_d2(d2_old[i], moved, g.v.h), limit_d2);
      } else {
        int rings;
        b = grid_nn_host(g.v, p[0], p[1], p[2], limit_d2, &rings);
      }
      out_idx[k * nq + i] = b.idx;
      out_d2[k * nq + i] = b.d2;
      d2_old[i] = (b.idx >= 0 && b.d2 <= limit_d2) ? b.d2 : INFINITY;  // a rejected query searches cold next time
    }
  }
}

// The warm ball search once more with counters (tests/debug/warm_search_anatomy.py): the loop of
// core_math.cuh : grid_ball_search, statement for statement.  out4 per query: rows of the ball's bounding box, rows
// that pass the slab test (two cell_start loads each), points scanned, improvements of the candidate.
HC_API void hc_warm_stats(const float* tgt, size_t n, size_t tstride, const float* q, size_t nq, size_t qstride,
                          float occupancy, const int32_t* prev, float limit_d2, int32_t* out_idx, int32_t* out4,
                          float* out_cell) {
  HostGrid hg;
  build_grid(tgt, n, tstride / 4, occupancy, 0.0f, hg);
  const GridView& g = hg.v;
  if (out_cell) *out_cell = g.h;
  std::vector<int> pos(n, -1);
  for (int j = 0; j < g.n; ++j) pos[point_index(hg.pts[j])] = j;
  for (size_t i = 0; i < nq; ++i) {
    const float* p = q + i * (qstride / 4);
    const float qx = p[0], qy = p[1], qz = p[2];
    NnBest best;
    const float4 t0 = g.pts[pos[prev[i]]];
    best.d2 = l2_simple(qx, qy, qz, t0.x, t0.y, t0.z);
    best.idx = point_index(t0);
    best.j = pos[prev[i]];
    int rows_box = 0, rows_scanned = 0, points = 0, improved = 0;
    const float fx = (qx - g.ox) * g.inv_h, fy = (qy - g.oy) * g.inv_h, fz = (qz - g.oz) * g.inv_h;
    const float inv_h2 = g.inv_h * g.inv_h;
    const float pad = 0.001f + 4.8e-7f * static_cast<float>(std::max(g.dx, std::max(g.dy, g.dz)));
    const float R = sqrtf(fminf(best.d2, limit_d2) * inv_h2) * 1.0001f + pad;
    const int y0 = grid_clamp_cell(fy - R, g.dy), y1 = grid_clamp_cell(fy + R, g.dy);
    const int z0 = grid_clamp_cell(fz - R, g.dz), z1 = grid_clamp_cell(fz + R, g.dz);
    for (int z = z0; z <= z1; ++z) {
      const float dz = grid_slab_dist_cells(fz, z);
      for (int y = y0; y <= y1; ++y) {
        ++rows_box;
        const float dy = grid_slab_dist_cells(fy, y);
        const float dyz2 = dy * dy + dz * dz;
        const float cur = fminf(best.d2, limit_d2) * inv_h2;
        if (dyz2 > cur) continue;
        ++rows_scanned;
        const float rx = sqrtf(cur - dyz2) * 1.0001f + pad;
        const int x0 = grid_clamp_cell(fx - rx, g.dx), x1 = grid_clamp_cell(fx + rx, g.dx);
        const int base = (z * g.dy + y) * g.dx;
        const uint32_t s = g.cell_start[base + x0], e = g.cell_start[base + x1 + 1];
        points += static_cast<int>(e - s);
        const int before = best.j;
        grid_scan_range(g, s, e, qx, qy, qz, best);
        improved += best.j != before;
      }
    }
    out_idx[i] = best.idx;
    out4[4 * i] = rows_box;
    out4[4 * i + 1] = rows_scanned;
    out4[4 * i + 2] = points;
    out4[4 * i + 3] = improved;
  }
}

// seeded search: query i is seeded with the exact match of seed query s[i] (found by the ring search)
HC_API void hc_grid_nn_seeded(const float* tgt, size_t n, size_t tstride, const float* q, const float* sq, size_t nq,
                              float occupancy, float limit_d2, int32_t* out_idx, float* out_d2) {
  HostGrid g;
  build_grid(tgt, n, tstride / 4, occupancy, 0.0f, g);
  for (size_t i = 0; i < nq; ++i) {
    int rings;
    NnBest sb = grid_nn_host(g.v, sq[3 * i], sq[3 * i + 1], sq[3 * i + 2], limit_d2, &rings);
    NnBest b;
    if (sb.j >= 0)
      b = grid_nn_seeded(g.v, q[3 * i], q[3 * i + 1], q[3 * i + 2], sb.j, sq[3 * i], sq[3 * i + 1], sq[3 * i + 2], limit_d2);
    else
      b = grid_nn_host(g.v, q[3 * i], q[3 * i + 1], q[3 * i + 2], limit_d2, &rings);
    out_idx[i] = b.idx;
    out_d2[i] = b.d2;
  }
}

// one ICP iteration's solve from explicit pairs, through the same moment sums the kernel builds
HC_API void hc_umeyama_pairs(const float* s3, const float* t3, size_t n, float* out_T) {
  double acc[kAccMax] = {0};
  for (size_t i = 0; i < n; ++i) {
    const double sx = s3[3 * i], sy = s3[3 * i + 1], sz = s3[3 * i + 2];
    const double tx = t3[3 * i], ty = t3[3 * i + 1], tz = t3[3 * i + 2];
    acc[0] += 1.0;
    acc[1] += sx; acc[2] += sy; acc[3] += sz;
    acc[4] += tx; acc[5] += ty; acc[6] += tz;
    acc[7] += tx * sx; acc[8] += tx * sy; acc[9] += tx * sz;
    acc[10] += ty * sx; acc[11] += ty * sy; acc[12] += ty * sz;
    acc[13] += tz * sx; acc[14] += tz * sy; acc[15] += tz * sz;
  }
  Mat4 T = umeyama_from_sums(acc);
  memcpy(out_T, T.m, sizeof(T.m));
}

HC_API void hc_lls_pairs(const float* s3, const float* d3, const float* n3, size_t n, float* out_T) {
  double acc[kAccMax] = {0};
  for (size_t i = 0; i < n; ++i) {
    const float sx = s3[3 * i], sy = s3[3 * i + 1], sz = s3[3 * i + 2];
    const float dx = d3[3 * i], dy = d3[3 * i + 1], dz = d3[3 * i + 2];
    const float nx = n3[3 * i], ny = n3[3 * i + 1], nz = n3[3 * i + 2];
    const double a = nz * sy - ny * sz, b = nx * sz - nz * sx, c = ny * sx - nx * sy;
    acc[0] += 1.0;
    acc[1] += a * a; acc[2] += a * b; acc[3] += a * c; acc[4] += a * nx; acc[5] += a * ny; acc[6] += a * nz;
    acc[7] += b * b; acc[8] += b * c; acc[9] += b * nx; acc[10] += b * ny; acc[11] += b * nz;
    acc[12] += c * c; acc[13] += c * nx; acc[14] += c * ny; acc[15] += c * nz;
    acc[16] += nx * nx; acc[17] += nx * ny; acc[18] += nx * nz; acc[19] += ny * ny; acc[20] += ny * nz;
    acc[21] += nz * nz;
    const double d = nx * dx + ny * dy + nz * dz - nx * sx - ny * sy - nz * sz;
    acc[22] += a * d; acc[23] += b * d; acc[24] += c * d; acc[25] += nx * d; acc[26] += ny * d; acc[27] += nz * d;
  }
  Mat4 T = lls_from_sums(acc);
  memcpy(out_T, T.m, sizeof(T.m));
}

// scripted convergence machine: per step an increment (column-major), the mse and the number of
// correspondences; returns state / converged / iterations after every step
HC_API void hc_criteria_script(const peb_icp_params* prm, const float* incs, const double* mses, const int32_t* ncorr,
                               size_t n, int32_t* out_state, int32_t* out_conv, int32_t* out_iter) {
  IcpCriteria cr;
  cr.max_iterations = prm->max_iterations;
  cr.min_correspondences = prm->min_correspondences;
  cr.max_similar = prm->max_iterations_similar;
  cr.estimator = 99;  // neither estimator: the scripted increment is injected below
  cr.mse_abs = prm->abs_mse_threshold;
  cr.mse_rel = prm->euclidean_fitness_epsilon;
  cr.translation_threshold = prm->transformation_epsilon;
  cr.rotation_threshold = prm->rotation_epsilon > 0 ? prm->rotation_epsilon : 1.0 - prm->transformation_epsilon;
  IcpState st;
  memset(&st, 0, sizeof(st));
  st.inc = mat4_identity();
  st.final_t = mat4_identity();
  st.prev_mse = st.cur_mse = DBL_MAX;
  st.active = 1;
  for (size_t i = 0; i < n; ++i) {
    if (st.active) {
      Mat4 inc;
      memcpy(inc.m, incs + 16 * i, sizeof(inc.m));
      icp_finish_iteration_scripted(st, cr, inc, mses[i] * ncorr[i], ncorr[i]);
    }
    out_state[i] = st.state;
    out_conv[i] = st.converged;
    out_iter[i] = st.iterations;
  }
}

HC_API void hc_normal_from_neighbours(const float* nb3, int cnt, const float* p3, const float* vp3, float* out8) {
  float accu[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int j = 0; j < cnt; ++j) {
    const float* c = nb3 + 3 * j;
    accu[0] += c[0] * c[0]; accu[1] += c[0] * c[1]; accu[2] += c[0] * c[2];
    accu[3] += c[1] * c[1]; accu[4] += c[1] * c[2]; accu[5] += c[2] * c[2];
    accu[6] += c[0]; accu[7] += c[1]; accu[8] += c[2];
  }
  normal_from_accu(accu, cnt, p3[0], p3[1], p3[2], vp3[0], vp3[1], vp3[2], out8);
}

HC_API void hc_transforms(const float* T, const float* p3, float* out_icp3, float* out_tpc3) {
  transform_icp(T, p3[0], p3[1], p3[2], out_icp3[0], out_icp3[1], out_icp3[2]);
  transform_tpc(T, p3[0], p3[1], p3[2], out_tpc3[0], out_tpc3[1], out_tpc3[2]);
}

HC_API void hc_mat4_mul(const float* A, const float* B, float* Cm) {
  Mat4 a, b;
  memcpy(a.m, A, 64);
  memcpy(b.m, B, 64);
  Mat4 c = mat4_mul(a, b);
  memcpy(Cm, c.m, 64);
}
}

extern "C" {
// rotation of umeyama from a 3x3 sigma: polar Newton path vs the Jacobi SVD path (returns 1 if polar converged)
__attribute__((visibility("default"))) int hc_rotation_paths(const double* sigma9, double* R_polar, double* R_svd) {
  double A[9], Rp[9], Rs[9];
  for (int i = 0; i < 9; ++i) A[i] = sigma9[i];
  const bool ok = polar_rotation3(A, Rp);
  umeyama_rotation_svd(A, Rs);
  for (int i = 0; i < 9; ++i) {
    R_polar[i] = ok ? Rp[i] : 0.0;
    R_svd[i] = Rs[i];
  }
  return ok ? 1 : 0;
}
}
