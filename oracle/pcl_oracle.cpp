// pcl_oracle.cpp — CPU ORACLE.  TEST INFRASTRUCTURE ONLY, NOT PRODUCT CODE.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load this library; the product (libpe_b200.so) never links, loads or calls it.
//
// What it is: a dependency-free C++17 restatement of the PCL 1.10.0 registration path that
// BASELINE.json's north star names (VoxelGrid -> NormalEstimation -> IterativeClosestPoint /
// IterativeClosestPointWithNormals, KdTreeFLANN correspondences).  The reference repo pins
// "PCL 1.10" (pose_estimation/CMakeLists.txt:36, README.md:53-61) but vendors none of it, and
// neither PCL, FLANN nor Eigen exist in this image, so each function below follows the
// published upstream algorithm and names the upstream file it restates ("[PCL] path",
// "[FLANN] path", "[EIGEN] path"; no line numbers: they cannot be checked here).
//
// PARITY UNPINNED (partially): the reference has no tests, fixtures or golden vectors for this
// path and real PCL cannot be run here.  What IS pinned, by tests/test_oracle.py:
//   - the kd-tree search against OpenCV's bundled FLANN KDTreeSingleIndex (cv2.flann_Index,
//     algorithm 4, leaf 15, exact search) and scipy cKDTree and brute force: committed golden
//     vectors in tests/golden/;
//   - umeyama against a float64 numpy Kabsch, the LLS step against numpy lstsq, eigen33 against
//     numpy eigh, VoxelGrid against an independent numpy dict-of-cells restatement.
//   - the plane RANSAC (pcl::SACSegmentation, SACMODEL_PLANE + SAC_RANSAC; tests/test_sac.py): its Mersenne twister
//     against the C++ standard's known answer and numpy's RandomState(12345) stream, the sequential loop against an
//     independent numpy restatement (bit-identical coefficients, iteration counts, inlier lists).
// The remaining statements (operation order, float/double choices, the convergence state
// machine) are recollections of upstream source and carry no external pin.  The same holds for
// the cv::ppf_match_3d::ICP restatement at the end of this file (opencv_contrib is not in the image
// either): it is checked for what an ICP must do (tests/test_cvicp.py), not against OpenCV.
//
// Canonical arithmetic: IEEE-754 binary32/binary64, round-to-nearest, NO fused contraction
// (build with -ffp-contract=off), PCL's source-level operation order.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/pe_b200.h"  // POD params / result layouts only (shared with the tests)

namespace orc {

struct P3 {
  float x, y, z;
};

static inline bool finite3(float x, float y, float z) {
  return std::isfinite(x) && std::isfinite(y) && std::isfinite(z);
}

static inline const float* rec(const void* base, size_t i, size_t stride) {
  return reinterpret_cast<const float*>(static_cast<const char*>(base) + i * stride);
}

// ------------------------------------------------------------------------------------------
// [FLANN] src/cpp/flann/algorithms/dist.h : L2_Simple<float>::operator()
//   result = 0; for each dim: diff = a[i]-b[i]; result += diff*diff;   (all float)
// ------------------------------------------------------------------------------------------
static inline float l2_simple(const float* a, const float* b) {
  float r = 0.0f;
  for (int i = 0; i < 3; ++i) {
    float diff = a[i] - b[i];
    r += diff * diff;
  }
  return r;
}

// ------------------------------------------------------------------------------------------
// [FLANN] src/cpp/flann/util/result_set.h : KNNSimpleResultSet<float>
// ------------------------------------------------------------------------------------------
struct KnnSet {
  int capacity;
  int count = 0;
  float worst;
  float* dist;
  int* index;
  KnnSet(int k, float* d, int* ix) : capacity(k), dist(d), index(ix) {
    worst = std::numeric_limits<float>::max();
    dist[capacity - 1] = worst;
  }
  float worstDist() const { return worst; }
  void addPoint(float d, int id) {
    if (d >= worst) return;
    if (count < capacity) ++count;
    int i;
    for (i = count - 1; i > 0; --i) {
      if (dist[i - 1] > d) {
        dist[i] = dist[i - 1];
        index[i] = index[i - 1];
      } else {
        break;
      }
    }
    dist[i] = d;
    index[i] = id;
    worst = dist[capacity - 1];
  }
};

// ------------------------------------------------------------------------------------------
// [FLANN] src/cpp/flann/algorithms/kdtree_single_index.h : KDTreeSingleIndex<L2_Simple<float>>
// as configured by [PCL] kdtree/include/pcl/kdtree/impl/kdtree_flann.hpp:
//   KDTreeSingleIndexParams(15 /*leaf_max_size*/), reorder = true, SearchParams(checks -1,
//   eps 0) => exact search, sorted results, squared distances.  Non-finite input points are
//   left out of the index and results are mapped back to original indices (index_mapping_).
// ------------------------------------------------------------------------------------------
struct KdTree {
  struct Interval {
    float low, high;
  };
  struct Node {
    int left, right;    // leaf: point range [left,right) in reordered data
    int divfeat;        // inner: split dimension
    float divlow, divhigh;
    int child1 = -1, child2 = -1;
  };
  std::vector<float> pts;      // dense finite points, 3 floats each (index = dense id)
  std::vector<int> mapping;    // dense id -> original index
  std::vector<int> vind;       // permutation built by the split
  std::vector<float> data;     // reordered copy
  std::vector<Node> nodes;
  Interval root_bbox[3];
  int leaf_max = 15;
  int n = 0;

  void build(const void* base, size_t count, size_t stride) {
    pts.clear();
    mapping.clear();
    nodes.clear();
    for (size_t i = 0; i < count; ++i) {
      const float* p = rec(base, i, stride);
      if (!finite3(p[0], p[1], p[2])) continue;
      pts.push_back(p[0]);
      pts.push_back(p[1]);
      pts.push_back(p[2]);
      mapping.push_back(static_cast<int>(i));
    }
    n = static_cast<int>(mapping.size());
    vind.resize(n);
    for (int i = 0; i < n; ++i) vind[i] = i;
    if (n == 0) return;
    for (int d = 0; d < 3; ++d) {
      root_bbox[d].low = root_bbox[d].high = pts[d];
    }
    for (int k = 1; k < n; ++k)
      for (int d = 0; d < 3; ++d) {
        float v = pts[3 * k + d];
        if (v < root_bbox[d].low) root_bbox[d].low = v;
        if (v > root_bbox[d].high) root_bbox[d].high = v;
      }
    nodes.reserve(2 * (n / 8 + 1));
    Interval bbox[3] = {root_bbox[0], root_bbox[1], root_bbox[2]};
    divide(0, n, bbox);
    data.resize(3 * static_cast<size_t>(n));
    for (int i = 0; i < n; ++i)
      for (int d = 0; d < 3; ++d) data[3 * static_cast<size_t>(i) + d] = pts[3 * static_cast<size_t>(vind[i]) + d];
  }

  float coord(int dense, int d) const { return pts[3 * static_cast<size_t>(dense) + d]; }

  void minmax(const int* ind, int count, int dim, float& mn, float& mx) const {
    mn = mx = coord(ind[0], dim);
    for (int i = 1; i < count; ++i) {
      float v = coord(ind[i], dim);
      if (v < mn) mn = v;
      if (v > mx) mx = v;
    }
  }

  void planeSplit(int* ind, int count, int cutfeat, float cutval, int& lim1, int& lim2) const {
    int left = 0, right = count - 1;
    for (;;) {
      while (left <= right && coord(ind[left], cutfeat) < cutval) ++left;
      while (left <= right && coord(ind[right], cutfeat) >= cutval) --right;
      if (left > right) break;
      std::swap(ind[left], ind[right]);
      ++left;
      --right;
    }
    lim1 = left;
    right = count - 1;
    for (;;) {
      while (left <= right && coord(ind[left], cutfeat) <= cutval) ++left;
      while (left <= right && coord(ind[right], cutfeat) > cutval) --right;
      if (left > right) break;
      std::swap(ind[left], ind[right]);
      ++left;
      --right;
    }
    lim2 = left;
  }

  void middleSplit(int* ind, int count, int& index, int& cutfeat, float& cutval, const Interval* bbox) const {
    const float EPS = 0.00001f;
    float max_span = bbox[0].high - bbox[0].low;
    for (int i = 1; i < 3; ++i) {
      float span = bbox[i].high - bbox[i].low;
      if (span > max_span) max_span = span;
    }
    float max_spread = -1;
    cutfeat = 0;
    for (int i = 0; i < 3; ++i) {
      float span = bbox[i].high - bbox[i].low;
      if (span > static_cast<float>((1 - EPS) * max_span)) {
        float mn, mx;
        minmax(ind, count, i, mn, mx);
        float spread = mx - mn;
        if (spread > max_spread) {
          cutfeat = i;
          max_spread = spread;
        }
      }
    }
    float split_val = (bbox[cutfeat].low + bbox[cutfeat].high) / 2;
    float mn, mx;
    minmax(ind, count, cutfeat, mn, mx);
    if (split_val < mn)
      cutval = mn;
    else if (split_val > mx)
      cutval = mx;
    else
      cutval = split_val;
    int lim1, lim2;
    planeSplit(ind, count, cutfeat, cutval, lim1, lim2);
    if (lim1 > count / 2)
      index = lim1;
    else if (lim2 < count / 2)
      index = lim2;
    else
      index = count / 2;
  }

  int divide(int left, int right, Interval* bbox) {
    int id = static_cast<int>(nodes.size());
    nodes.emplace_back();
    if (right - left <= leaf_max) {
      nodes[id].left = left;
      nodes[id].right = right;
      nodes[id].child1 = nodes[id].child2 = -1;
      for (int d = 0; d < 3; ++d) bbox[d].low = bbox[d].high = coord(vind[left], d);
      for (int k = left + 1; k < right; ++k)
        for (int d = 0; d < 3; ++d) {
          float v = coord(vind[k], d);
          if (bbox[d].low > v) bbox[d].low = v;
          if (bbox[d].high < v) bbox[d].high = v;
        }
    } else {
      int idx, cutfeat;
      float cutval;
      middleSplit(&vind[left], right - left, idx, cutfeat, cutval, bbox);
      nodes[id].divfeat = cutfeat;
      Interval lb[3] = {bbox[0], bbox[1], bbox[2]};
      lb[cutfeat].high = cutval;
      int c1 = divide(left, left + idx, lb);
      Interval rb[3] = {bbox[0], bbox[1], bbox[2]};
      rb[cutfeat].low = cutval;
      int c2 = divide(left + idx, right, rb);
      nodes[id].child1 = c1;
      nodes[id].child2 = c2;
      nodes[id].divlow = lb[cutfeat].high;
      nodes[id].divhigh = rb[cutfeat].low;
      for (int d = 0; d < 3; ++d) {
        bbox[d].low = std::min(lb[d].low, rb[d].low);
        bbox[d].high = std::max(lb[d].high, rb[d].high);
      }
    }
    return id;
  }

  void searchLevel(KnnSet& rs, const float* vec, int node, float mindistsq, float* dists) const {
    const Node& nd = nodes[node];
    if (nd.child1 < 0 && nd.child2 < 0) {
      float worst = rs.worstDist();
      for (int i = nd.left; i < nd.right; ++i) {
        float d = l2_simple(vec, &data[3 * static_cast<size_t>(i)]);
        if (d < worst) rs.addPoint(d, vind[i]);
      }
      return;
    }
    int idx = nd.divfeat;
    float val = vec[idx];
    float diff1 = val - nd.divlow;
    float diff2 = val - nd.divhigh;
    int best, other;
    float cut_dist;
    if ((diff1 + diff2) < 0) {
      best = nd.child1;
      other = nd.child2;
      cut_dist = (val - nd.divhigh) * (val - nd.divhigh);  // L2_Simple::accum_dist
    } else {
      best = nd.child2;
      other = nd.child1;
      cut_dist = (val - nd.divlow) * (val - nd.divlow);
    }
    searchLevel(rs, vec, best, mindistsq, dists);
    float dst = dists[idx];
    mindistsq = mindistsq + cut_dist - dst;
    dists[idx] = cut_dist;
    if (mindistsq * 1.0f /*epsError = 1+eps, eps = 0*/ <= rs.worstDist()) searchLevel(rs, vec, other, mindistsq, dists);
    dists[idx] = dst;
  }

  // returns number of neighbours found (min(k, n)); indices are ORIGINAL indices
  int knn(const float* q, int k, int* out_idx, float* out_d2) const {
    if (n == 0) return 0;
    if (k > n) k = n;
    KnnSet rs(k, out_d2, out_idx);
    float dists[3] = {0, 0, 0};
    float distsq = 0;
    for (int i = 0; i < 3; ++i) {
      if (q[i] < root_bbox[i].low) {
        dists[i] = (q[i] - root_bbox[i].low) * (q[i] - root_bbox[i].low);
        distsq += dists[i];
      }
      if (q[i] > root_bbox[i].high) {
        dists[i] = (q[i] - root_bbox[i].high) * (q[i] - root_bbox[i].high);
        distsq += dists[i];
      }
    }
    searchLevel(rs, q, 0, distsq, dists);
    for (int i = 0; i < rs.count; ++i) out_idx[i] = mapping[out_idx[i]];
    return rs.count;
  }
};

// ------------------------------------------------------------------------------------------
// 4x4 float matrices, column-major (Eigen::Matrix4f storage).
// ------------------------------------------------------------------------------------------
struct M4 {
  float m[16];
  float& operator()(int r, int c) { return m[c * 4 + r]; }
  float operator()(int r, int c) const { return m[c * 4 + r]; }
  static M4 identity() {
    M4 a;
    for (int i = 0; i < 16; ++i) a.m[i] = 0;
    a.m[0] = a.m[5] = a.m[10] = a.m[15] = 1;
    return a;
  }
};

// [EIGEN] Matrix4f * Matrix4f: column j of the result = sum_k lhs.col(k) * rhs(k,j), k ascending.
static M4 mul(const M4& a, const M4& b) {
  M4 r;
  for (int j = 0; j < 4; ++j)
    for (int i = 0; i < 4; ++i) {
      float acc = a(i, 0) * b(0, j);
      acc = acc + a(i, 1) * b(1, j);
      acc = acc + a(i, 2) * b(2, j);
      acc = acc + a(i, 3) * b(3, j);
      r(i, j) = acc;
    }
  return r;
}

// [PCL] registration/include/pcl/registration/impl/icp.hpp : IterativeClosestPoint::transformCloud
//   pt_t = tr * pt with pt = (x,y,z,1): Eigen 4x4 * 4-vector = ((c0*x + c1*y) + c2*z) + c3*1.
static inline void transform_icp(const M4& t, const float* p, float* o) {
  for (int r = 0; r < 3; ++r) {
    float acc = t(r, 0) * p[0];
    acc = acc + t(r, 1) * p[1];
    acc = acc + t(r, 2) * p[2];
    acc = acc + t(r, 3) * 1.0f;
    o[r] = acc;
  }
}

// [PCL] common/include/pcl/common/impl/transforms.hpp : pcl::detail::Transformer<float>::se3
//   tgt = c0*x + (c1*y + (c2*z + c3))
static inline void transform_tpc(const M4& t, const float* p, float* o) {
  for (int r = 0; r < 3; ++r) {
    float p0 = p[0] * t(r, 0);
    float p1 = p[1] * t(r, 1);
    float p2 = p[2] * t(r, 2);
    o[r] = p0 + (p1 + (p2 + t(r, 3)));
  }
}

// ------------------------------------------------------------------------------------------
// [EIGEN] Eigen/src/Jacobi/Jacobi.h + Eigen/src/SVD/JacobiSVD.h, 3x3 real case (no QR
// preconditioner for square input): two-sided Jacobi sweeps built from real_2x2_jacobi_svd.
// Templated so the tests can run it in double as a sanity reference.
// ------------------------------------------------------------------------------------------
template <typename S>
struct Rot {
  S c, s;
};

template <typename S>
static bool makeJacobi(S x, S y, S z, Rot<S>& j) {
  S deno = S(2) * std::fabs(y);
  if (deno < std::numeric_limits<S>::min()) {
    j.c = S(1);
    j.s = S(0);
    return false;
  }
  S tau = (x - z) / deno;
  S w = std::sqrt(tau * tau + S(1));
  S t = (tau > S(0)) ? S(1) / (tau + w) : S(1) / (tau - w);
  S sign_t = t > S(0) ? S(1) : S(-1);
  S nn = S(1) / std::sqrt(t * t + S(1));
  j.s = -sign_t * (y / std::fabs(y)) * std::fabs(t) * nn;
  j.c = nn;
  return true;
}

template <typename S>
static void applyLeft(S* W, int p, int q, const Rot<S>& j) {  // rows p,q of 3x3 row-major W
  for (int i = 0; i < 3; ++i) {
    S xi = W[p * 3 + i], yi = W[q * 3 + i];
    W[p * 3 + i] = j.c * xi + j.s * yi;
    W[q * 3 + i] = -j.s * xi + j.c * yi;
  }
}

template <typename S>
static void applyRight(S* W, int p, int q, const Rot<S>& j) {  // cols p,q, rotation j (uses j^T form)
  for (int i = 0; i < 3; ++i) {
    S xi = W[i * 3 + p], yi = W[i * 3 + q];
    W[i * 3 + p] = j.c * xi - j.s * yi;
    W[i * 3 + q] = j.s * xi + j.c * yi;
  }
}

// A (row-major 3x3) = U * diag(sv) * V^T, sv descending, U and V orthogonal.
template <typename S>
static void jacobi_svd3(const S* A, S* U, S* sv, S* V) {
  const S precision = S(2) * std::numeric_limits<S>::epsilon();
  const S considerAsZero = std::numeric_limits<S>::min();
  S scale = 0;
  for (int i = 0; i < 9; ++i) scale = std::max(scale, std::fabs(A[i]));
  if (scale == S(0)) scale = S(1);
  S W[9];
  for (int i = 0; i < 9; ++i) W[i] = A[i] / scale;
  for (int i = 0; i < 9; ++i) U[i] = V[i] = (i % 4 == 0) ? S(1) : S(0);
  S maxDiag = std::max(std::fabs(W[0]), std::max(std::fabs(W[4]), std::fabs(W[8])));
  bool finished = false;
  int sweeps = 0;
  while (!finished && sweeps++ < 64) {
    finished = true;
    for (int p = 1; p < 3; ++p)
      for (int q = 0; q < p; ++q) {
        S threshold = std::max(considerAsZero, precision * maxDiag);
        if (std::fabs(W[p * 3 + q]) > threshold || std::fabs(W[q * 3 + p]) > threshold) {
          finished = false;
          // real_2x2_jacobi_svd on the (p,q) block
          S m00 = W[p * 3 + p], m01 = W[p * 3 + q], m10 = W[q * 3 + p], m11 = W[q * 3 + q];
          Rot<S> rot1;
          S t = m00 + m11;
          S d = m10 - m01;
          if (std::fabs(d) < std::numeric_limits<S>::min()) {
            rot1.s = S(0);
            rot1.c = S(1);
          } else {
            S u = t / d;
            S tmp = std::sqrt(S(1) + u * u);
            rot1.s = S(1) / tmp;
            rot1.c = u / tmp;
          }
          // m.applyOnTheLeft(0,1,rot1)
          S n00 = rot1.c * m00 + rot1.s * m10, n01 = rot1.c * m01 + rot1.s * m11;
          S n11 = -rot1.s * m01 + rot1.c * m11;
          Rot<S> jr;
          makeJacobi(n00, n01, n11, jr);
          // j_left = rot1 * j_right^T   (composition of plane rotations)
          Rot<S> jl;
          jl.c = rot1.c * jr.c + rot1.s * jr.s;
          jl.s = rot1.s * jr.c - rot1.c * jr.s;
          applyLeft(W, p, q, jl);
          Rot<S> jlt{jl.c, -jl.s};
          applyRight(U, p, q, jlt);  // U.applyOnTheRight(p,q,j_left.transpose())
          applyRight(W, p, q, jr);  // W.applyOnTheRight(p,q,j_right)
          applyRight(V, p, q, jr);
          maxDiag = std::max(maxDiag, std::max(std::fabs(W[p * 3 + p]), std::fabs(W[q * 3 + q])));
        }
      }
  }
  for (int i = 0; i < 3; ++i) {
    S a = W[i * 3 + i];
    sv[i] = std::fabs(a) * scale;
    if (a < S(0))
      for (int r = 0; r < 3; ++r) U[r * 3 + i] = -U[r * 3 + i];
  }
  for (int i = 0; i < 3; ++i) {  // selection sort, descending, swapping columns
    int pos = i;
    for (int k = i + 1; k < 3; ++k)
      if (sv[k] > sv[pos]) pos = k;
    if (sv[pos] == S(0)) break;
    if (pos != i) {
      std::swap(sv[i], sv[pos]);
      for (int r = 0; r < 3; ++r) {
        std::swap(U[r * 3 + i], U[r * 3 + pos]);
        std::swap(V[r * 3 + i], V[r * 3 + pos]);
      }
    }
  }
}

template <typename S>
static S det3(const S* a) {
  return a[0] * (a[4] * a[8] - a[5] * a[7]) - a[1] * (a[3] * a[8] - a[5] * a[6]) + a[2] * (a[3] * a[7] - a[4] * a[6]);
}

// ------------------------------------------------------------------------------------------
// [PCL] common/include/pcl/common/impl/eigen.hpp : pcl::umeyama(src, dst, with_scaling=false)
// called by [PCL] registration/.../impl/transformation_estimation_svd.hpp (use_umeyama_ = true)
// on 3 x n matrices of the matched pairs, Scalar S (float in PCL).
// ------------------------------------------------------------------------------------------
template <typename S>
static M4 umeyama(const std::vector<S>& src, const std::vector<S>& dst, size_t n) {
  const S one_over_n = S(1) / static_cast<S>(n);
  S sm[3] = {0, 0, 0}, dm[3] = {0, 0, 0};
  for (size_t i = 0; i < n; ++i)
    for (int d = 0; d < 3; ++d) {
      sm[d] += src[3 * i + d];
      dm[d] += dst[3 * i + d];
    }
  for (int d = 0; d < 3; ++d) {
    sm[d] *= one_over_n;
    dm[d] *= one_over_n;
  }
  // sigma = one_over_n * dst_demean * src_demean^T
  S sigma[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (size_t i = 0; i < n; ++i) {
    S sd[3], dd[3];
    for (int d = 0; d < 3; ++d) {
      sd[d] = src[3 * i + d] - sm[d];
      dd[d] = dst[3 * i + d] - dm[d];
    }
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) sigma[r * 3 + c] += dd[r] * sd[c];
  }
  for (int i = 0; i < 9; ++i) sigma[i] *= one_over_n;
  S U[9], V[9], sv[3];
  jacobi_svd3<S>(sigma, U, sv, V);
  S Sd[3] = {1, 1, 1};
  if (det3<S>(sigma) < S(0)) Sd[2] = -1;
  int rank = 0;
  for (int i = 0; i < 3; ++i) {
    // !isMuchSmallerThan(d_i, d_0): |d_i| > |d_0| * dummy_precision (1e-5 float / 1e-12 double)
    const S prec = sizeof(S) == 4 ? S(1e-5) : S(1e-12);
    if (!(std::fabs(sv[i]) <= std::fabs(sv[0]) * prec)) ++rank;
  }
  S R[9];
  auto usvt = [&](const S* sdiag) {
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) {
        S acc = 0;
        for (int k = 0; k < 3; ++k) acc += U[r * 3 + k] * sdiag[k] * V[c * 3 + k];
        R[r * 3 + c] = acc;
      }
  };
  if (rank == 2) {
    if (det3<S>(U) * det3<S>(V) > S(0)) {
      S ones[3] = {1, 1, 1};
      usvt(ones);
    } else {
      S s2[3] = {1, 1, -1};
      usvt(s2);
    }
  } else {
    usvt(Sd);
  }
  M4 T = M4::identity();
  for (int r = 0; r < 3; ++r) {
    S acc = 0;
    for (int c = 0; c < 3; ++c) {
      T(r, c) = static_cast<float>(R[r * 3 + c]);
      acc += R[r * 3 + c] * sm[c];
    }
    T(r, 3) = static_cast<float>(dm[r] - acc);
  }
  return T;
}

// ------------------------------------------------------------------------------------------
// [EIGEN] Matrix<double,6,6>::inverse() -> PartialPivLU, then inverse * b.
// ------------------------------------------------------------------------------------------
static bool inverse6(const double* A, double* inv) {
  double lu[36];
  std::memcpy(lu, A, sizeof(lu));
  int perm[6];
  for (int i = 0; i < 6; ++i) perm[i] = i;
  for (int k = 0; k < 6; ++k) {
    int piv = k;
    double best = std::fabs(lu[k * 6 + k]);
    for (int r = k + 1; r < 6; ++r)
      if (std::fabs(lu[r * 6 + k]) > best) {
        best = std::fabs(lu[r * 6 + k]);
        piv = r;
      }
    if (piv != k) {
      for (int c = 0; c < 6; ++c) std::swap(lu[k * 6 + c], lu[piv * 6 + c]);
      std::swap(perm[k], perm[piv]);
    }
    if (lu[k * 6 + k] == 0.0) continue;  // singular: Eigen carries on producing inf/nan
    for (int r = k + 1; r < 6; ++r) {
      lu[r * 6 + k] /= lu[k * 6 + k];
      for (int c = k + 1; c < 6; ++c) lu[r * 6 + c] -= lu[r * 6 + k] * lu[k * 6 + c];
    }
  }
  for (int col = 0; col < 6; ++col) {
    double y[6];
    for (int r = 0; r < 6; ++r) {
      double v = (perm[r] == col) ? 1.0 : 0.0;
      for (int c = 0; c < r; ++c) v -= lu[r * 6 + c] * y[c];
      y[r] = v;
    }
    for (int r = 5; r >= 0; --r) {
      double v = y[r];
      for (int c = r + 1; c < 6; ++c) v -= lu[r * 6 + c] * inv[c * 6 + col];
      inv[r * 6 + col] = v / lu[r * 6 + r];
    }
  }
  return true;
}

// [PCL] registration/.../impl/transformation_estimation_point_to_plane_lls.hpp : constructTransformationMatrix
static M4 construct_rt(double alpha, double beta, double gamma, double tx, double ty, double tz) {
  M4 T;
  for (int i = 0; i < 16; ++i) T.m[i] = 0;
  T(0, 0) = static_cast<float>(cos(gamma) * cos(beta));
  T(0, 1) = static_cast<float>(-sin(gamma) * cos(alpha) + cos(gamma) * sin(beta) * sin(alpha));
  T(0, 2) = static_cast<float>(sin(gamma) * sin(alpha) + cos(gamma) * sin(beta) * cos(alpha));
  T(1, 0) = static_cast<float>(sin(gamma) * cos(beta));
  T(1, 1) = static_cast<float>(cos(gamma) * cos(alpha) + sin(gamma) * sin(beta) * sin(alpha));
  T(1, 2) = static_cast<float>(-cos(gamma) * sin(alpha) + sin(gamma) * sin(beta) * cos(alpha));
  T(2, 0) = static_cast<float>(-sin(beta));
  T(2, 1) = static_cast<float>(cos(beta) * sin(alpha));
  T(2, 2) = static_cast<float>(cos(beta) * cos(alpha));
  T(0, 3) = static_cast<float>(tx);
  T(1, 3) = static_cast<float>(ty);
  T(2, 3) = static_cast<float>(tz);
  T(3, 3) = 1.0f;
  return T;
}

// [PCL] registration/.../impl/transformation_estimation_point_to_plane_lls.hpp : estimateRigidTransformation
struct Corr {
  int q, m;
  float d2;
};
static M4 point_to_plane_lls(const std::vector<float>& work, const float* tgt, size_t tstride_f,
                             const float* tn, size_t nstride_f, const std::vector<Corr>& corrs) {
  double ATA[36], ATb[6];
  for (double& v : ATA) v = 0;
  for (double& v : ATb) v = 0;
  for (const Corr& c : corrs) {
    const float* s = &work[4 * static_cast<size_t>(c.q)];
    const float* d = tgt + static_cast<size_t>(c.m) * tstride_f;
    const float* nr = tn + static_cast<size_t>(c.m) * nstride_f;
    if (!finite3(s[0], s[1], s[2]) || !finite3(d[0], d[1], d[2]) || !finite3(nr[0], nr[1], nr[2])) continue;
    const float sx = s[0], sy = s[1], sz = s[2];
    const float dx = d[0], dy = d[1], dz = d[2];
    const float nx = nr[0], ny = nr[1], nz = nr[2];
    double a = nz * sy - ny * sz;  // float expression, widened
    double b = nx * sz - nz * sx;
    double cc = ny * sx - nx * sy;
    ATA[0] += a * a;
    ATA[1] += a * b;
    ATA[2] += a * cc;
    ATA[3] += a * nx;
    ATA[4] += a * ny;
    ATA[5] += a * nz;
    ATA[7] += b * b;
    ATA[8] += b * cc;
    ATA[9] += b * nx;
    ATA[10] += b * ny;
    ATA[11] += b * nz;
    ATA[14] += cc * cc;
    ATA[15] += cc * nx;
    ATA[16] += cc * ny;
    ATA[17] += cc * nz;
    ATA[21] += nx * nx;  // float products
    ATA[22] += nx * ny;
    ATA[23] += nx * nz;
    ATA[28] += ny * ny;
    ATA[29] += ny * nz;
    ATA[35] += nz * nz;
    double dd = nx * dx + ny * dy + nz * dz - nx * sx - ny * sy - nz * sz;  // float expression
    ATb[0] += a * dd;
    ATb[1] += b * dd;
    ATb[2] += cc * dd;
    ATb[3] += nx * dd;
    ATb[4] += ny * dd;
    ATb[5] += nz * dd;
  }
  for (int r = 1; r < 6; ++r)
    for (int c = 0; c < r; ++c) ATA[r * 6 + c] = ATA[c * 6 + r];
  double inv[36];
  inverse6(ATA, inv);
  double x[6];
  for (int r = 0; r < 6; ++r) {
    double acc = 0;
    for (int c = 0; c < 6; ++c) acc += inv[r * 6 + c] * ATb[c];
    x[r] = acc;
  }
  return construct_rt(x[0], x[1], x[2], x[3], x[4], x[5]);
}

// ------------------------------------------------------------------------------------------
// [PCL] registration/.../default_convergence_criteria.h + impl/default_convergence_criteria.hpp
// ------------------------------------------------------------------------------------------
struct Criteria {
  int max_iterations = 1000;
  bool failure_after_max_iter = false;
  double rotation_threshold = 0.99999;
  double translation_threshold = 3e-4 * 3e-4;
  double mse_rel = 0.00001;
  double mse_abs = 1e-12;
  int max_similar = 0;
  int similar = 0;
  double prev_mse = std::numeric_limits<double>::max();
  double cur_mse = std::numeric_limits<double>::max();
  int state = PEB_NOT_CONVERGED;

  bool hasConverged(int iterations, const M4& tr, const std::vector<Corr>& corrs) {
    if (state != PEB_NOT_CONVERGED) {
      similar = 0;
      state = PEB_NOT_CONVERGED;
    }
    bool is_similar = false;
    if (iterations >= max_iterations) {
      if (!failure_after_max_iter) {
        state = PEB_ITERATIONS;
        return true;
      }
      state = PEB_FAILURE_AFTER_MAX_ITERATIONS;
    }
    double cos_angle = 0.5 * (tr(0, 0) + tr(1, 1) + tr(2, 2) - 1);               // float sum, then * 0.5 in double
    double translation_sqr = tr(0, 3) * tr(0, 3) + tr(1, 3) * tr(1, 3) + tr(2, 3) * tr(2, 3);  // float expression
    if (cos_angle >= rotation_threshold && translation_sqr <= translation_threshold) {
      if (similar >= max_similar) {
        state = PEB_TRANSFORM;
        return true;
      }
      is_similar = true;
    }
    double mse = 0;
    for (const Corr& c : corrs) mse += c.d2;
    mse /= static_cast<double>(corrs.size());
    cur_mse = mse;
    if (std::fabs(cur_mse - prev_mse) < mse_abs) {
      if (similar >= max_similar) {
        state = PEB_ABS_MSE;
        return true;
      }
      is_similar = true;
    }
    if (std::fabs(cur_mse - prev_mse) / prev_mse < mse_rel) {
      if (similar >= max_similar) {
        state = PEB_REL_MSE;
        return true;
      }
      is_similar = true;
    }
    if (is_similar)
      ++similar;
    else
      similar = 0;
    prev_mse = cur_mse;
    return false;
  }
};

// ------------------------------------------------------------------------------------------
// ICP.  [PCL] registration/.../impl/icp.hpp (computeTransformation, transformCloud),
// impl/registration.hpp (align, getFitnessScore), impl/correspondence_estimation.hpp
// (determineCorrespondences), registration/src/correspondence_rejection_distance.cpp.
// ------------------------------------------------------------------------------------------
struct IcpTrace {
  std::vector<M4> increments;
  std::vector<double> mse;
  std::vector<int> ncorr;
};

struct Icp {
  KdTree tree;
  const void* tgt = nullptr;
  size_t tgt_n = 0, tgt_stride = 0;
  const void* tgt_normals = nullptr;
  size_t tgt_nstride = 0;
  bool wide_accum = false;  // diagnostic only: run umeyama in double (quantifies PCL's float noise)

  void setTarget(const void* pts, size_t n, size_t stride, const void* normals, size_t nstride) {
    tgt = pts;
    tgt_n = n;
    tgt_stride = stride;
    tgt_normals = normals;
    tgt_nstride = nstride;
    tree.build(pts, n, stride);
  }

  double fitness(const void* src, size_t n, size_t stride, const M4& T, double max_range, int* n_in) const {
    double score = 0;
    int nr = 0;
    for (size_t i = 0; i < n; ++i) {
      const float* p = rec(src, i, stride);
      float q[3];
      if (!finite3(p[0], p[1], p[2])) continue;  // KdTreeFLANN asserts on invalid queries; skipped here
      transform_tpc(T, p, q);
      int id;
      float d2;
      if (tree.knn(q, 1, &id, &d2) == 0) continue;
      if (d2 <= max_range) {
        score += d2;
        ++nr;
      }
    }
    if (n_in) *n_in = nr;
    return nr > 0 ? score / nr : std::numeric_limits<double>::max();
  }

  void align(const void* src, size_t n, size_t stride, const M4& guess, const peb_icp_params& prm,
             peb_icp_result* res, float* out_aligned, int32_t* out_idx, float* out_d2, IcpTrace* trace) const {
    // working cloud (input_transformed), xyz1 records
    std::vector<float> work(4 * n);
    std::vector<char> valid(n);
    bool guess_is_identity = true;
    {
      M4 I = M4::identity();
      for (int i = 0; i < 16; ++i)
        if (guess.m[i] != I.m[i]) guess_is_identity = false;
    }
    for (size_t i = 0; i < n; ++i) {
      const float* p = rec(src, i, stride);
      valid[i] = finite3(p[0], p[1], p[2]);
      work[4 * i + 0] = p[0];
      work[4 * i + 1] = p[1];
      work[4 * i + 2] = p[2];
      work[4 * i + 3] = 1.0f;
      if (!guess_is_identity && valid[i]) transform_icp(guess, p, &work[4 * i]);
    }
    M4 final_t = guess;
    M4 transformation = M4::identity();
    int nr_iterations = 0;
    bool converged = false;

    Criteria crit;
    crit.max_similar = prm.max_iterations_similar;
    crit.mse_abs = prm.abs_mse_threshold;
    crit.max_iterations = prm.max_iterations;
    crit.mse_rel = prm.euclidean_fitness_epsilon;
    crit.translation_threshold = prm.transformation_epsilon;
    crit.rotation_threshold = prm.rotation_epsilon > 0 ? prm.rotation_epsilon : 1.0 - prm.transformation_epsilon;

    const double max_dist_sqr = prm.max_corr_dist * prm.max_corr_dist;
    const bool use_rejector = prm.rejector_max_dist > 0;
    const float rej_max2 = static_cast<float>(prm.rejector_max_dist * prm.rejector_max_dist);
    const float* tn = static_cast<const float*>(tgt_normals);
    std::vector<Corr> corrs;
    corrs.reserve(n);
    std::vector<float> ms, md;
    if (out_idx)
      for (size_t i = 0; i < n; ++i) out_idx[i] = -1;
    if (out_d2)
      for (size_t i = 0; i < n; ++i) out_d2[i] = 0.0f;

    do {
      corrs.clear();
      for (size_t i = 0; i < n; ++i) {
        if (!valid[i]) continue;
        int id;
        float d2;
        if (tree.knn(&work[4 * i], 1, &id, &d2) == 0) continue;
        if (d2 > max_dist_sqr) continue;                 // determineCorrespondences: strict >
        if (use_rejector && !(d2 < rej_max2)) continue;  // CorrespondenceRejectorDistance: keep iff <
        corrs.push_back({static_cast<int>(i), id, d2});
      }
      if (out_idx || out_d2) {
        if (out_idx)
          for (size_t i = 0; i < n; ++i) out_idx[i] = -1;
        if (out_d2)
          for (size_t i = 0; i < n; ++i) out_d2[i] = 0.0f;
        for (const Corr& c : corrs) {
          if (out_idx) out_idx[c.q] = c.m;
          if (out_d2) out_d2[c.q] = c.d2;
        }
      }
      if (static_cast<int>(corrs.size()) < prm.min_correspondences) {
        crit.state = PEB_NO_CORRESPONDENCES;
        converged = false;
        break;
      }
      if (prm.estimator == PEB_ESTIMATOR_SVD) {
        size_t nc = corrs.size();
        if (wide_accum) {
          std::vector<double> s(3 * nc), d(3 * nc);
          for (size_t k = 0; k < nc; ++k) {
            const float* t = rec(tgt, corrs[k].m, tgt_stride);
            for (int c = 0; c < 3; ++c) {
              s[3 * k + c] = work[4 * static_cast<size_t>(corrs[k].q) + c];
              d[3 * k + c] = t[c];
            }
          }
          transformation = umeyama<double>(s, d, nc);
        } else {
          ms.resize(3 * nc);
          md.resize(3 * nc);
          for (size_t k = 0; k < nc; ++k) {
            const float* t = rec(tgt, corrs[k].m, tgt_stride);
            for (int c = 0; c < 3; ++c) {
              ms[3 * k + c] = work[4 * static_cast<size_t>(corrs[k].q) + c];
              md[3 * k + c] = t[c];
            }
          }
          transformation = umeyama<float>(ms, md, nc);
        }
      } else {
        transformation = point_to_plane_lls(work, static_cast<const float*>(tgt), tgt_stride / 4, tn,
                                            tgt_nstride / 4, corrs);
      }
      for (size_t i = 0; i < n; ++i) {
        if (!valid[i]) continue;
        float o[3];
        transform_icp(transformation, &work[4 * i], o);
        work[4 * i] = o[0];
        work[4 * i + 1] = o[1];
        work[4 * i + 2] = o[2];
      }
      final_t = mul(transformation, final_t);
      ++nr_iterations;
      if (trace) {
        trace->increments.push_back(transformation);
        trace->ncorr.push_back(static_cast<int>(corrs.size()));
      }
      converged = crit.hasConverged(nr_iterations, transformation, corrs);
      if (trace) trace->mse.push_back(crit.cur_mse);
    } while (crit.state == PEB_NOT_CONVERGED);

    if (out_aligned) {
      for (size_t i = 0; i < n; ++i) {
        const float* p = rec(src, i, stride);
        float* o = out_aligned + 4 * i;
        o[0] = p[0];
        o[1] = p[1];
        o[2] = p[2];
        o[3] = 1.0f;
        if (finite3(p[0], p[1], p[2])) transform_icp(final_t, p, o);
      }
    }
    std::memcpy(res->T, final_t.m, sizeof(res->T));
    res->iterations = nr_iterations;
    res->converged = converged ? 1 : 0;
    res->state = crit.state;
    res->last_mse = crit.cur_mse;
    res->n_correspondences = static_cast<int32_t>(corrs.size());
    res->fitness = fitness(src, n, stride, final_t, prm.fitness_max_range, nullptr);
  }
};

// ------------------------------------------------------------------------------------------
// [PCL] filters/include/pcl/filters/impl/voxel_grid.hpp : VoxelGrid<PointXYZ>::applyFilter
//   (downsample_all_data_ = true -> CentroidPoint<PointXYZ>, [PCL] common/.../impl/centroid.hpp;
//    getMinMax3D, [PCL] common/.../impl/common.hpp).
// PCL's std::sort is unstable on equal voxel ids, so the within-voxel float summation order is
// implementation-defined upstream; the oracle fixes it to ascending original index.
// returns: number of output points; *unchanged = 1 if the overflow guard returned the input.
// ------------------------------------------------------------------------------------------
static size_t voxel_grid(const void* pts, size_t n, size_t stride, const float leaf[3], unsigned min_pts,
                         float* out, int* unchanged) {
  *unchanged = 0;
  float inv[3] = {1.0f / leaf[0], 1.0f / leaf[1], 1.0f / leaf[2]};
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (size_t i = 0; i < n; ++i) {
    const float* p = rec(pts, i, stride);
    if (!finite3(p[0], p[1], p[2])) continue;
    for (int d = 0; d < 3; ++d) {
      mn[d] = std::min(mn[d], p[d]);
      mx[d] = std::max(mx[d], p[d]);
    }
  }
  int64_t dx = static_cast<int64_t>((mx[0] - mn[0]) * inv[0]) + 1;
  int64_t dy = static_cast<int64_t>((mx[1] - mn[1]) * inv[1]) + 1;
  int64_t dz = static_cast<int64_t>((mx[2] - mn[2]) * inv[2]) + 1;
  if (dx * dy * dz > static_cast<int64_t>(std::numeric_limits<int32_t>::max())) {
    for (size_t i = 0; i < n; ++i) {
      const float* p = rec(pts, i, stride);
      out[4 * i] = p[0];
      out[4 * i + 1] = p[1];
      out[4 * i + 2] = p[2];
      out[4 * i + 3] = 1.0f;
    }
    *unchanged = 1;
    return n;
  }
  int min_b[3], max_b[3], div_b[3];
  for (int d = 0; d < 3; ++d) {
    min_b[d] = static_cast<int>(std::floor(mn[d] * inv[d]));
    max_b[d] = static_cast<int>(std::floor(mx[d] * inv[d]));
    div_b[d] = max_b[d] - min_b[d] + 1;
  }
  int mul[3] = {1, div_b[0], div_b[0] * div_b[1]};
  struct IdxPt {
    unsigned idx;
    unsigned pt;
  };
  std::vector<IdxPt> iv;
  iv.reserve(n);
  for (size_t i = 0; i < n; ++i) {
    const float* p = rec(pts, i, stride);
    if (!finite3(p[0], p[1], p[2])) continue;
    int ijk0 = static_cast<int>(std::floor(p[0] * inv[0]) - static_cast<float>(min_b[0]));
    int ijk1 = static_cast<int>(std::floor(p[1] * inv[1]) - static_cast<float>(min_b[1]));
    int ijk2 = static_cast<int>(std::floor(p[2] * inv[2]) - static_cast<float>(min_b[2]));
    int idx = ijk0 * mul[0] + ijk1 * mul[1] + ijk2 * mul[2];
    iv.push_back({static_cast<unsigned>(idx), static_cast<unsigned>(i)});
  }
  std::stable_sort(iv.begin(), iv.end(), [](const IdxPt& a, const IdxPt& b) { return a.idx < b.idx; });
  size_t total = 0, index = 0;
  while (index < iv.size()) {
    size_t i = index + 1;
    while (i < iv.size() && iv[i].idx == iv[index].idx) ++i;
    if (i - index >= min_pts) {
      float sx = 0, sy = 0, sz = 0;
      for (size_t li = index; li < i; ++li) {
        const float* p = rec(pts, iv[li].pt, stride);
        sx += p[0];
        sy += p[1];
        sz += p[2];
      }
      float cnt = static_cast<float>(i - index);
      out[4 * total] = sx / cnt;
      out[4 * total + 1] = sy / cnt;
      out[4 * total + 2] = sz / cnt;
      out[4 * total + 3] = 1.0f;
      ++total;
    }
    index = i;
  }
  return total;
}

// ------------------------------------------------------------------------------------------
// [PCL] common/include/pcl/common/impl/eigen.hpp : computeRoots2, computeRoots, eigen33
// (smallest eigenvalue + eigenvector), Scalar = float.
// ------------------------------------------------------------------------------------------
static void computeRoots2(float b, float c, float* roots) {
  roots[0] = 0.0f;
  float d = static_cast<float>(b * b - 4.0 * c);
  if (d < 0.0) d = 0.0;
  float sd = std::sqrt(d);
  roots[2] = 0.5f * (b + sd);
  roots[1] = 0.5f * (b - sd);
}

static void computeRoots(const float* m /*row-major 3x3 symmetric*/, float* roots) {
  float c0 = m[0] * m[4] * m[8] + 2.0f * m[1] * m[2] * m[5] - m[0] * m[5] * m[5] - m[4] * m[2] * m[2] -
             m[8] * m[1] * m[1];
  float c1 = m[0] * m[4] - m[1] * m[1] + m[0] * m[8] - m[2] * m[2] + m[4] * m[8] - m[5] * m[5];
  float c2 = m[0] + m[4] + m[8];
  if (std::fabs(c0) < std::numeric_limits<float>::epsilon()) {
    computeRoots2(c2, c1, roots);
  } else {
    const float s_inv3 = static_cast<float>(1.0 / 3.0);
    const float s_sqrt3 = std::sqrt(3.0f);
    float c2_over_3 = c2 * s_inv3;
    float a_over_3 = (c1 - c2 * c2_over_3) * s_inv3;
    if (a_over_3 > 0.0f) a_over_3 = 0.0f;
    float half_b = 0.5f * (c0 + c2_over_3 * (2.0f * c2_over_3 * c2_over_3 - c1));
    float q = half_b * half_b + a_over_3 * a_over_3 * a_over_3;
    if (q > 0.0f) q = 0.0f;
    float rho = std::sqrt(-a_over_3);
    float theta = std::atan2(std::sqrt(-q), half_b) * s_inv3;
    float cos_theta = std::cos(theta);
    float sin_theta = std::sin(theta);
    roots[0] = c2_over_3 + 2.0f * rho * cos_theta;
    roots[1] = c2_over_3 - rho * (cos_theta + s_sqrt3 * sin_theta);
    roots[2] = c2_over_3 - rho * (cos_theta - s_sqrt3 * sin_theta);
    if (roots[0] >= roots[1]) std::swap(roots[0], roots[1]);
    if (roots[1] >= roots[2]) {
      std::swap(roots[1], roots[2]);
      if (roots[0] >= roots[1]) std::swap(roots[0], roots[1]);
    }
    if (roots[0] <= 0) computeRoots2(c2, c1, roots);
  }
}

static void eigen33(const float* mat, float& eigenvalue, float* ev) {
  float scale = 0;
  for (int i = 0; i < 9; ++i) scale = std::max(scale, std::fabs(mat[i]));
  if (scale <= std::numeric_limits<float>::min()) scale = 1.0f;
  float s[9];
  for (int i = 0; i < 9; ++i) s[i] = mat[i] / scale;
  float roots[3];
  computeRoots(s, roots);
  eigenvalue = roots[0] * scale;
  s[0] -= roots[0];
  s[4] -= roots[0];
  s[8] -= roots[0];
  auto cross = [](const float* a, const float* b, float* o) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
  };
  float v1[3], v2[3], v3[3];
  cross(&s[0], &s[3], v1);
  cross(&s[0], &s[6], v2);
  cross(&s[3], &s[6], v3);
  float l1 = v1[0] * v1[0] + v1[1] * v1[1] + v1[2] * v1[2];
  float l2 = v2[0] * v2[0] + v2[1] * v2[1] + v2[2] * v2[2];
  float l3 = v3[0] * v3[0] + v3[1] * v3[1] + v3[2] * v3[2];
  const float* v;
  float l;
  if (l1 >= l2 && l1 >= l3) {
    v = v1;
    l = l1;
  } else if (l2 >= l1 && l2 >= l3) {
    v = v2;
    l = l2;
  } else {
    v = v3;
    l = l3;
  }
  float sl = std::sqrt(l);
  ev[0] = v[0] / sl;
  ev[1] = v[1] / sl;
  ev[2] = v[2] / sl;
}

// [PCL] features/.../impl/normal_3d.hpp (computeFeature), features/.../normal_3d.h
// (computePointNormal, flipNormalTowardsViewpoint), features/.../impl/feature.hpp
// (solvePlaneParameters), common/.../impl/centroid.hpp (computeMeanAndCovarianceMatrix, 1.10:
// single pass, float, no origin shift).
static void normal_from_neighbours(const void* pts, size_t stride, const int* nn, int cnt, const float* p,
                                   const float vp[3], float* out8) {
  const float qnan = std::numeric_limits<float>::quiet_NaN();
  for (int i = 0; i < 8; ++i) out8[i] = 0.0f;
  if (cnt < 3) {
    out8[0] = out8[1] = out8[2] = out8[4] = qnan;
    return;
  }
  float accu[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int j = 0; j < cnt; ++j) {
    const float* c = rec(pts, nn[j], stride);
    accu[0] += c[0] * c[0];
    accu[1] += c[0] * c[1];
    accu[2] += c[0] * c[2];
    accu[3] += c[1] * c[1];
    accu[4] += c[1] * c[2];
    accu[5] += c[2] * c[2];
    accu[6] += c[0];
    accu[7] += c[1];
    accu[8] += c[2];
  }
  float fc = static_cast<float>(cnt);
  for (int i = 0; i < 9; ++i) accu[i] /= fc;
  float cov[9];
  cov[0] = accu[0] - accu[6] * accu[6];
  cov[1] = accu[1] - accu[6] * accu[7];
  cov[2] = accu[2] - accu[6] * accu[8];
  cov[4] = accu[3] - accu[7] * accu[7];
  cov[5] = accu[4] - accu[7] * accu[8];
  cov[8] = accu[5] - accu[8] * accu[8];
  cov[3] = cov[1];
  cov[6] = cov[2];
  cov[7] = cov[5];
  float ev, vec[3];
  eigen33(cov, ev, vec);
  float eig_sum = cov[0] + cov[4] + cov[8];
  float curvature = (eig_sum != 0) ? std::fabs(ev / eig_sum) : 0.0f;
  float vx = vp[0] - p[0], vy = vp[1] - p[1], vz = vp[2] - p[2];
  float cos_theta = (vx * vec[0] + vy * vec[1] + vz * vec[2]);
  if (cos_theta < 0) {
    vec[0] *= -1;
    vec[1] *= -1;
    vec[2] *= -1;
  }
  out8[0] = vec[0];
  out8[1] = vec[1];
  out8[2] = vec[2];
  out8[4] = curvature;
}

static void normals_knn(const void* pts, size_t n, size_t stride, int k, const float vp[3], float* out8,
                        int32_t* out_nn /*nullable, n*k, -1 padded*/, int threads) {
  KdTree tree;
  tree.build(pts, n, stride);
  const float qnan = std::numeric_limits<float>::quiet_NaN();
#ifdef _OPENMP
  if (threads <= 0) threads = omp_get_max_threads();
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1024)
#endif
  for (int64_t i = 0; i < static_cast<int64_t>(n); ++i) {
    std::vector<int> idx(k);
    std::vector<float> d2(k);
    const float* p = rec(pts, i, stride);
    float* o = out8 + 8 * i;
    if (out_nn)
      for (int j = 0; j < k; ++j) out_nn[i * k + j] = -1;
    if (!finite3(p[0], p[1], p[2])) {
      for (int j = 0; j < 8; ++j) o[j] = 0.0f;
      o[0] = o[1] = o[2] = o[4] = qnan;
      continue;
    }
    int cnt = tree.knn(p, k, idx.data(), d2.data());
    if (out_nn)
      for (int j = 0; j < cnt; ++j) out_nn[i * k + j] = idx[j];
    normal_from_neighbours(pts, stride, idx.data(), cnt, p, vp, o);
  }
  (void)threads;
}

}  // namespace orc

// ==========================================================================================
// C entry points for ctypes (tests / bench cpu_baseline only)
// ==========================================================================================
extern "C" {

#define ORC_API __attribute__((visibility("default")))

ORC_API const char* orc_version(void) { return "pcl-1.10 restatement oracle r1 (parity unpinned: see header)"; }

ORC_API void* orc_kdtree_create(const void* pts, size_t n, size_t stride) {
  auto* t = new orc::KdTree();
  t->build(pts, n, stride);
  return t;
}
ORC_API void orc_kdtree_destroy(void* t) { delete static_cast<orc::KdTree*>(t); }

// out_idx/out_d2: nq*k, padded with -1 / inf when fewer than k points exist
ORC_API void orc_kdtree_knn(const void* tree, const void* q, size_t nq, size_t stride, int k, int32_t* out_idx,
                            float* out_d2) {
  const auto* t = static_cast<const orc::KdTree*>(tree);
  std::vector<int> idx(k);
  std::vector<float> d2(k);
  for (size_t i = 0; i < nq; ++i) {
    const float* p = orc::rec(q, i, stride);
    int cnt = orc::finite3(p[0], p[1], p[2]) ? t->knn(p, k, idx.data(), d2.data()) : 0;
    for (int j = 0; j < k; ++j) {
      out_idx[i * k + j] = j < cnt ? idx[j] : -1;
      out_d2[i * k + j] = j < cnt ? d2[j] : std::numeric_limits<float>::infinity();
    }
  }
}

// brute force, L2_Simple float order, lowest index wins ties
ORC_API void orc_nn_bruteforce(const void* tgt, size_t n, size_t tstride, const void* q, size_t nq, size_t qstride,
                               int32_t* out_idx, float* out_d2) {
  for (size_t i = 0; i < nq; ++i) {
    const float* p = orc::rec(q, i, qstride);
    float best = std::numeric_limits<float>::infinity();
    int bi = -1;
    for (size_t j = 0; j < n; ++j) {
      const float* t = orc::rec(tgt, j, tstride);
      if (!orc::finite3(t[0], t[1], t[2])) continue;
      float d = orc::l2_simple(p, t);
      if (d < best) {
        best = d;
        bi = static_cast<int>(j);
      }
    }
    out_idx[i] = bi;
    out_d2[i] = best;
  }
}

ORC_API size_t orc_voxel_grid(const void* pts, size_t n, size_t stride, float lx, float ly, float lz, unsigned min_pts,
                              float* out_xyz4, int* unchanged) {
  float leaf[3] = {lx, ly, lz};
  return orc::voxel_grid(pts, n, stride, leaf, min_pts, out_xyz4, unchanged);
}

ORC_API void orc_normals_knn(const void* pts, size_t n, size_t stride, int k, const float* vp, float* out8,
                             int32_t* out_nn, int threads) {
  orc::normals_knn(pts, n, stride, k, vp, out8, out_nn, threads);
}

ORC_API void* orc_icp_create(void) { return new orc::Icp(); }
ORC_API void orc_icp_destroy(void* h) { delete static_cast<orc::Icp*>(h); }
ORC_API void orc_icp_set_wide_accum(void* h, int on) { static_cast<orc::Icp*>(h)->wide_accum = on != 0; }
// the caller keeps pts / normals alive while the handle is used
ORC_API void orc_icp_set_target(void* h, const void* pts, size_t n, size_t stride, const void* normals,
                                size_t nstride) {
  static_cast<orc::Icp*>(h)->setTarget(pts, n, stride, normals, nstride);
}

// trace_T: nullable, cap*16 floats (column-major increments); trace_mse: cap doubles; trace_n -> count
ORC_API void orc_icp_align(const void* h, const void* src, size_t n, size_t stride, const float* guess,
                           const peb_icp_params* prm, peb_icp_result* res, float* out_aligned, int32_t* out_idx,
                           float* out_d2, float* trace_T, double* trace_mse, size_t cap, size_t* trace_n) {
  orc::M4 g = orc::M4::identity();
  if (guess) std::memcpy(g.m, guess, sizeof(g.m));
  orc::IcpTrace tr;
  static_cast<const orc::Icp*>(h)->align(src, n, stride, g, *prm, res, out_aligned, out_idx, out_d2,
                                         (trace_T || trace_mse) ? &tr : nullptr);
  size_t cnt = std::min(cap, tr.increments.size());
  for (size_t i = 0; i < cnt; ++i) {
    if (trace_T) std::memcpy(trace_T + 16 * i, tr.increments[i].m, 16 * sizeof(float));
    if (trace_mse) trace_mse[i] = tr.mse[i];
  }
  if (trace_n) *trace_n = cnt;
}

// H independent aligns (the registerModelToScene shape); OpenMP over hypotheses like OpenCV does
ORC_API void orc_icp_align_batch(const void* h, const void* src, size_t n, size_t stride, const float* guesses,
                                 size_t H, const peb_icp_params* prm, peb_icp_result* res, int threads) {
#ifdef _OPENMP
  if (threads <= 0) threads = omp_get_max_threads();
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1)
#endif
  for (int64_t i = 0; i < static_cast<int64_t>(H); ++i) {
    orc::M4 g;
    std::memcpy(g.m, guesses + 16 * i, sizeof(g.m));
    static_cast<const orc::Icp*>(h)->align(src, n, stride, g, *prm, &res[i], nullptr, nullptr, nullptr, nullptr);
  }
  (void)threads;
}

ORC_API double orc_icp_fitness(const void* h, const void* src, size_t n, size_t stride, const float* T,
                               double max_range, int32_t* n_inliers) {
  orc::M4 t;
  std::memcpy(t.m, T, sizeof(t.m));
  int ni = 0;
  double f = static_cast<const orc::Icp*>(h)->fitness(src, n, stride, t, max_range, &ni);
  if (n_inliers) *n_inliers = ni;
  return f;
}

// ---- small pieces exposed so the tests can pin them individually -------------------------
ORC_API void orc_umeyama(const float* src3, const float* dst3, size_t n, int use_double, float* out_T) {
  orc::M4 T;
  if (use_double) {
    std::vector<double> s(src3, src3 + 3 * n), d(dst3, dst3 + 3 * n);
    T = orc::umeyama<double>(s, d, n);
  } else {
    std::vector<float> s(src3, src3 + 3 * n), d(dst3, dst3 + 3 * n);
    T = orc::umeyama<float>(s, d, n);
  }
  std::memcpy(out_T, T.m, sizeof(T.m));
}

ORC_API void orc_svd3f(const float* A, float* U, float* S, float* V) { orc::jacobi_svd3<float>(A, U, S, V); }
ORC_API void orc_svd3d(const double* A, double* U, double* S, double* V) { orc::jacobi_svd3<double>(A, U, S, V); }
ORC_API void orc_eigen33(const float* cov, float* eigenvalue, float* eigenvector) {
  orc::eigen33(cov, *eigenvalue, eigenvector);
}
ORC_API void orc_inverse6(const double* A, double* inv) { orc::inverse6(A, inv); }
ORC_API void orc_transform_icp(const float* T, const float* p, float* o) {
  orc::M4 t;
  std::memcpy(t.m, T, sizeof(t.m));
  orc::transform_icp(t, p, o);
}
ORC_API void orc_transform_tpc(const float* T, const float* p, float* o) {
  orc::M4 t;
  std::memcpy(t.m, T, sizeof(t.m));
  orc::transform_tpc(t, p, o);
}
ORC_API void orc_mul4(const float* A, const float* B, float* C) {
  orc::M4 a, b;
  std::memcpy(a.m, A, sizeof(a.m));
  std::memcpy(b.m, B, sizeof(b.m));
  orc::M4 c = orc::mul(a, b);
  std::memcpy(C, c.m, sizeof(c.m));
}
// point-to-plane LLS step on explicit pairs (source xyz, target xyz, target normal), n pairs
ORC_API void orc_point_to_plane_lls(const float* s3, const float* d3, const float* n3, size_t n, float* out_T) {
  std::vector<float> work(4 * n);
  std::vector<orc::Corr> corrs(n);
  for (size_t i = 0; i < n; ++i) {
    work[4 * i] = s3[3 * i];
    work[4 * i + 1] = s3[3 * i + 1];
    work[4 * i + 2] = s3[3 * i + 2];
    work[4 * i + 3] = 1.0f;
    corrs[i] = {static_cast<int>(i), static_cast<int>(i), 0.0f};
  }
  orc::M4 T = orc::point_to_plane_lls(work, d3, 3, n3, 3, corrs);
  std::memcpy(out_T, T.m, sizeof(T.m));
}
// scripted convergence-state machine: feed (increment, mse) pairs, get the state sequence
ORC_API void orc_criteria_script(const peb_icp_params* prm, const float* incs, const double* mses, size_t n,
                                 int32_t* out_states, int32_t* out_ret) {
  orc::Criteria crit;
  crit.max_similar = prm->max_iterations_similar;
  crit.mse_abs = prm->abs_mse_threshold;
  crit.max_iterations = prm->max_iterations;
  crit.mse_rel = prm->euclidean_fitness_epsilon;
  crit.translation_threshold = prm->transformation_epsilon;
  crit.rotation_threshold = prm->rotation_epsilon > 0 ? prm->rotation_epsilon : 1.0 - prm->transformation_epsilon;
  for (size_t i = 0; i < n; ++i) {
    orc::M4 t;
    std::memcpy(t.m, incs + 16 * i, sizeof(t.m));
    std::vector<orc::Corr> c(1);
    c[0] = {0, 0, 0.0f};
    // a single pseudo-correspondence whose distance reproduces the scripted mse in float
    c[0].d2 = static_cast<float>(mses[i]);
    out_ret[i] = crit.hasConverged(static_cast<int>(i + 1), t, c) ? 1 : 0;
    out_states[i] = crit.state;
  }
}

// ------------------------------------------------------------------------------------------
// Scene pre-filter: restates the reference's OWN code (verified in /root/reference), in order:
//   pcl::removeNaNFromPointCloud          pose_estimation/src/pose_estimation.cpp:246-248
//   PoseEstimation::filter_points         pose_estimation/src/pose_estimation.cpp:347-372
//   the band test of remove_planes        pose_estimation/src/pose_estimation.cpp:309-333
// Survivors are kept in original order (the reference's order depends on OpenMP scheduling of its
// push_back under `omp critical`; ExtractIndices with setNegative(true) restores index order).
// ------------------------------------------------------------------------------------------
ORC_API size_t orc_scene_prefilter(const void* pts, size_t n, size_t stride, const peb_prefilter_params* f,
                                   float* out_xyz4) {
  size_t m = 0;
  for (size_t i = 0; i < n; ++i) {
    const float* p = orc::rec(pts, i, stride);
    const float x = p[0], y = p[1], z = p[2];
    if (!std::isfinite(x) || !std::isfinite(y) || !std::isfinite(z)) continue;
    if (f->use_sphere) {
      float dx = f->sphere_center[0] - x;
      float dy = f->sphere_center[1] - y;
      float dz = f->sphere_center[2] - z;
      float d = std::sqrt(dx * dx + dy * dy + dz * dz);
      bool inside = d <= f->sphere_radius;
      if (f->remove_inliers ? inside : !inside) continue;
    }
    bool near_plane = false;
    for (int k = 0; k < f->n_planes && !near_plane; ++k) {
      const float* c = f->planes + 4 * k;
      float d = (x * c[0] + y * c[1] + z * c[2] + c[3]) / std::sqrt(x * x + y * y + z * z);
      if (std::abs(d) <= f->plane_band) near_plane = true;
    }
    if (near_plane) continue;
    out_xyz4[4 * m] = x;
    out_xyz4[4 * m + 1] = y;
    out_xyz4[4 * m + 2] = z;
    out_xyz4[4 * m + 3] = 1.0f;
    ++m;
  }
  return m;
}

// ------------------------------------------------------------------------------------------
// pcl::SACSegmentation<PointXYZ>::segment with SACMODEL_PLANE / SAC_RANSAC, as PCL 1.10 runs it
// (the reference: pose_estimation/src/pose_estimation.cpp:285-297).
// [PCL] sample_consensus/impl/ransac.hpp : RandomSampleConsensus::computeModel
// [PCL] sample_consensus/sac_model.h : getSamples, drawIndexSample, rnd()
//       (boost::mt19937 seeded with 12345, boost::uniform_int<>(0, INT_MAX) == mt() >> 1)
// [PCL] sample_consensus/impl/sac_model_plane.hpp : isSampleGood, computeModelCoefficients,
//       countWithinDistance, selectWithinDistance, optimizeModelCoefficients
// [PCL] segmentation/impl/sac_segmentation.hpp : segment (optimise, then re-select the inliers)
// Written as the sequential loop PCL runs (the product draws all samples first and counts in one
// pass).  The 4-term dot product is summed left to right (Eigen's packet reduction order depends on
// the SSE level PCL was built with).  wide = 1: inlier moments in double (the product's choice);
// wide = 0: PCL 1.10's single-pass float sums in index order.
// ------------------------------------------------------------------------------------------
}  // extern "C"  (helpers below are C++)

namespace orc {

struct Mt19937 {  // boost::mt19937 == std::mt19937; restated so that the oracle pins the stream itself
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
        uint32_t y = (mt[i] & 0x80000000u) | (mt[(i + 1) % 624] & 0x7fffffffu);
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

static bool plane_sample_collinear(const float* p0, const float* p1, const float* p2) {
  // Eigen::Array4f dy1dy2 = (p1 - p0) / (p2 - p0); the 4th lanes (w = 1 - 1 = 0 -> 0/0) are not looked at
  float r[3];
  for (int i = 0; i < 3; ++i) r[i] = (p1[i] - p0[i]) / (p2[i] - p0[i]);
  return (r[0] == r[1]) && (r[2] == r[1]);
}

static float plane_dot(const float* mc, const float* p) { return ((mc[0] * p[0] + mc[1] * p[1]) + mc[2] * p[2]) + mc[3] * 1.0f; }

static bool plane_from_sample(const float* p0, const float* p1, const float* p2, float* mc) {
  if (plane_sample_collinear(p0, p1, p2)) return false;
  float a[3], b[3];
  for (int i = 0; i < 3; ++i) {
    a[i] = p1[i] - p0[i];
    b[i] = p2[i] - p0[i];
  }
  mc[0] = a[1] * b[2] - a[2] * b[1];
  mc[1] = a[2] * b[0] - a[0] * b[2];
  mc[2] = a[0] * b[1] - a[1] * b[0];
  mc[3] = 0.0f;
  // Eigen normalize(): v /= sqrt(squaredNorm)
  float nn = std::sqrt(((mc[0] * mc[0] + mc[1] * mc[1]) + mc[2] * mc[2]) + mc[3] * mc[3]);
  for (int i = 0; i < 4; ++i) mc[i] /= nn;
  mc[3] = -1.0f * (((mc[0] * p0[0] + mc[1] * p0[1]) + mc[2] * p0[2]) + mc[3] * 1.0f);
  return true;
}

}  // namespace orc

extern "C" {

// returns 1 if a model was found; coeff[4]; inliers (nullable) ascending indices; iterations run
ORC_API int orc_sac_plane(const void* pts, size_t n, size_t stride, const peb_sac_params* prm, int wide, float* coeff,
                          int32_t* inliers, size_t* n_inliers, int32_t* iterations_out) {
  using namespace orc;
  for (int i = 0; i < 4; ++i) coeff[i] = 0.0f;
  *n_inliers = 0;
  if (iterations_out) *iterations_out = 0;
  if (n < 3) return 0;
  std::vector<int> shuffled(n);
  for (size_t i = 0; i < n; ++i) shuffled[i] = static_cast<int>(i);
  Mt19937 rng(prm->seed);
  auto rnd = [&]() { return static_cast<int>(rng.next() >> 1); };  // uniform_int<>(0, INT_MAX) over a 32-bit engine
  const double threshold = prm->distance_threshold;
  int iterations = 0;
  int n_best = -std::numeric_limits<int>::max();
  double k = 1.0;
  const double log_probability = std::log(1.0 - prm->probability);
  const double one_over_indices = 1.0 / static_cast<double>(n);
  unsigned skipped = 0;
  const unsigned max_skip = static_cast<unsigned>(prm->max_iterations) * 10u;
  float best[4] = {0, 0, 0, 0};
  bool have = false;
  int sel[3];
  while (iterations < k && skipped < max_skip) {
    // getSamples: up to 1000 draws until isSampleGood
    bool good = false;
    for (unsigned it = 0; it < 1000 && !good; ++it) {
      for (size_t i = 0; i < 3; ++i) std::swap(shuffled[i], shuffled[i + (rnd() % (n - i))]);
      for (int i = 0; i < 3; ++i) sel[i] = shuffled[i];
      good = !plane_sample_collinear(rec(pts, sel[0], stride), rec(pts, sel[1], stride), rec(pts, sel[2], stride));
    }
    if (!good) break;  // "No samples could be selected!"
    float mc[4];
    if (!plane_from_sample(rec(pts, sel[0], stride), rec(pts, sel[1], stride), rec(pts, sel[2], stride), mc)) {
      ++skipped;
      continue;
    }
    int cnt = 0;
    for (size_t i = 0; i < n; ++i)
      if (static_cast<double>(std::abs(plane_dot(mc, rec(pts, i, stride)))) < threshold) ++cnt;
    if (cnt > n_best) {
      n_best = cnt;
      have = true;
      for (int i = 0; i < 4; ++i) best[i] = mc[i];
      double w = static_cast<double>(n_best) * one_over_indices;
      double p_no_outliers = 1.0 - std::pow(w, 3.0);
      p_no_outliers = std::max(std::numeric_limits<double>::epsilon(), p_no_outliers);
      p_no_outliers = std::min(1.0 - std::numeric_limits<double>::epsilon(), p_no_outliers);
      k = log_probability / std::log(p_no_outliers);
    }
    ++iterations;
    if (iterations > prm->max_iterations) break;
  }
  if (iterations_out) *iterations_out = iterations;
  if (!have) return 0;
  auto select = [&](const float* mc, std::vector<int>& out) {
    out.clear();
    for (size_t i = 0; i < n; ++i)
      if (static_cast<double>(std::abs(plane_dot(mc, rec(pts, i, stride)))) < threshold) out.push_back(static_cast<int>(i));
  };
  std::vector<int> inl;
  select(best, inl);
  float final_c[4] = {best[0], best[1], best[2], best[3]};
  if (prm->optimize_coefficients) {
    if (inl.size() > 3) {
      float cov[9], cen[3];
      if (wide) {
        double a[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (int j : inl) {
          const float* c = rec(pts, j, stride);
          const double x = c[0], y = c[1], z = c[2];
          a[0] += x * x, a[1] += x * y, a[2] += x * z, a[3] += y * y, a[4] += y * z, a[5] += z * z, a[6] += x, a[7] += y, a[8] += z;
        }
        for (int i = 0; i < 9; ++i) a[i] /= static_cast<double>(inl.size());
        cov[0] = static_cast<float>(a[0] - a[6] * a[6]);
        cov[1] = static_cast<float>(a[1] - a[6] * a[7]);
        cov[2] = static_cast<float>(a[2] - a[6] * a[8]);
        cov[4] = static_cast<float>(a[3] - a[7] * a[7]);
        cov[5] = static_cast<float>(a[4] - a[7] * a[8]);
        cov[8] = static_cast<float>(a[5] - a[8] * a[8]);
        for (int i = 0; i < 3; ++i) cen[i] = static_cast<float>(a[6 + i]);
      } else {
        float a[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (int j : inl) {
          const float* c = rec(pts, j, stride);
          a[0] += c[0] * c[0], a[1] += c[0] * c[1], a[2] += c[0] * c[2], a[3] += c[1] * c[1], a[4] += c[1] * c[2];
          a[5] += c[2] * c[2], a[6] += c[0], a[7] += c[1], a[8] += c[2];
        }
        const float fc = static_cast<float>(inl.size());
        for (int i = 0; i < 9; ++i) a[i] /= fc;
        cov[0] = a[0] - a[6] * a[6];
        cov[1] = a[1] - a[6] * a[7];
        cov[2] = a[2] - a[6] * a[8];
        cov[4] = a[3] - a[7] * a[7];
        cov[5] = a[4] - a[7] * a[8];
        cov[8] = a[5] - a[8] * a[8];
        for (int i = 0; i < 3; ++i) cen[i] = a[6 + i];
      }
      cov[3] = cov[1];
      cov[6] = cov[2];
      cov[7] = cov[5];
      float ev, vec[3];
      eigen33(cov, ev, vec);
      final_c[0] = vec[0];
      final_c[1] = vec[1];
      final_c[2] = vec[2];
      final_c[3] = 0.0f;
      final_c[3] = -1.0f * (((final_c[0] * cen[0] + final_c[1] * cen[1]) + final_c[2] * cen[2]) + final_c[3] * 1.0f);
    }
    select(final_c, inl);  // segment(): "Refine inliers"
  }
  for (int i = 0; i < 4; ++i) coeff[i] = final_c[i];
  *n_inliers = inl.size();
  if (inliers) std::copy(inl.begin(), inl.end(), inliers);
  return 1;
}

ORC_API uint32_t orc_mt19937_nth(uint32_t seed, uint32_t nth) {
  orc::Mt19937 r(seed);
  uint32_t v = 0;
  for (uint32_t i = 0; i < nth; ++i) v = r.next();
  return v;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------
// cv::ppf_match_3d::ICP::registerModelToScene — what the reference runs in the refinement slot
// (pose_estimation/src/opencv_surface_match.cpp:85-94).  [CV] opencv_contrib/modules/
// surface_matching/src/icp.cpp, ppf_helpers.cpp (transformPCPose, samplePCUniform), c_utils.hpp
// (eulerToDCM, rtToPose).  PARITY UNPINNED: the contrib sources are in neither /root/reference nor
// this image; this is the algorithm as recollected, restated sequentially in
// the arithmetic OpenCV uses (float clouds, double poses and solves).  Choices where OpenCV's order
// depends on its hash table: of several sources matched to one scene point the smallest distance
// wins, ties go to the smaller source index; pairs enter the solve in ascending scene index.
// ------------------------------------------------------------------------------------------
namespace orc {
namespace cvicp {

static const double kEps = 1.192092896e-07;  // OpenCV's EPS (FLT_EPSILON)

// [CV] ppf_helpers.cpp : transformPCPose — points through the 4x4 (homogeneous divide), normals through its 3x3, renormalised
static void transform_pc_pose(const std::vector<float>& pc, const double* P, std::vector<float>& out) {
  const size_t n = pc.size() / 6;
  out.resize(pc.size());
  for (size_t i = 0; i < n; ++i) {
    const float* r = &pc[6 * i];
    double p[4];
    for (int k = 0; k < 4; ++k) p[k] = P[4 * k] * r[0] + P[4 * k + 1] * r[1] + P[4 * k + 2] * r[2] + P[4 * k + 3];
    float* o = &out[6 * i];
    if (std::fabs(p[3]) > kEps) {
      o[0] = static_cast<float>(p[0] / p[3]);
      o[1] = static_cast<float>(p[1] / p[3]);
      o[2] = static_cast<float>(p[2] / p[3]);
    } else {
      o[0] = o[1] = o[2] = 0.0f;
    }
    double nn[3];
    for (int k = 0; k < 3; ++k) nn[k] = P[4 * k] * r[3] + P[4 * k + 1] * r[4] + P[4 * k + 2] * r[5];
    const double norm = std::sqrt(nn[0] * nn[0] + nn[1] * nn[1] + nn[2] * nn[2]);
    if (norm > kEps) {
      o[3] = static_cast<float>(nn[0] / norm);
      o[4] = static_cast<float>(nn[1] / norm);
      o[5] = static_cast<float>(nn[2] / norm);
    } else {
      o[3] = o[4] = o[5] = 0.0f;
    }
  }
}

// [CV] ppf_helpers.cpp : samplePCUniform — numRows = rows / sampleStep (integer division) rows, taken at 0, step, 2 step, ...
static void sample_uniform(const std::vector<float>& pc, int step, std::vector<float>& out) {
  const size_t n = pc.size() / 6, rows = n / static_cast<size_t>(step);
  out.clear();
  for (size_t c = 0; c < rows; ++c) out.insert(out.end(), &pc[6 * c * step], &pc[6 * c * step] + 6);
}

static int cv_round(double v) { return static_cast<int>(std::lrint(v)); }  // cvRound: to nearest, ties to even

static void mat44_mul(const double* A, const double* B, double* C) {
  double t[16];
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) {
      double a = 0.0;
      for (int k = 0; k < 4; ++k) a += A[4 * r + k] * B[4 * k + c];
      t[4 * r + c] = a;
    }
  std::memcpy(C, t, sizeof(t));
}

// [CV] c_utils.hpp : eulerToDCM (R = Rx * (Ry * Rz)) + rtToPose
static void pose_from_euler(const double* e, const double* t, double* P) {
  const double cx = std::cos(e[0]), sx = std::sin(e[0]), cy = std::cos(e[1]), sy = std::sin(e[1]), cz = std::cos(e[2]),
               sz = std::sin(e[2]);
  const double Rx[9] = {1, 0, 0, 0, cx, -sx, 0, sx, cx};
  const double Ry[9] = {cy, 0, sy, 0, 1, 0, -sy, 0, cy};
  const double Rz[9] = {cz, -sz, 0, sz, cz, 0, 0, 0, 1};
  double T[9], R[9];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) T[3 * r + c] = Ry[3 * r] * Rz[c] + Ry[3 * r + 1] * Rz[3 + c] + Ry[3 * r + 2] * Rz[6 + c];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) R[3 * r + c] = Rx[3 * r] * T[c] + Rx[3 * r + 1] * T[3 + c] + Rx[3 * r + 2] * T[6 + c];
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) P[4 * r + c] = R[3 * r + c];
    P[4 * r + 3] = t[r];
  }
  P[12] = P[13] = P[14] = 0.0;
  P[15] = 1.0;
}

// least squares of the n x 6 system through its normal equations (Gaussian elimination with partial pivoting, double);
// cv::solve(A, b, x, DECOMP_SVD) gives the same minimiser for a full-rank A
static bool solve6(double* N /*6x6*/, double* r /*6*/, double* x) {
  for (int c = 0; c < 6; ++c) {
    int best = c;
    for (int rr = c + 1; rr < 6; ++rr)
      if (std::fabs(N[6 * rr + c]) > std::fabs(N[6 * best + c])) best = rr;
    if (!(std::fabs(N[6 * best + c]) > 0.0)) return false;
    if (best != c) {
      for (int k = 0; k < 6; ++k) std::swap(N[6 * c + k], N[6 * best + k]);
      std::swap(r[c], r[best]);
    }
    for (int rr = c + 1; rr < 6; ++rr) {
      const double f = N[6 * rr + c] / N[6 * c + c];
      for (int k = c; k < 6; ++k) N[6 * rr + k] -= f * N[6 * c + k];
      r[rr] -= f * r[c];
    }
  }
  for (int c = 5; c >= 0; --c) {
    double a = r[c];
    for (int k = c + 1; k < 6; ++k) a -= N[6 * c + k] * x[k];
    x[c] = a / N[6 * c + c];
  }
  return true;
}

static float nth_value(std::vector<float> v, size_t nth) {
  std::nth_element(v.begin(), v.begin() + static_cast<long>(nth), v.end());
  return v[nth];
}

// [CV] icp.cpp : getRejectionThreshold — median + scale * 1.48257968 * MAD, both taken at index m / 2
static float rejection_threshold(const std::vector<float>& d, float scale) {
  const size_t m = d.size();
  const float med = nth_value(d, m / 2);
  std::vector<float> a(m);
  for (size_t i = 0; i < m; ++i) a[i] = std::fabs(d[i] - med);
  const float mad = nth_value(a, m / 2);
  const float sgm = 1.48257968f * mad;
  return scale * sgm + med;
}

// one pose; src = the model already moved by the pose (n x 6 floats); returns the 4x4 ICP pose and the residual
static void register_one(const std::vector<float>& src_in, const std::vector<float>& dst_in, const peb_cvicp_params& prm,
                         double* pose, double* residual) {
  const int n = static_cast<int>(src_in.size() / 6);
  std::vector<float> src0 = src_in, dst0 = dst_in;
  auto mean_cols = [](const std::vector<float>& pc, double* m) {
    m[0] = m[1] = m[2] = 0.0;
    const size_t r = pc.size() / 6;
    for (size_t i = 0; i < r; ++i)
      for (int k = 0; k < 3; ++k) m[k] += pc[6 * i + k];
    for (int k = 0; k < 3; ++k) m[k] /= static_cast<double>(r);
  };
  double ms[3], md[3], mean_avg[3];
  mean_cols(src0, ms);
  mean_cols(dst0, md);
  for (int k = 0; k < 3; ++k) mean_avg[k] = 0.5 * (ms[k] + md[k]);
  auto subtract = [&](std::vector<float>& pc) {
    for (size_t i = 0; i < pc.size() / 6; ++i)
      for (int k = 0; k < 3; ++k) pc[6 * i + k] -= static_cast<float>(mean_avg[k]);
  };
  subtract(src0);
  subtract(dst0);
  auto dist_to_origin = [](const std::vector<float>& pc) {
    double d = 0.0;
    for (size_t i = 0; i < pc.size() / 6; ++i) {
      const double x = pc[6 * i], y = pc[6 * i + 1], z = pc[6 * i + 2];
      d += std::sqrt(x * x + y * y + z * z);
    }
    return d;
  };
  const double scale = static_cast<double>(n) / ((dist_to_origin(src0) + dist_to_origin(dst0)) * 0.5);
  auto scale_pc = [&](std::vector<float>& pc) {
    for (size_t i = 0; i < pc.size() / 6; ++i)
      for (int k = 0; k < 3; ++k) pc[6 * i + k] = static_cast<float>(pc[6 * i + k] * scale);
  };
  scale_pc(src0);
  scale_pc(dst0);
  for (int i = 0; i < 16; ++i) pose[i] = (i % 5 == 0) ? 1.0 : 0.0;
  double temp_residual = 0.0;
  const bool robust = prm.rejection_scale > 0.0f;
  for (int level = prm.num_levels - 1; level >= 0; --level) {
    const double div = std::pow(2.0, static_cast<double>(level));
    const int num_samples = cv_round(static_cast<double>(n) / div);
    const double tol_p = static_cast<double>(prm.tolerance) * static_cast<double>(level + 1) * (level + 1);
    const int max_it = cv_round(static_cast<double>(prm.iterations) / (level + 1));
    std::vector<float> moved_full, src_pct, dst_pcs;
    transform_pc_pose(src0, pose, moved_full);
    const int step = std::max(1, cv_round(static_cast<double>(n) / static_cast<double>(std::max(num_samples, 1))));
    sample_uniform(moved_full, step, src_pct);
    sample_uniform(dst0, step, dst_pcs);
    if (src_pct.empty() || dst_pcs.empty()) continue;  // (a level coarser than the clouds: nothing to iterate on)
    KdTree tree;
    tree.build(dst_pcs.data(), dst_pcs.size() / 6, 24);
    double fval_old = 9999999999.0, fval_perc = 0.0, fval_min = 9999999999.0;
    std::vector<float> src_moved = src_pct;
    const size_t m = src_pct.size() / 6;
    std::vector<int> idx(m);
    std::vector<float> dist(m);
    double pose_x[16];
    for (int i = 0; i < 16; ++i) pose_x[i] = (i % 5 == 0) ? 1.0 : 0.0;
    int it = 0;
    while (!(fval_perc < (1.0 + tol_p) && fval_perc > (1.0 - tol_p)) && it < max_it) {
      for (size_t i = 0; i < m; ++i) {
        int j = -1;
        float d2 = 0.0f;
        tree.knn(&src_moved[6 * i], 1, &j, &d2);
        idx[i] = j;
        dist[i] = d2;
      }
      float thr = std::numeric_limits<float>::infinity();
      if (robust) thr = rejection_threshold(dist, prm.rejection_scale);
      // picky ICP: per scene point the closest accepted source (ties: smaller source index)
      std::vector<int> winner(dst_pcs.size() / 6, -1);
      for (size_t i = 0; i < m; ++i) {
        if (idx[i] < 0) continue;
        if (robust && !(dist[i] < thr)) continue;
        int& w = winner[static_cast<size_t>(idx[i])];
        if (w < 0 || dist[i] < dist[static_cast<size_t>(w)]) w = static_cast<int>(i);
      }
      double N[36] = {0}, r[6] = {0}, fsum = 0.0;
      int sel = 0;
      for (size_t j = 0; j < winner.size(); ++j) {
        if (winner[j] < 0) continue;
        ++sel;
        const float* sp = &src_pct[6 * static_cast<size_t>(winner[j])];
        const float* dp = &dst_pcs[6 * j];
        const double sx = sp[0], sy = sp[1], sz = sp[2], dx = dp[0], dy = dp[1], dz = dp[2], nx = dp[3], ny = dp[4], nz = dp[5];
        const double row[6] = {sy * nz - sz * ny, sz * nx - sx * nz, sx * ny - sy * nx, nx, ny, nz};
        const double b = (dx - sx) * nx + (dy - sy) * ny + (dz - sz) * nz;
        for (int a = 0; a < 6; ++a) {
          for (int c = 0; c < 6; ++c) N[6 * a + c] += row[a] * row[c];
          r[a] += row[a] * b;
        }
        for (int k = 0; k < 6; ++k) {
          const double e = static_cast<double>(sp[k]) - static_cast<double>(dp[k]);
          fsum += e * e;
        }
      }
      if (sel < 6) break;
      double x[6];
      if (!solve6(N, r, x)) break;
      bool bad = false;
      for (int k = 0; k < 6; ++k) bad = bad || std::isnan(x[k]);
      if (bad) break;
      pose_from_euler(x, x + 3, pose_x);
      transform_pc_pose(src_pct, pose_x, src_moved);
      const double fval = std::sqrt(fsum) / static_cast<double>(m);
      fval_perc = fval / fval_old;
      fval_old = fval;
      if (fval < fval_min) fval_min = fval;
      ++it;
    }
    mat44_mul(pose_x, pose, pose);
    temp_residual = fval_min;
  }
  // undo the normalisation: t = t / scale + meanAvg - R * meanAvg
  for (int r = 0; r < 3; ++r) {
    const double rm = pose[4 * r] * mean_avg[0] + pose[4 * r + 1] * mean_avg[1] + pose[4 * r + 2] * mean_avg[2];
    pose[4 * r + 3] = pose[4 * r + 3] / scale + mean_avg[r] - rm;
  }
  *residual = temp_residual;
}

}  // namespace cvicp
}  // namespace orc

extern "C" {

// poses: n_poses x 16 doubles (row-major 4x4), updated in place (Pose3D::appendPose); residuals: n_poses
ORC_API void orc_cvicp_register(const float* model, size_t n_model, const float* scene, size_t n_scene,
                                const peb_cvicp_params* prm, double* poses, size_t n_poses, double* residuals) {
  std::vector<float> m(model, model + 6 * n_model), sc(scene, scene + 6 * n_scene);
#pragma omp parallel for schedule(dynamic)
  for (long long i = 0; i < static_cast<long long>(n_poses); ++i) {
    std::vector<float> moved;
    orc::cvicp::transform_pc_pose(m, poses + 16 * i, moved);
    double icp_pose[16], res = 0.0;
    orc::cvicp::register_one(moved, sc, *prm, icp_pose, &res);
    orc::cvicp::mat44_mul(icp_pose, poses + 16 * i, poses + 16 * i);
    residuals[i] = res;
  }
}

ORC_API int orc_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

}  // extern "C"
