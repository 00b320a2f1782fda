// ppf_oracle.cpp — CPU ORACLE of the coarse matcher.  TEST INFRASTRUCTURE ONLY, NOT PRODUCT CODE
// (same rules as pcl_oracle.cpp: only tests/, __graft_entry__.smoke() and bench.py's CPU legs load it).
//
// What it restates: cv::ppf_match_3d::PPF3DDetector (opencv_contrib 4.x, modules/surface_matching/src/
// ppf_match_3d.cpp, ppf_helpers.cpp, pose_3d.cpp, c_utils.hpp) as the reference uses it —
//   PPF3DDetector(0.03, 0.03, 40).trainModel(model)            pose_estimation/src/opencv_surface_match.cpp:37-51
//   detector.match(scene_with_normals, results, 1.0, 0.03)     pose_estimation/src/opencv_surface_match.cpp:65
// opencv_contrib is neither in /root/reference nor in this image (cv2 4.13 is built without ppf_match_3d:
// profiles/r2_probe_image.txt), so every function below follows the upstream algorithm FROM RECOLLECTION and names
// the upstream function it restates.  PARITY UNPINNED: there is no OpenCV output to compare with.  What the tests
// pin instead (tests/test_ppf.py): the pieces against independent numpy restatements (sampling, the four-component
// feature, the alpha angle, the voting of one reference point, the clustering), and the whole matcher on what a
// matcher must do — the pose of a synthetic object recovered from a cluttered scene.
//
// ONE DELIBERATE DIFFERENCE, stated here because it defines what "votes" means for the CUDA path as well:
// OpenCV stores the model's point pairs in a chained hash table keyed by a MurmurHash of the four quantised feature
// components and, when voting, walks the whole chain of a bucket WITHOUT comparing keys — two different features
// that fall into the same bucket vote for each other.  Which features collide depends on the hash function, on its
// 32/64-bit build variant and on the table size; none of it can be restated reliably and the colliding votes are
// noise by construction.  Here a scene pair votes for exactly the model pairs with the SAME quantised feature
// (an ideal hash).  Vote counts are therefore a lower bound of OpenCV's; the argmax and the poses agree wherever
// OpenCV's collisions do not flip a maximum.
//
// Arithmetic: double where OpenCV uses double (features, alpha, transforms), float where it stores float (the
// sampled cloud, the model's alpha, the distance step); libm acos / atan2 / sin.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <map>
#include <unordered_map>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/pe_b200.h"  // POD params / pose layouts only

namespace orc {
namespace ppf {

static const double kEps = 1.192092896e-07;  // [CV] c_utils.hpp : EPS

struct V3 {
  double x, y, z;
};
static inline V3 operator-(const V3& a, const V3& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline double dot(const V3& a, const V3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline double norm(const V3& a) { return std::sqrt(dot(a, a)); }

// [CV] ppf_helpers.cpp : computeBboxStd — min / max per axis (float)
static void bbox(const float* pc, size_t n, size_t cols, float lo[3], float hi[3]) {
  for (int a = 0; a < 3; ++a) {
    lo[a] = std::numeric_limits<float>::infinity();
    hi[a] = -std::numeric_limits<float>::infinity();
  }
  for (size_t i = 0; i < n; ++i) {
    const float* p = pc + i * cols;
    if (!(std::isfinite(p[0]) && std::isfinite(p[1]) && std::isfinite(p[2]))) continue;
    for (int a = 0; a < 3; ++a) {
      if (p[a] < lo[a]) lo[a] = p[a];
      if (p[a] > hi[a]) hi[a] = p[a];
    }
  }
}

// [CV] ppf_helpers.cpp : samplePCByQuantization(pc, xrange, yrange, zrange, sampleStep, weightByCenter = 0)
// cells of a (numSamplesDim)^3 lattice over the bounding box; the points (and normals) of a cell are averaged in
// double in input order, the normal re-normalised; output in ascending cell index.  The cell index uses
// numSamplesDim as the multiplier although a coordinate at the upper bound quantises to numSamplesDim itself
// (such points alias into the next row's first cell: upstream behaviour, kept).
static void sample_by_quantization(const float* pc, size_t n, size_t cols, const float lo[3], const float hi[3],
                                   float sample_step, std::vector<float>& out) {
  const int nsd = static_cast<int>(1.0 / sample_step);
  const float xr = hi[0] - lo[0], yr = hi[1] - lo[1], zr = hi[2] - lo[2];
  const size_t cells = static_cast<size_t>(nsd + 1) * (nsd + 1) * (nsd + 1);
  std::vector<std::vector<int>> map(cells);
  for (size_t i = 0; i < n; ++i) {
    const float* p = pc + i * cols;
    if (!(std::isfinite(p[0]) && std::isfinite(p[1]) && std::isfinite(p[2]))) continue;  // (upstream expects clean clouds)
    const int xc = static_cast<int>(static_cast<float>(nsd) * (p[0] - lo[0]) / xr);
    const int yc = static_cast<int>(static_cast<float>(nsd) * (p[1] - lo[1]) / yr);
    const int zc = static_cast<int>(static_cast<float>(nsd) * (p[2] - lo[2]) / zr);
    const int index = xc * nsd * nsd + yc * nsd + zc;
    map[static_cast<size_t>(index)].push_back(static_cast<int>(i));
  }
  out.clear();
  for (size_t c = 0; c < cells; ++c) {
    const std::vector<int>& cur = map[c];
    const int cn = static_cast<int>(cur.size());
    if (cn == 0) continue;
    double px = 0, py = 0, pz = 0, nx = 0, ny = 0, nz = 0;
    for (int j = 0; j < cn; ++j) {
      const float* p = pc + static_cast<size_t>(cur[j]) * cols;
      px += static_cast<double>(p[0]);
      py += static_cast<double>(p[1]);
      pz += static_cast<double>(p[2]);
      if (cols == 6) {
        nx += static_cast<double>(p[3]);
        ny += static_cast<double>(p[4]);
        nz += static_cast<double>(p[5]);
      }
    }
    px /= static_cast<double>(cn);
    py /= static_cast<double>(cn);
    pz /= static_cast<double>(cn);
    float rec[6] = {static_cast<float>(px), static_cast<float>(py), static_cast<float>(pz), 0.f, 0.f, 0.f};
    if (cols == 6) {
      nx /= static_cast<double>(cn);
      ny /= static_cast<double>(cn);
      nz /= static_cast<double>(cn);
      const double nn = std::sqrt(nx * nx + ny * ny + nz * nz);
      if (nn > kEps) {
        rec[3] = static_cast<float>(nx / nn);
        rec[4] = static_cast<float>(ny / nn);
        rec[5] = static_cast<float>(nz / nn);
      }
    }
    out.insert(out.end(), rec, rec + 6);
  }
}

// [CV] c_utils.hpp : TAngle3Normalized — acos(a . b) for unit vectors (the 4.x form)
static inline double angle3_normalized(const V3& a, const V3& b) { return std::acos(dot(a, b)); }

// [CV] ppf_match_3d.cpp : PPF3DDetector::computePPFFeatures.  Returns false for coincident points (f stays 0).
static bool ppf_features(const V3& p1, const V3& n1, const V3& p2, const V3& n2, double f[4]) {
  f[0] = f[1] = f[2] = f[3] = 0.0;
  V3 d = p2 - p1;
  f[3] = norm(d);
  if (f[3] <= kEps) return false;
  const double inv = 1.0 / f[3];
  d = {d.x * inv, d.y * inv, d.z * inv};
  f[0] = angle3_normalized(n1, d);
  f[1] = angle3_normalized(n2, d);
  f[2] = angle3_normalized(n1, n2);
  return true;
}

// the four quantised components ([CV] hashPPF: (int)(f / step)); the hash of upstream is replaced by the key itself
// A NaN component (acos of a dot product that rounding pushed past 1) would be undefined behaviour in upstream's
// cast; here such a pair gets a negative component and is left out of the table and of the voting (valid_key).
static inline void ppf_key(const double f[4], double angle_step, double distance_step, int key[4]) {
  const double q[4] = {f[0] / angle_step, f[1] / angle_step, f[2] / angle_step, f[3] / distance_step};
  for (int k = 0; k < 4; ++k) key[k] = (q[k] == q[k] && q[k] < 32767.0) ? static_cast<int>(q[k]) : -1;
}
static inline bool valid_key(const int key[4]) { return key[0] >= 0 && key[1] >= 0 && key[2] >= 0 && key[3] >= 0; }
static inline uint64_t pack_key(const int key[4]) {
  return (static_cast<uint64_t>(static_cast<uint16_t>(key[0])) << 48) | (static_cast<uint64_t>(static_cast<uint16_t>(key[1])) << 32) |
         (static_cast<uint64_t>(static_cast<uint16_t>(key[2])) << 16) | static_cast<uint64_t>(static_cast<uint16_t>(key[3]));
}

// [CV] c_utils.hpp : aaToR — Rodrigues: R = cos I + (1 - cos) a a^T + sin [a]x   (row-major)
static void aa_to_r(const V3& a, double angle, double R[9]) {
  const double c = std::cos(angle), s = std::sin(angle), omc = 1.0 - c;
  R[0] = c + omc * a.x * a.x;
  R[1] = omc * a.x * a.y - s * a.z;
  R[2] = omc * a.x * a.z + s * a.y;
  R[3] = omc * a.y * a.x + s * a.z;
  R[4] = c + omc * a.y * a.y;
  R[5] = omc * a.y * a.z - s * a.x;
  R[6] = omc * a.z * a.x - s * a.y;
  R[7] = omc * a.z * a.y + s * a.x;
  R[8] = c + omc * a.z * a.z;
}

// [CV] c_utils.hpp : computeTransformRT — the rotation that takes n1 onto the x axis, t = -R p1
static void transform_rt(const V3& p1, const V3& n1, double R[9], double t[3]) {
  const double angle = std::acos(n1.x);
  V3 axis = {0.0, n1.z, -n1.y};
  if (n1.y == 0.0 && n1.z == 0.0) {
    axis.y = 1.0;
    axis.z = 0.0;
  } else {
    const double nn = norm(axis);
    if (nn > kEps) {  // [CV] TNormalize3
      const double inv = 1.0 / nn;
      axis = {axis.x * inv, axis.y * inv, axis.z * inv};
    }
  }
  aa_to_r(axis, angle, R);
  t[0] = -(R[0] * p1.x + R[1] * p1.y + R[2] * p1.z);
  t[1] = -(R[3] * p1.x + R[4] * p1.y + R[5] * p1.z);
  t[2] = -(R[6] * p1.x + R[7] * p1.y + R[8] * p1.z);
}

// the planar angle of p2 in the frame of (p1, n1): [CV] computeAlpha and the same lines inside match()
static inline double alpha_of(const double R[9], const double t[3], const V3& p2) {
  const double my = t[1] + (R[3] * p2.x + R[4] * p2.y + R[5] * p2.z);
  const double mz = t[2] + (R[6] * p2.x + R[7] * p2.y + R[8] * p2.z);
  double alpha = std::atan2(-mz, my);
  if (alpha != alpha) return 0.0;
  if (std::sin(alpha) * mz < 0.0) alpha = -alpha;
  return -alpha;
}

static inline V3 pt(const float* rec) { return {rec[0], rec[1], rec[2]}; }
static inline V3 nr(const float* rec) { return {rec[3], rec[4], rec[5]}; }

struct Node {
  int i;        // model reference point
  float alpha;  // stored as float like ppf.ptr<float>(ppfInd)[4]
};

struct Detector {
  peb_ppf_params prm;
  double angle_step = 0.0;      // radians
  float distance_step = 0.0f;   // float like upstream
  double position_threshold = 0.0, rotation_threshold = 0.0;
  std::vector<float> sampled;   // n x 6
  int n = 0;
  std::unordered_map<uint64_t, std::vector<Node>> table;
};

// [CV] PPF3DDetector::PPF3DDetector + setSearchParams + trainModel
static void train(Detector& d, const float* model6, size_t n_model, const peb_ppf_params& prm) {
  d.prm = prm;
  d.angle_step = (360.0 / prm.num_angles) * M_PI / 180.0;
  d.position_threshold = prm.position_threshold < 0 ? prm.relative_sampling_step : prm.position_threshold;
  d.rotation_threshold = prm.rotation_threshold < 0 ? ((360 / d.angle_step) / 180.0 * M_PI) : prm.rotation_threshold;
  float lo[3], hi[3];
  bbox(model6, n_model, 6, lo, hi);
  const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
  const float diameter = std::sqrt(dx * dx + dy * dy + dz * dz);
  d.distance_step = static_cast<float>(diameter * prm.relative_sampling_step);
  sample_by_quantization(model6, n_model, 6, lo, hi, static_cast<float>(prm.relative_sampling_step), d.sampled);
  d.n = static_cast<int>(d.sampled.size() / 6);
  d.table.clear();
  for (int i = 0; i < d.n; ++i) {
    const V3 p1 = pt(&d.sampled[6 * i]), n1 = nr(&d.sampled[6 * i]);
    double R[9], t[3];
    transform_rt(p1, n1, R, t);
    for (int j = 0; j < d.n; ++j) {
      if (i == j) continue;
      const V3 p2 = pt(&d.sampled[6 * j]), n2 = nr(&d.sampled[6 * j]);
      double f[4];
      ppf_features(p1, n1, p2, n2, f);  // (coincident sampled points keep f = 0 and are inserted like upstream)
      int key[4];
      ppf_key(f, d.angle_step, d.distance_step, key);
      if (!valid_key(key)) continue;
      d.table[pack_key(key)].push_back({i, static_cast<float>(alpha_of(R, t, p2))});
    }
  }
}

static void mat44_mul(const double* A, const double* B, double* C) {
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) {
      double s = 0.0;
      for (int k = 0; k < 4; ++k) s += A[4 * r + k] * B[4 * k + c];
      C[4 * r + c] = s;
    }
}
static void rt_to_pose(const double R[9], const double t[3], double P[16]) {
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) P[4 * r + c] = R[3 * r + c];
    P[4 * r + 3] = t[r];
  }
  P[12] = P[13] = P[14] = 0.0;
  P[15] = 1.0;
}

// [CV] c_utils.hpp : dcmToQuat (w x y z)
static void dcm_to_quat(const double R[9], double q[4]) {
  const double tr = R[0] + R[4] + R[8];
  if (tr > 0) {
    q[0] = tr + 1.0;
    q[1] = R[5] - R[7];
    q[2] = R[6] - R[2];
    q[3] = R[1] - R[3];
  } else {
    int i = 0;
    if (R[4] > R[0]) i = 1;
    if (R[8] > R[3 * i + i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    q[i + 1] = R[3 * i + i] - R[3 * j + j] - R[3 * k + k] + 1.0;
    q[j + 1] = R[3 * i + j] + R[3 * j + i];
    q[k + 1] = R[3 * i + k] + R[3 * k + i];
    q[0] = R[3 * j + k] - R[3 * k + j];
  }
  const double nn = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  const double s = 1.0 / nn;  // (upstream: q *= 0.5 / sqrt(...); the unit quaternion is the same)
  for (int a = 0; a < 4; ++a) q[a] *= s;
}
// [CV] c_utils.hpp : quatToDCM (w x y z)
static void quat_to_dcm(const double q[4], double R[9]) {
  const double sqw = q[0] * q[0], sqx = q[1] * q[1], sqy = q[2] * q[2], sqz = q[3] * q[3];
  R[0] = sqx - sqy - sqz + sqw;
  R[4] = -sqx + sqy - sqz + sqw;
  R[8] = -sqx - sqy + sqz + sqw;
  double t1 = q[1] * q[2], t2 = q[3] * q[0];
  R[3] = 2.0 * (t1 + t2);
  R[1] = 2.0 * (t1 - t2);
  t1 = q[1] * q[3];
  t2 = q[2] * q[0];
  R[6] = 2.0 * (t1 - t2);
  R[2] = 2.0 * (t1 + t2);
  t1 = q[2] * q[3];
  t2 = q[1] * q[0];
  R[7] = 2.0 * (t1 + t2);
  R[5] = 2.0 * (t1 - t2);
}
// the rotation angle from the trace: [CV] pose_3d.cpp : Pose3D::updatePose
static double angle_of(const double R[9]) {
  const double trace = R[0] + R[4] + R[8];
  if (std::fabs(trace - 3) <= kEps) return 0.0;
  if (std::fabs(trace + 1) <= kEps) return M_PI;
  return std::acos((trace - 1) / 2);
}
static void update_pose(peb_ppf_pose& p, const double P[16]) {
  std::memcpy(p.pose, P, sizeof(p.pose));
  double R[9];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) R[3 * r + c] = P[4 * r + c];
  p.t[0] = P[3];
  p.t[1] = P[7];
  p.t[2] = P[11];
  p.angle = angle_of(R);
  dcm_to_quat(R, p.q);
}
// [CV] Pose3D::updatePoseQuat
static void update_pose_quat(peb_ppf_pose& p, const double q[4], const double t[3]) {
  double R[9];
  quat_to_dcm(q, R);
  for (int a = 0; a < 4; ++a) p.q[a] = q[a];
  for (int a = 0; a < 3; ++a) p.t[a] = t[a];
  double P[16];
  rt_to_pose(R, t, P);
  std::memcpy(p.pose, P, sizeof(p.pose));
  p.angle = angle_of(R);
}

// [CV] PPF3DDetector::clusterPoses (+ matchPose, pose3DPtrCompare, sortPoseClusters).  std::sort is not stable
// upstream; ties keep the input order here (a documented choice: the CUDA path uses the same rule).
static void cluster_poses(const Detector& d, std::vector<peb_ppf_pose> poses, std::vector<peb_ppf_pose>& out) {
  std::stable_sort(poses.begin(), poses.end(), [](const peb_ppf_pose& a, const peb_ppf_pose& b) { return a.num_votes > b.num_votes; });
  struct Cluster {
    std::vector<int> members;
    uint64_t votes = 0;
  };
  std::vector<Cluster> clusters;
  for (size_t i = 0; i < poses.size(); ++i) {
    bool assigned = false;
    for (size_t c = 0; c < clusters.size() && !assigned; ++c) {
      const peb_ppf_pose& centre = poses[static_cast<size_t>(clusters[c].members[0])];
      const double dvx = centre.t[0] - poses[i].t[0], dvy = centre.t[1] - poses[i].t[1], dvz = centre.t[2] - poses[i].t[2];
      const double dn = std::sqrt(dvx * dvx + dvy * dvy + dvz * dvz);
      const double phi = std::fabs(poses[i].angle - centre.angle);
      if (phi < d.rotation_threshold && dn < d.position_threshold) {
        clusters[c].members.push_back(static_cast<int>(i));
        clusters[c].votes += poses[i].num_votes;
        assigned = true;
      }
    }
    if (!assigned) {
      Cluster c;
      c.members.push_back(static_cast<int>(i));
      c.votes = poses[i].num_votes;
      clusters.push_back(c);
    }
  }
  std::stable_sort(clusters.begin(), clusters.end(), [](const Cluster& a, const Cluster& b) { return a.votes > b.votes; });
  out.clear();
  for (const Cluster& c : clusters) {
    double q[4] = {0, 0, 0, 0}, t[3] = {0, 0, 0};
    const int sz = static_cast<int>(c.members.size());
    if (d.prm.use_weighted_avg) {
      double wsum = 0;
      for (int m : c.members) {
        const peb_ppf_pose& p = poses[static_cast<size_t>(m)];
        const double w = static_cast<double>(p.num_votes);
        for (int a = 0; a < 4; ++a) q[a] += w * p.q[a];
        for (int a = 0; a < 3; ++a) t[a] += w * p.t[a];
        wsum += w;
      }
      for (int a = 0; a < 3; ++a) t[a] *= 1.0 / wsum;
      for (int a = 0; a < 4; ++a) q[a] *= 1.0 / wsum;
    } else {
      for (int m : c.members) {
        const peb_ppf_pose& p = poses[static_cast<size_t>(m)];
        for (int a = 0; a < 4; ++a) q[a] += p.q[a];
        for (int a = 0; a < 3; ++a) t[a] += p.t[a];
      }
      for (int a = 0; a < 3; ++a) t[a] *= 1.0 / sz;
      for (int a = 0; a < 4; ++a) q[a] *= 1.0 / sz;
    }
    peb_ppf_pose r = poses[static_cast<size_t>(c.members[0])];
    update_pose_quat(r, q, t);
    r.num_votes = c.votes;
    out.push_back(r);
  }
}

// [CV] PPF3DDetector::match: the voting of every scene reference point, then clusterPoses
static void match(const Detector& d, const float* scene6, size_t n_scene, double rel_scene_sample_step,
                  double rel_scene_distance, std::vector<float>& sampled_out, std::vector<peb_ppf_pose>& raw,
                  std::vector<peb_ppf_pose>& results) {
  const int num_angles = static_cast<int>(std::floor(2 * M_PI / d.angle_step));
  const int step = static_cast<int>(1.0 / rel_scene_sample_step);
  float lo[3], hi[3];
  bbox(scene6, n_scene, 6, lo, hi);
  std::vector<float>& s = sampled_out;
  sample_by_quantization(scene6, n_scene, 6, lo, hi, static_cast<float>(rel_scene_distance), s);
  const int m = static_cast<int>(s.size() / 6);
  const int n_ref = (m + step - 1) / step;
  raw.assign(static_cast<size_t>(n_ref), peb_ppf_pose());
#pragma omp parallel for schedule(dynamic, 4)
  for (int ri = 0; ri < n_ref; ++ri) {
    const int i = ri * step;
    const V3 p1 = pt(&s[6 * i]), n1 = nr(&s[6 * i]);
    double Rsg[9], tsg[3];
    transform_rt(p1, n1, Rsg, tsg);
    std::vector<uint32_t> acc(static_cast<size_t>(num_angles) * d.n, 0u);
    for (int j = 0; j < m; ++j) {
      if (i == j) continue;
      const V3 p2 = pt(&s[6 * j]), n2 = nr(&s[6 * j]);
      double f[4];
      ppf_features(p1, n1, p2, n2, f);
      int key[4];
      ppf_key(f, d.angle_step, d.distance_step, key);
      if (!valid_key(key)) continue;
      const double alpha_scene = alpha_of(Rsg, tsg, p2);
      const auto it = d.table.find(pack_key(key));
      if (it == d.table.end()) continue;
      for (const Node& nd : it->second) {
        const double alpha = static_cast<double>(nd.alpha) - alpha_scene;
        const int alpha_index = static_cast<int>(num_angles * (alpha + 2 * M_PI) / (4 * M_PI));
        // (alpha == 2 pi exactly would index one past the row upstream: it goes to the last bin here and on the device)
        const int a_idx = alpha_index < num_angles ? alpha_index : num_angles - 1;
        acc[static_cast<size_t>(nd.i) * num_angles + a_idx]++;
      }
    }
    uint32_t max_votes = 0;
    int ref_max = 0, alpha_max = 0;
    for (int k = 0; k < d.n; ++k)
      for (int a = 0; a < num_angles; ++a) {
        const uint32_t v = acc[static_cast<size_t>(k) * num_angles + a];
        if (v > max_votes) {
          max_votes = v;
          ref_max = k;
          alpha_max = a;
        }
      }
    // pose = Tsg^-1 * Rx(alpha) * Tmg
    double RInv[9], tInv[3];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) RInv[3 * r + c] = Rsg[3 * c + r];
    for (int r = 0; r < 3; ++r) tInv[r] = -(RInv[3 * r] * tsg[0] + RInv[3 * r + 1] * tsg[1] + RInv[3 * r + 2] * tsg[2]);
    double TsgInv[16], Tmg[16], Talpha[16], tmp[16], rawPose[16];
    rt_to_pose(RInv, tInv, TsgInv);
    double Rmg[9], tmg[3];
    transform_rt(pt(&d.sampled[6 * ref_max]), nr(&d.sampled[6 * ref_max]), Rmg, tmg);
    rt_to_pose(Rmg, tmg, Tmg);
    const double alpha = (alpha_max * (4 * M_PI)) / num_angles - 2 * M_PI;
    const double sa = std::sin(alpha), ca = std::cos(alpha);
    const double Rx[9] = {1, 0, 0, 0, ca, -sa, 0, sa, ca};  // [CV] getUnitXRotation
    const double tz[3] = {0, 0, 0};
    rt_to_pose(Rx, tz, Talpha);
    mat44_mul(Talpha, Tmg, tmp);
    mat44_mul(TsgInv, tmp, rawPose);
    peb_ppf_pose p;
    std::memset(&p, 0, sizeof(p));
    p.alpha = alpha;
    p.model_index = static_cast<uint64_t>(ref_max);
    p.num_votes = max_votes;
    p.residual = 0.0;
    update_pose(p, rawPose);
    raw[static_cast<size_t>(ri)] = p;
  }
  cluster_poses(d, raw, results);
}

}  // namespace ppf
}  // namespace orc

#define ORC_API __attribute__((visibility("default")))
extern "C" {

ORC_API void* orc_ppf_train(const float* model6, size_t n, const peb_ppf_params* prm) {
  auto* d = new orc::ppf::Detector();
  orc::ppf::train(*d, model6, n, *prm);
  return d;
}
ORC_API void orc_ppf_destroy(void* h) { delete static_cast<orc::ppf::Detector*>(h); }
ORC_API size_t orc_ppf_model_size(const void* h) { return static_cast<size_t>(static_cast<const orc::ppf::Detector*>(h)->n); }
ORC_API void orc_ppf_model_sampled(const void* h, float* out6) {
  const auto* d = static_cast<const orc::ppf::Detector*>(h);
  std::memcpy(out6, d->sampled.data(), d->sampled.size() * sizeof(float));
}
ORC_API double orc_ppf_distance_step(const void* h) { return static_cast<const orc::ppf::Detector*>(h)->distance_step; }
// number of model pairs stored under the quantised feature of the ordered pair (i, j) of the sampled model
ORC_API size_t orc_ppf_table_size(const void* h) {
  size_t s = 0;
  for (const auto& kv : static_cast<const orc::ppf::Detector*>(h)->table) s += kv.second.size();
  return s;
}
// match: raw = one pose per scene reference point (before clustering), results = clustered poses (most votes first).
// Returns the number of clustered poses; copies at most cap of each.  sampled6 (nullable, cap_sampled records)
// receives the sampled scene, *n_sampled its size.
ORC_API size_t orc_ppf_match(const void* h, const float* scene6, size_t n, double rel_sample_step, double rel_distance,
                             peb_ppf_pose* raw, size_t cap_raw, size_t* n_raw, peb_ppf_pose* results, size_t cap,
                             float* sampled6, size_t cap_sampled, size_t* n_sampled) {
  std::vector<float> s;
  std::vector<peb_ppf_pose> r, c;
  orc::ppf::match(*static_cast<const orc::ppf::Detector*>(h), scene6, n, rel_sample_step, rel_distance, s, r, c);
  if (n_raw) *n_raw = r.size();
  if (raw) std::memcpy(raw, r.data(), std::min(cap_raw, r.size()) * sizeof(peb_ppf_pose));
  if (results) std::memcpy(results, c.data(), std::min(cap, c.size()) * sizeof(peb_ppf_pose));
  if (n_sampled) *n_sampled = s.size() / 6;
  if (sampled6) std::memcpy(sampled6, s.data(), std::min(cap_sampled * 6, s.size()) * sizeof(float));
  return c.size();
}
// pieces for the tests
ORC_API void orc_ppf_feature(const double* p1, const double* n1, const double* p2, const double* n2, double* f4, double* alpha) {
  const orc::ppf::V3 a = {p1[0], p1[1], p1[2]}, an = {n1[0], n1[1], n1[2]}, b = {p2[0], p2[1], p2[2]}, bn = {n2[0], n2[1], n2[2]};
  orc::ppf::ppf_features(a, an, b, bn, f4);
  double R[9], t[3];
  orc::ppf::transform_rt(a, an, R, t);
  *alpha = orc::ppf::alpha_of(R, t, b);
}
ORC_API void orc_ppf_transform_rt(const double* p1, const double* n1, double* R9, double* t3) {
  orc::ppf::transform_rt({p1[0], p1[1], p1[2]}, {n1[0], n1[1], n1[2]}, R9, t3);
}
ORC_API size_t orc_ppf_sample(const float* pc6, size_t n, float step, float* out6, size_t cap) {
  float lo[3], hi[3];
  orc::ppf::bbox(pc6, n, 6, lo, hi);
  std::vector<float> s;
  orc::ppf::sample_by_quantization(pc6, n, 6, lo, hi, step, s);
  std::memcpy(out6, s.data(), std::min(cap * 6, s.size()) * sizeof(float));
  return s.size() / 6;
}
ORC_API size_t orc_ppf_cluster(const void* h, const peb_ppf_pose* poses, size_t n, peb_ppf_pose* out, size_t cap) {
  std::vector<peb_ppf_pose> in(poses, poses + n), res;
  orc::ppf::cluster_poses(*static_cast<const orc::ppf::Detector*>(h), in, res);
  std::memcpy(out, res.data(), std::min(cap, res.size()) * sizeof(peb_ppf_pose));
  return res.size();
}
}
