// facade_check.cpp — exercises include/pe_b200/pcl_facade.hpp exactly the way host code written
// against PCL would (tests/test_cpp_facade.py compiles it with g++ and links libpe_b200.so).
//   facade_check <source.f32> <n_src> <target.f32> <n_tgt> <guess16.f32|-> <leaf> <k>
// Clouds are raw little-endian float32 x y z 1 records.  Prints one line per stage.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "pe_b200/pcl_facade.hpp"

static std::vector<float> read_f32(const char* path, size_t count) {
  std::vector<float> v(count);
  FILE* f = std::fopen(path, "rb");
  if (!f || std::fread(v.data(), sizeof(float), count, f) != count) {
    std::fprintf(stderr, "cannot read %zu floats from %s\n", count, path);
    std::exit(2);
  }
  std::fclose(f);
  return v;
}

static void print_T(const char* tag, const peb_icp_result& r) {
  std::printf("%s iterations %d state %d converged %d fitness %.17g T", tag, r.iterations, r.state, r.converged, r.fitness);
  for (int i = 0; i < 16; ++i) std::printf(" %.9g", r.T[i]);
  std::printf("\n");
}

int main(int argc, char** argv) {
  if (argc != 8) {
    std::fprintf(stderr, "usage: %s source.f32 n_src target.f32 n_tgt guess16.f32|- leaf k\n", argv[0]);
    return 2;
  }
  const size_t n_src = std::strtoull(argv[2], nullptr, 10), n_tgt = std::strtoull(argv[4], nullptr, 10);
  const float leaf = std::strtof(argv[6], nullptr);
  const int k = std::atoi(argv[7]);
  try {
    pe_b200::Context ctx(0);
    std::vector<float> src = read_f32(argv[1], 4 * n_src), tgt = read_f32(argv[3], 4 * n_tgt), guess;
    const bool have_guess = argv[5][0] != '-' || argv[5][1] != 0;
    if (have_guess) guess = read_f32(argv[5], 16);

    pe_b200::VoxelGrid vg(ctx);
    vg.setInputCloud(tgt.data(), n_tgt, 16);
    vg.setLeafSize(leaf, leaf, leaf);
    std::vector<float> ds;
    vg.filter(ds);
    std::printf("voxel_grid %zu -> %zu\n", n_tgt, ds.size() / 4);

    pe_b200::SACSegmentation seg(ctx);
    seg.setModelType(pe_b200::SACSegmentation::SACMODEL_PLANE);
    seg.setMethodType(pe_b200::SACSegmentation::SAC_RANSAC);
    seg.setOptimizeCoefficients(true);
    seg.setDistanceThreshold(0.001);
    seg.setMaxIterations(100);
    seg.setInputCloud(tgt.data(), n_tgt, 16);
    std::vector<int32_t> inliers;
    std::vector<float> coeff;
    seg.segment(inliers, coeff);
    std::printf("sac_plane inliers %zu iterations %d coeff", inliers.size(), seg.iterations());
    for (float v : coeff) std::printf(" %.9g", v);
    std::printf("\n");

    pe_b200::NormalEstimation ne(ctx);
    ne.setInputCloud(ds.data(), ds.size() / 4, 16);
    ne.setKSearch(k);
    std::vector<float> normals;
    ne.compute(normals);
    std::printf("normals %zu first %.9g %.9g %.9g curvature %.9g\n", normals.size() / 8, normals[0], normals[1], normals[2],
                normals[4]);

    pe_b200::IterativeClosestPoint icp(ctx);
    icp.setInputSource(src.data(), n_src, 16);
    icp.setInputTarget(ds.data(), ds.size() / 4, 16);
    icp.setMaximumIterations(30);
    icp.getConvergeCriteria()->setAbsoluteMSE(-1.0);
    std::vector<float> aligned;
    icp.align(&aligned, have_guess ? guess.data() : nullptr);
    print_T("icp_p2p", icp.result());

    pe_b200::IterativeClosestPointWithNormals icpn(ctx);
    icpn.setInputSource(src.data(), n_src, 16);
    icpn.setInputTarget(ds.data(), ds.size() / 4, 16, normals.data(), 32);
    icpn.setMaximumIterations(30);
    icpn.getConvergeCriteria()->setAbsoluteMSE(-1.0);
    icpn.align(nullptr, have_guess ? guess.data() : nullptr);
    print_T("icp_p2plane", icpn.result());

    // batch of two poses: the guess (or identity) and the p2p result
    std::vector<float> poses(32, 0.0f);
    for (int i = 0; i < 4; ++i) poses[5 * i] = 1.0f;
    if (have_guess) std::copy(guess.begin(), guess.end(), poses.begin());
    std::copy(icp.getFinalTransformation(), icp.getFinalTransformation() + 16, poses.begin() + 16);
    std::vector<peb_icp_result> res;
    icp.alignBatch(poses.data(), 2, res);
    print_T("batch0", res[0]);
    print_T("batch1", res[1]);

    // the same batch sharded over two contexts on this device (peb_multi_*): must reproduce batch0 / batch1
    {
      pe_b200::MultiContext many(std::vector<int>{0, 0});
      pe_b200::MultiDeviceICP micp(many, icp);
      micp.setInputSource(src.data(), n_src, 16);
      micp.setInputTarget(ds.data(), ds.size() / 4, 16);
      std::vector<peb_icp_result> mres;
      micp.alignBatch(poses.data(), 2, mres);
      std::printf("multi_batch contexts %d identical %d\n", many.size(),
                  static_cast<int>(std::memcmp(mres.data(), res.data(), 2 * sizeof(peb_icp_result)) == 0));
    }

    // cv::ppf_match_3d::ICP as the reference calls it: model / scene with the normals computed above, two poses
    {
      pe_b200::NormalEstimation nes(ctx);
      nes.setInputCloud(src.data(), n_src, 16);
      nes.setKSearch(k);
      std::vector<float> src_normals;
      nes.compute(src_normals);
      auto pack6 = [](const std::vector<float>& xyz4, const std::vector<float>& nrm8, std::vector<float>& out) {
        out.clear();
        for (size_t i = 0; i < xyz4.size() / 4; ++i) {
          const float* p = &xyz4[4 * i];
          const float* q = &nrm8[8 * i];
          if (!(q[0] == q[0])) continue;  // NaN normal
          out.insert(out.end(), {p[0], p[1], p[2], q[0], q[1], q[2]});
        }
      };
      std::vector<float> model6, scene6;
      pack6(src, src_normals, model6);
      pack6(ds, normals, scene6);
      std::vector<double> cvposes(32, 0.0), residuals(2, 0.0);
      for (int p = 0; p < 2; ++p)
        for (int r = 0; r < 4; ++r)
          for (int c = 0; c < 4; ++c) cvposes[16 * p + 4 * r + c] = poses[16 * p + 4 * c + r];  // column-major -> row-major
      pe_b200::CvIcp cvicp(ctx, 250, 0.005f, 2.5f, 8);
      cvicp.registerModelToScene(model6.data(), model6.size() / 6, scene6.data(), scene6.size() / 6, cvposes.data(), 2,
                                 residuals.data());
      std::printf("cvicp residuals %.17g %.17g pose0", residuals[0], residuals[1]);
      for (int i = 0; i < 16; ++i) std::printf(" %.17g", cvposes[i]);
      std::printf("\n");
      // the coarse matcher with the reference's parameters shape (opencv_surface_match.cpp:45-46, :65)
      pe_b200::PPF3DDetector det(ctx, 0.08, 0.08, 30);
      det.trainModel(model6.data(), model6.size() / 6);
      std::vector<peb_ppf_pose> found;
      det.match(scene6.data(), scene6.size() / 6, found, 0.5, 0.08);
      std::printf("ppf clusters %zu votes %llu pose0", found.size(), found.empty() ? 0ull : (unsigned long long)found[0].num_votes);
      for (int i = 0; i < 16 && !found.empty(); ++i) std::printf(" %.17g", found[0].pose[i]);
      std::printf("\n");
    }

    // PCL options without a CUDA path must be refused, not emulated
    try {
      icp.setUseReciprocalCorrespondences(true);
      std::printf("unsupported-option check FAILED\n");
      return 1;
    } catch (const pe_b200::Error& e) {
      std::printf("unsupported-option refused: code %d\n", e.code);
    }
  } catch (const pe_b200::Error& e) {
    std::fprintf(stderr, "pe_b200 error %d: %s\n", e.code, e.what());
    return 3;
  }
  return 0;
}
