"""The C++ facade (include/pe_b200/pcl_facade.hpp): compiles and links against libpe_b200.so with
plain g++ (CPU), fails loudly without a GPU, and on the B200 reproduces the Python host classes
bit for bit (both are thin layers over the same C ABI)."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

from pose_estimation_b200.testing import synth

ROOT = Path(__file__).resolve().parents[1]
EXE = ROOT / "tests" / "cpp" / "facade_check"


@pytest.fixture(scope="module")
def exe():
    subprocess.run(["make", "-C", str(ROOT / "pose_estimation_b200" / "csrc")], check=True, capture_output=True)
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-I", str(ROOT / "include"), str(EXE) + ".cpp", "-o", str(EXE),
           "-L", str(ROOT / "pose_estimation_b200"), "-lpe_b200", f"-Wl,-rpath,{ROOT / 'pose_estimation_b200'}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return EXE


def _write_case(tmp_path):
    prob = synth.make_c1(3000, seed=5)
    (tmp_path / "src.f32").write_bytes(np.ascontiguousarray(prob.source, np.float32).tobytes())
    (tmp_path / "tgt.f32").write_bytes(np.ascontiguousarray(prob.target, np.float32).tobytes())
    return prob


def test_facade_builds_with_plain_gxx_and_fails_loudly_without_gpu(exe, tmp_path):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    prob = _write_case(tmp_path)
    r = subprocess.run([str(exe), str(tmp_path / "src.f32"), str(len(prob.source)), str(tmp_path / "tgt.f32"),
                        str(len(prob.target)), "-", "0.002", "12"], capture_output=True, text=True)
    assert r.returncode == 3
    assert "no CUDA device" in r.stderr and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_facade_matches_python_host_classes(exe, tmp_path):
    from pose_estimation_b200 import pcl

    prob = _write_case(tmp_path)
    r = subprocess.run([str(exe), str(tmp_path / "src.f32"), str(len(prob.source)), str(tmp_path / "tgt.f32"),
                        str(len(prob.target)), "-", "0.002", "12"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lines = {ln.split()[0]: ln.split() for ln in r.stdout.strip().splitlines()}
    ctx = pcl.Context(0)
    vg = pcl.VoxelGrid(ctx)
    vg.setInputCloud(prob.target)
    vg.setLeafSize(0.002)
    ds = vg.filter()
    assert int(lines["voxel_grid"][3]) == len(ds)
    seg = pcl.SACSegmentation(ctx)
    seg.setModelType(pcl.SACSegmentation.SACMODEL_PLANE)
    seg.setMethodType(pcl.SACSegmentation.SAC_RANSAC)
    seg.setDistanceThreshold(0.001)
    seg.setMaxIterations(100)
    seg.setInputCloud(prob.target)
    inl, coeff = seg.segment()
    assert int(lines["sac_plane"][2]) == len(inl) and int(lines["sac_plane"][4]) == seg.iterations_
    assert np.array_equal(np.array([float(v) for v in lines["sac_plane"][6:10]], np.float32), coeff)
    ne = pcl.NormalEstimation(ctx)
    ne.setInputCloud(ds)
    ne.setKSearch(12)
    nrm = ne.compute()
    assert np.allclose([float(v) for v in lines["normals"][3:6]], nrm[0, :3], atol=1e-7)

    def parse(tag):
        t = lines[tag]
        return int(t[2]), int(t[4]), float(t[8]), np.array([float(v) for v in t[10:26]], np.float32).reshape(4, 4).T

    for tag, cls, normals in (("icp_p2p", pcl.IterativeClosestPoint, None),
                              ("icp_p2plane", pcl.IterativeClosestPointWithNormals, nrm)):
        icp = cls(ctx)
        icp.setInputSource(prob.source)
        icp.setInputTarget(ds, normals)
        icp.setMaximumIterations(30)
        icp.getConvergeCriteria().setAbsoluteMSE(-1.0)
        icp.align(want_output=False)
        it, state, fit, T = parse(tag)
        assert it == icp.nr_iterations_ and state == icp.result.state
        assert np.array_equal(T, icp.getFinalTransformation())
        assert fit == icp.getFitnessScore()
        if tag == "icp_p2p":
            T_p2p = T
    it0, _, fit0, T0 = parse("batch0")
    it1, _, fit1, _ = parse("batch1")
    _, _, fit, T = parse("icp_p2p")
    assert np.array_equal(T0, T) and fit0 == fit and it0 == 30 and it1 == 30 and fit1 <= fit0 * 1.01
    assert "unsupported-option refused: code -6" in r.stdout
    assert lines["multi_batch"][2] == "2" and lines["multi_batch"][4] == "1"  # two contexts, records identical to alignBatch
    # the cv ICP of the facade against the Python mirror (same library underneath: bit-identical)
    nes = pcl.NormalEstimation(ctx)
    nes.setInputCloud(prob.source)
    nes.setKSearch(12)
    sn = nes.compute()
    ok_m, ok_s = np.isfinite(sn[:, 0]), np.isfinite(nrm[:, 0])
    model6 = np.concatenate([prob.source[ok_m, :3], sn[ok_m, :3]], 1).astype(np.float32)
    scene6 = np.concatenate([ds[ok_s, :3], nrm[ok_s, :3]], 1).astype(np.float32)
    poses2 = np.stack([np.eye(4), np.asarray(T_p2p, np.float64)])
    got, res = pcl.CvIcp(250, 0.005, 2.5, 8, ctx=ctx).registerModelToScene(model6, scene6, poses2)
    cv = [float(v) for v in lines["cvicp"][2:4]] , [float(v) for v in lines["cvicp"][5:21]]
    assert np.array_equal(np.array(cv[0]), res) and np.array_equal(np.array(cv[1]).reshape(4, 4), got[0])
    # PPF3DDetector of the facade against the Python mirror (same library underneath: identical clusters)
    det = pcl.PPF3DDetector(0.08, 0.08, 30, ctx=ctx)
    det.trainModel(model6)
    found = det.match(scene6, 0.5, 0.08)
    assert int(lines["ppf"][2]) == len(found) and int(lines["ppf"][4]) == found[0].num_votes
    assert np.array_equal(np.array([float(v) for v in lines["ppf"][6:22]]).reshape(4, 4), found[0].matrix)
    det.close()
    ctx.close()
