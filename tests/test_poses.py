"""Selection and packing of the refined pose: the Python mirror against scipy and
against the C++ functions of the facade (compiled with plain g++, no GPU)."""
import subprocess
from pathlib import Path

import numpy as np
from scipy.spatial.transform import Rotation

from pose_estimation_b200 import poses
from pose_estimation_b200.testing import synth

ROOT = Path(__file__).resolve().parents[1]


def test_quaternion_packing_matches_scipy():
    rng = np.random.default_rng(0)
    for _ in range(200):
        R = Rotation.random(random_state=rng.integers(1 << 31)).as_matrix()
        T = synth.make_pose(R, rng.normal(size=3))
        p = poses.pack_pose(T)
        q = Rotation.from_matrix(R).as_quat()  # x y z w
        q = -q if q[3] < 0 else q
        assert np.allclose(p[3:], q, atol=1e-6) and np.allclose(p[:3], T[:3, 3], atol=1e-6)
        ref = poses.pack_pose(T, reference_layout=True)
        assert ref[5] == 0.0 and np.allclose(ref[[0, 1, 2, 3, 4, 6]], p[[0, 1, 2, 3, 4, 6]])  # the reference loses qz


def test_selection_rule_of_the_reference():
    # most votes wins when the matcher returned <= 5 results
    assert poses.select_best_pose([100, 900, 500], [0.3, 0.2, 0.1], 3) == 1
    # > 5 results: lowest residual among > 400 votes overrides ...
    assert poses.select_best_pose([100, 900, 500, 50, 40, 30], [0.3, 0.2, 0.1, 0.0, 0.0, 0.0], 8) == 2
    # ... but a later pose with more votes takes over again (last assignment wins, as written)
    assert poses.select_best_pose([500, 900], [0.1, 0.2], 8) == 1
    assert poses.select_best_pose([], [], 0) == 0


def test_cpp_functions_agree_with_the_python_mirror(tmp_path):
    src = tmp_path / "poses_check.cpp"
    src.write_text(r'''
#include <cstdio>
#include "pe_b200/pcl_facade.hpp"
int main() {
  const int votes[6] = {100, 900, 500, 50, 40, 30};
  const double res[6] = {0.3, 0.2, 0.1, 0.0, 0.0, 0.0};
  std::printf("%zu %zu\n", pe_b200::select_best_pose(votes, res, 3, 3), pe_b200::select_best_pose(votes, res, 6, 8));
  float T[16];
  while (std::scanf("%f %f %f %f %f %f %f %f %f %f %f %f %f %f %f %f", T, T+1, T+2, T+3, T+4, T+5, T+6, T+7, T+8, T+9, T+10,
                    T+11, T+12, T+13, T+14, T+15) == 16) {
    float a[7], b[7];
    pe_b200::pack_pose(T, a, false);
    pe_b200::pack_pose(T, b, true);
    for (int i = 0; i < 7; ++i) std::printf("%.9g ", a[i]);
    for (int i = 0; i < 7; ++i) std::printf("%.9g ", b[i]);
    std::printf("\n");
  }
}
''')
    exe = tmp_path / "poses_check"
    r = subprocess.run(["g++", "-std=c++17", "-I", str(ROOT / "include"), str(src), "-o", str(exe), "-L",
                        str(ROOT / "pose_estimation_b200"), "-lpe_b200", f"-Wl,-rpath,{ROOT / 'pose_estimation_b200'}"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    rng = np.random.default_rng(1)
    Ts = [synth.make_pose(Rotation.random(random_state=int(rng.integers(1 << 31))).as_matrix(), rng.normal(size=3))
          for _ in range(50)]
    inp = "\n".join(" ".join(f"{v:.9g}" for v in np.asarray(T, np.float32).T.reshape(-1)) for T in Ts)
    out = subprocess.run([str(exe)], input=inp, capture_output=True, text=True).stdout.strip().splitlines()
    assert out[0].split() == ["1", "2"]
    for T, line in zip(Ts, out[1:]):
        vals = np.array([float(v) for v in line.split()], np.float32)
        T32 = np.asarray(T, np.float32)
        assert np.allclose(vals[:7], poses.pack_pose(T32), atol=1e-6)
        assert np.allclose(vals[7:], poses.pack_pose(T32, reference_layout=True), atol=1e-6)


def test_grasp_frame_of_the_manager():
    """pose_transformer.cpp:78-121 against an independent float64 construction with scipy."""
    rng = np.random.default_rng(3)
    for trial in range(200):
        R_cam = Rotation.random(random_state=int(rng.integers(1 << 31)))
        t_cam = rng.normal(size=3)
        q = R_cam.as_quat() * rng.uniform(0.5, 2.0)  # the reference normalises whatever it receives
        he = synth.make_pose(Rotation.random(random_state=int(rng.integers(1 << 31))).as_matrix(), rng.normal(size=3))
        out = poses.obj_in_base_frame(np.concatenate([t_cam, q]), he)
        base = he @ synth.make_pose(R_cam.as_matrix(), t_cam)
        y = base[:3, 1]
        zb = np.array([1.0, 0, 0]) if abs(y[2]) > 0.6 else np.array([0, 0, -1.0])
        if abs(abs(y[2]) - 0.6) < 1e-4:
            continue  # the float32 / float64 branch may differ exactly at the switch
        z = zb - (zb @ y) / (y @ y) * y
        x = np.cross(y, z)
        rot = np.stack([x / np.linalg.norm(x), y / np.linalg.norm(y), z / np.linalg.norm(z)], 1)
        assert np.allclose(out[:3], base[:3, 3], atol=2e-6 * max(1.0, np.abs(base[:3, 3]).max()))
        got = Rotation.from_quat(out[3:]).as_matrix()
        assert np.allclose(got, rot, atol=5e-6)
        assert abs(np.linalg.det(got) - 1) < 1e-5 and np.allclose(got[:, 1], y / np.linalg.norm(y), atol=5e-6)
        assert np.allclose(poses.hover_pose(np.concatenate([t_cam, q]), he) - out, [0, 0, 0.1, 0, 0, 0, 0])


def test_cpp_grasp_frame_agrees_with_the_python_mirror(tmp_path):
    src = tmp_path / "grasp_check.cpp"
    src.write_text(r"""
#include <cstdio>
#include "pe_b200/pcl_facade.hpp"
int main() {
  float in[23];
  for (;;) {
    for (int i = 0; i < 23; ++i) if (std::scanf("%f", in + i) != 1) return 0;
    double out[7];
    pe_b200::obj_in_base_frame(in, in + 7, out);
    for (int i = 0; i < 7; ++i) std::printf("%.9g ", out[i]);
    std::printf("\n");
  }
}
""")
    exe = tmp_path / "grasp_check"
    r = subprocess.run(["g++", "-std=c++17", "-I", str(ROOT / "include"), str(src), "-o", str(exe), "-L",
                        str(ROOT / "pose_estimation_b200"), "-lpe_b200", f"-Wl,-rpath,{ROOT / 'pose_estimation_b200'}"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    rng = np.random.default_rng(9)
    cases = []
    for _ in range(100):
        q = Rotation.random(random_state=int(rng.integers(1 << 31))).as_quat() * rng.uniform(0.5, 2.0)
        he = synth.make_pose(Rotation.random(random_state=int(rng.integers(1 << 31))).as_matrix(), rng.normal(size=3))
        cases.append((np.concatenate([rng.normal(size=3), q]).astype(np.float32), np.asarray(he, np.float32)))
    inp = "\n".join(" ".join(f"{v:.9g}" for v in np.concatenate([p, he.reshape(-1)])) for p, he in cases)
    out = subprocess.run([str(exe)], input=inp, capture_output=True, text=True).stdout.strip().splitlines()
    assert len(out) == len(cases)
    for (p, he), line in zip(cases, out):
        got = np.array([float(v) for v in line.split()])
        ref = poses.obj_in_base_frame(p, he)
        y2 = (he @ synth.make_pose(Rotation.from_quat(p[3:] / np.linalg.norm(p[3:])).as_matrix(), p[:3]))[2, 1]
        if abs(abs(y2) - 0.6) < 1e-3:
            continue  # the two float32 evaluation orders may take different branches exactly at the switch
        assert np.allclose(got[:3], ref[:3], atol=5e-6) and np.allclose(got[3:], ref[3:], atol=5e-6)
