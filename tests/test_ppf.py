"""cv::ppf_match_3d::PPF3DDetector (pose_estimation/src/opencv_surface_match.cpp:37-51, :65).

CPU part: the oracle's restatement (oracle/ppf_oracle.cpp, parity unpinned — opencv_contrib is not in the image) against
independent numpy restatements of its pieces, and against what a matcher must do.  GPU part: the CUDA path
(csrc/ppf.cu through the C ABI) against the oracle on the same clouds."""
import math

import numpy as np
import pytest

import oracle
from pose_estimation_b200.testing import synth


@pytest.fixture(scope="module")
def orc():
    return oracle.Oracle()


def make_case(orc, seed=3, width=486, height=300, n_model=20000, leaf=0.002):
    """model (n x 6, object frame), scene (m x 6: rendered view with clutter, background plane cut off, voxel-down-sampled,
    k = 20 normals towards the camera like computeNormalsPC3d(.., 20, true, (0,0,0))), ground-truth pose."""
    rng = np.random.default_rng(seed)
    surf = synth.Surface(1)
    gt = synth.default_gt_pose(rng)
    mp, mn = surf.sample(n_model, rng)
    model6 = np.concatenate([mp, mn], 1).astype(np.float32)
    scene = synth.render_scene(surf, gt, rng, width, height)
    fin = scene[np.isfinite(scene).all(1)]
    keep = fin[fin[:, 2] < 0.735]
    ds = orc.voxel_grid(keep, leaf)[0]
    nrm = orc.normals(ds, 20)
    ok = np.isfinite(nrm[:, :3]).all(1)
    scene6 = np.concatenate([ds[ok, :3], nrm[ok, :3]], 1).astype(np.float32)
    return model6, scene6, gt


@pytest.fixture(scope="module")
def case(orc):
    return make_case(orc)


# ---- independent numpy restatements ---------------------------------------------------------------------------------
def np_sample(pc6, step):
    """samplePCByQuantization as a dict of cells (float32 cell arithmetic, float64 means)."""
    pc6 = np.asarray(pc6, np.float32)
    lo = pc6[:, :3].min(0)
    rng_ = pc6[:, :3].max(0) - lo
    nsd = int(1.0 / float(np.float32(step)))  # (double division of the float step, like upstream: 0.1f -> 9)
    cells = {}
    idx = (np.float32(nsd) * (pc6[:, :3] - lo) / rng_).astype(np.int32)
    key = idx[:, 0] * nsd * nsd + idx[:, 1] * nsd + idx[:, 2]
    for i, k in enumerate(key):
        cells.setdefault(int(k), []).append(i)
    out = []
    for k in sorted(cells):
        rows = pc6[cells[k]].astype(np.float64)
        s = np.zeros(6)
        for r in rows:  # sequential, input order
            s += r
        s /= len(rows)
        n = s[3:]
        nn = math.sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2])
        n = n / nn if nn > 1.192092896e-07 else np.zeros(3)
        out.append(np.concatenate([s[:3], n]).astype(np.float32))
    return np.array(out, np.float32)


def np_frame(p1, n1):
    """The frame of computeTransformRT by its defining properties, built differently: Rodrigues about n1 x e_x."""
    n1 = np.asarray(n1, np.float64)
    ex = np.array([1.0, 0.0, 0.0])
    axis = np.cross(n1, ex)
    s = np.linalg.norm(axis)
    angle = math.acos(n1[0])
    if s == 0:
        axis = np.array([0.0, 1.0, 0.0])
    else:
        axis = axis / s
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    R = np.eye(3) + math.sin(angle) * K + (1 - math.cos(angle)) * (K @ K)
    return R, -R @ np.asarray(p1, np.float64)


def np_features(P, N, i):
    """f[4] and alpha of the ordered pairs (i, j) for every j, vectorised (float64)."""
    p1, n1 = P[i], N[i]
    d = P - p1
    f3 = np.linalg.norm(d, axis=1)
    with np.errstate(invalid="ignore", divide="ignore"):
        dn = d / f3[:, None]
        f0 = np.arccos(dn @ n1)
        f1 = np.arccos(np.einsum("ij,ij->i", N, dn))
        f2 = np.arccos(N @ n1)
    R, t = np_frame(p1, n1)
    m = P @ R.T + t
    alpha = np.arctan2(-m[:, 2], m[:, 1])
    return np.stack([f0, f1, f2, f3], 1), alpha


def test_sampling_matches_numpy_restatement(orc, case):
    model6, scene6, _ = case
    for pc, step in ((model6, 0.03), (scene6, 0.03), (model6[:3000], 0.1)):
        got = orc.ppf_sample(pc, step)
        ref = np_sample(pc, step)
        assert got.shape == ref.shape
        assert np.array_equal(got, ref)


def test_frame_and_feature_match_numpy(orc, case):
    model6, _, _ = case
    s = orc.ppf_sample(model6, 0.05)
    P, N = s[:, :3].astype(np.float64), s[:, 3:].astype(np.float64)
    for i in (0, 7, len(s) // 2):
        R, t = orc.ppf_transform_rt(P[i], N[i])
        Rn, tn = np_frame(P[i], N[i])
        assert np.allclose(R, Rn, atol=1e-12) and np.allclose(t, tn, atol=1e-12)
        assert np.allclose(R @ N[i], [np.linalg.norm(N[i]), 0, 0], atol=1e-7)  # (normals are unit to float precision)
        F, A = np_features(P, N, i)
        for j in (1, 5, len(s) - 1):
            if j == i:
                continue
            f, a = orc.ppf_feature(P[i], N[i], P[j], N[j])
            assert np.allclose(f, F[j], atol=1e-12)
            assert abs(a - A[j]) < 1e-12 or abs(abs(a - A[j]) - 2 * math.pi) < 1e-12


def np_vote(model_s, scene_s, i_ref, angle_step, dist_step, num_angles):
    """The voting of one scene reference point against an exact-key table, numpy + a dict."""
    Pm, Nm = model_s[:, :3].astype(np.float64), model_s[:, 3:].astype(np.float64)
    table = {}
    for i in range(len(model_s)):
        F, A = np_features(Pm, Nm, i)
        K = np.stack([F[:, 0] / angle_step, F[:, 1] / angle_step, F[:, 2] / angle_step, F[:, 3] / dist_step], 1)
        A32 = A.astype(np.float32)
        for j in range(len(model_s)):
            if j == i or not np.isfinite(K[j]).all():
                continue
            table.setdefault(tuple(K[j].astype(np.int64)), []).append((i, A32[j]))
    Ps, Ns = scene_s[:, :3].astype(np.float64), scene_s[:, 3:].astype(np.float64)
    F, A = np_features(Ps, Ns, i_ref)
    K = np.stack([F[:, 0] / angle_step, F[:, 1] / angle_step, F[:, 2] / angle_step, F[:, 3] / dist_step], 1)
    acc = np.zeros((len(model_s), num_angles), np.int64)
    for j in range(len(scene_s)):
        if j == i_ref or not np.isfinite(K[j]).all():
            continue
        for (mi, am) in table.get(tuple(K[j].astype(np.int64)), ()):
            alpha = float(am) - A[j]
            acc[mi, min(int(num_angles * (alpha + 2 * math.pi) / (4 * math.pi)), num_angles - 1)] += 1
    return acc


def test_votes_of_reference_points_match_numpy(orc):
    model6, scene6, _ = make_case(orc, seed=5, width=243, height=150, n_model=4000, leaf=0.004)
    prm = oracle.ppf_params(0.08, 0.08, 30)
    det = orc.ppf_train(model6, prm)
    ms = det.sampled()
    assert 40 < len(ms) < 400
    res, raw, ss = det.match(scene6, 1.0, 0.08)
    angle_step = (360.0 / 30) * math.pi / 180.0
    num_angles = int(math.floor(2 * math.pi / angle_step))
    for i_ref in (0, len(ss) // 3, len(ss) - 1):
        acc = np_vote(ms, ss, i_ref, angle_step, det.distance_step, num_angles)
        flat = int(acc.argmax())  # first maximum in (model reference, alpha bin) order, like upstream's scan
        assert raw[i_ref].num_votes == acc.max()
        if acc.max() > 0:
            assert (raw[i_ref].model_index, round((raw[i_ref].alpha + 2 * math.pi) * num_angles / (4 * math.pi))) == divmod(flat, num_angles)


def py_cluster(poses, pos_thr, rot_thr):
    order = sorted(range(len(poses)), key=lambda i: -poses[i].num_votes)
    clusters = []
    for i in order:
        p = poses[i]
        for c in clusters:
            q = poses[c[0]]
            if abs(p.angle - q.angle) < rot_thr and np.linalg.norm(np.array(q.t) - np.array(p.t)) < pos_thr:
                c.append(i)
                break
        else:
            clusters.append([i])
    clusters.sort(key=lambda c: -sum(poses[i].num_votes for i in c))
    return clusters


def test_clustering_matches_python_restatement(orc, case):
    model6, scene6, _ = case
    det = orc.ppf_train(model6)
    res, raw, _ = det.match(scene6, 1.0, 0.03)
    angle_step = (360.0 / 40) * math.pi / 180.0
    clusters = py_cluster(raw, 0.03, (360 / angle_step) / 180.0 * math.pi)
    assert len(clusters) == len(res)
    for c, r in zip(clusters, res):
        assert r.num_votes == sum(raw[i].num_votes for i in c)
        t = np.mean([raw[i].t for i in c], 0)
        assert np.allclose(r.t, t, atol=1e-12)
        q = np.mean([raw[i].q for i in c], 0)
        assert np.allclose(r.q, q, atol=1e-12)
    again = det.cluster(raw)
    assert [p.num_votes for p in again] == [p.num_votes for p in res]


def test_matcher_finds_the_object_and_the_icp_finishes_it(orc, case):
    """What the reference's find_object_in_scene does with the detector's output (opencv_surface_match.cpp:65-94): the
    first poses go through cv::ppf_match_3d::ICP.  The coarse pose must land in the object's basin (alpha bins are
    4 pi / 40 wide and the bin's lower edge is used: ~0.1 rad of systematic rotation error is upstream's), the refined
    one on the ground truth."""
    model6, scene6, gt = case
    det = orc.ppf_train(model6)
    res, raw, ss = det.match(scene6, 1.0, 0.03)
    assert len(raw) == len(ss)
    top = res[0]
    rot, trans = synth.pose_error(top.matrix, gt)
    assert trans < 0.01 and rot < 0.35
    assert top.num_votes > 5 * res[1].num_votes
    P, resid = orc.cvicp_register(model6, scene6, np.array([top.matrix]), oracle.cvicp_params(250, 0.005, 2.5, 8))
    rot2, trans2 = synth.pose_error(P[0], gt)
    assert rot2 < 0.01 and trans2 < 0.002


# ---- the CUDA path ----------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_gpu_sampling_training_and_raw_poses_match_oracle(orc, case):
    from pose_estimation_b200 import pcl

    model6, scene6, gt = case
    det = pcl.PPF3DDetector(0.03, 0.03, 40)
    det.trainModel(model6)
    ref = orc.ppf_train(model6)
    assert np.array_equal(det.sampled_model(), ref.sampled())  # bit-exact sampling (sequential double means)
    res, raw = det.match(scene6, 1.0, 0.03, return_raw=True)
    ores, oraw, _ = ref.match(scene6, 1.0, 0.03)
    assert len(raw) == len(oraw)
    same = [(a.num_votes, a.model_index) == (b.num_votes, b.model_index) and abs(a.alpha - b.alpha) < 1e-12 for a, b in zip(raw, oraw)]
    # a feature within an ulp of a bin edge may quantise differently under CUDA's acos / atan2 (<= 2 ulp from glibc's)
    assert np.mean(same) >= 0.995, f"{len(same) - sum(same)} of {len(same)} reference points differ"
    for a, b, s in zip(raw, oraw, same):
        if s:
            assert np.allclose(a.matrix, b.matrix, atol=1e-9)
            assert np.allclose(a.q, b.q, atol=1e-9) and abs(a.angle - b.angle) < 1e-7
    if all(same):
        assert len(res) == len(ores)
        for a, b in zip(res, ores):
            assert a.num_votes == b.num_votes and np.allclose(a.matrix, b.matrix, atol=1e-9)
    rot, trans = synth.pose_error(res[0].matrix, gt)
    assert trans < 0.01 and rot < 0.35
    det.close()


@pytest.mark.gpu
def test_gpu_match_with_scene_sample_step_and_other_parameters(orc):
    from pose_estimation_b200 import pcl

    model6, scene6, _ = make_case(orc, seed=9, width=243, height=150, n_model=6000, leaf=0.004)
    for (s, d, a, step, dist) in ((0.05, 0.05, 30, 1.0 / 5.0, 0.05), (0.08, 0.08, 24, 0.5, 0.04)):
        det = pcl.PPF3DDetector(s, d, a)
        det.trainModel(model6)
        ref = orc.ppf_train(model6, oracle.ppf_params(s, d, a))
        assert np.array_equal(det.sampled_model(), ref.sampled())
        res, raw = det.match(scene6, step, dist, return_raw=True)
        ores, oraw, _ = ref.match(scene6, step, dist)
        assert len(raw) == len(oraw)
        same = [(x.num_votes, x.model_index) == (y.num_votes, y.model_index) and abs(x.alpha - y.alpha) < 1e-12 for x, y in zip(raw, oraw)]
        assert np.mean(same) >= 0.99
        det.close()


@pytest.mark.gpu
def test_gpu_ppf_errors_and_edge_cases(orc):
    from pose_estimation_b200 import pcl

    det = pcl.PPF3DDetector(0.03, 0.03, 40)
    with pytest.raises(pcl.PebError):
        det.match(np.zeros((10, 6), np.float32))  # not trained
    with pytest.raises(pcl.PebError):
        det.trainModel(np.zeros((10, 4), np.float32))  # wrong shape
    with pytest.raises(pcl.PebError):
        det.trainModel(np.zeros((10, 6), np.float32))  # degenerate: one sampled point
    rng = np.random.default_rng(0)
    surf = synth.Surface(1)
    mp, mn = surf.sample(3000, rng)
    model6 = np.concatenate([mp, mn], 1).astype(np.float32)
    model6[5, :3] = np.nan  # non-finite rows are skipped by the sampler
    det.trainModel(model6)
    ref = orc.ppf_train(model6)
    assert np.array_equal(det.sampled_model(), ref.sampled())
    assert det.match(np.full((4, 6), np.nan, np.float32)) == []  # nothing finite: no poses
    det.close()
