"""The C++ oracle (oracle/pcl_oracle.cpp) against a SECOND, independently written restatement (tests/witness.py: numpy
+ a real FLANN kd-tree from cv2) of the same PCL 1.10 functions, function by function (real PCL, Eigen,
FLANN headers and cv2.ppf_match_3d are absent from the CPU image AND from the GPU image — tools/probe_image.sh,
profiles/r2_probe_image.txt — so no oracle/_ref can exist).

Lock-step mode: the witness computes its own increment in every iteration but MOVES by the oracle's, so both walk the
same trajectory and every piece is compared in isolation:
  correspondences, thresholds, MSE, counts, state, iteration count, final transform chain   -> exactly equal
  the estimator's increment (different SVD / solve algorithm)                                 -> to float rounding
Free-running mode: both run on their own; final poses within the north star's 1e-5 rad / 1e-5 m.
"""
import numpy as np
import pytest

import witness as W
from oracle import DBL_MAX, default_params
from pose_estimation_b200.testing import synth
from util import pose_delta

INC_TOL = 1.5e-6  # a float32 Jacobi SVD (oracle, as Eigen) against a float64 LAPACK SVD (witness) of the same float sigma: a dozen float eps


def _run_pair(oracle, target, source, guess, normals=None, **kw):
    prm = default_params(**kw)
    o = oracle.icp(target, normals).align(source, guess, prm, trace_cap=max(prm.max_iterations, 1))
    wkw = {k: v for k, v in kw.items()}
    icp = W.ICP(target, normals)
    forced = icp.align(source, guess, forced_increments=list(o["trace_T"]), **wkw)
    free = icp.align(source, guess, **wkw)
    return o, forced, free, icp


def _check_lockstep(o, w):
    r = o["result"]
    assert (w["iterations"], w["state"], int(w["converged"])) == (r.iterations, r.state, r.converged)
    assert w["n_correspondences"] == r.n_correspondences
    k = len(o["trace_T"])
    assert len(w["trace"]["inc"]) == k == len(w["trace"]["mse"])
    assert np.array_equal(np.asarray(w["trace"]["mse"]), o["trace_mse"])  # double sums of the same float distances, same order
    for a, b in zip(w["trace"]["inc"], o["trace_T"]):
        assert np.abs(a - b).max() < INC_TOL
    if r.state != W.NO_CORRESPONDENCES or k:
        assert np.array_equal(w["T"], r.matrix())  # final = inc_k * ... * inc_1 * guess, float products in the same order
    assert w["last_mse"] == r.last_mse
    # the correspondences of the last iteration, index for index
    q, m = w["trace"]["match"][-1]
    ref = o["corr_idx"]
    got = np.full(len(ref), -1, np.int32)
    got[q] = m
    assert np.array_equal(got, ref)


@pytest.fixture(scope="module")
def c1():
    return synth.make_c1(3000, seed=21)


@pytest.mark.parametrize("kw", [
    dict(),                                                          # PCL defaults: 10 iterations, |dMSE| < 1e-12 may fire
    dict(max_iterations=40, transformation_epsilon=1e-10),           # -> TRANSFORM
    dict(max_iterations=40, euclidean_fitness_epsilon=1e-4),         # -> REL_MSE
    dict(max_iterations=40, abs_mse_threshold=1e-9),                 # -> ABS_MSE
    dict(max_iterations=60, transformation_epsilon=1e-10, max_iterations_similar=3),  # the similar-transforms gate
    dict(max_iterations=25, abs_mse_threshold=-1.0),                 # fixed count (the benchmark's setting)
    dict(max_iterations=25, max_corr_dist=0.004),                    # distance threshold drops pairs
    dict(max_iterations=25, max_corr_dist=0.02, rejector_max_dist=0.003),  # + the rejector's strict <
    dict(max_iterations=0),                                          # the loop body still runs once
    dict(max_iterations=5, max_corr_dist=1e-5),                      # too few correspondences
])
def test_icp_loop_and_criteria_in_lock_step(oracle, c1, kw):
    o, forced, free, _ = _run_pair(oracle, c1.target, c1.source, None, **kw)
    _check_lockstep(o, forced)
    r = o["result"]
    if r.state != W.NO_CORRESPONDENCES:
        rot, tr = pose_delta(free["T"], r.matrix())
        # PCL's float umeyama is order / algorithm noisy at the 3e-5 rad level; a stop criterion
        # that fires on a creeping tail can fire some iterations apart, and the two runs then stop at different points
        # of the same slowly converging sequence
        if free["iterations"] == r.iterations:
            assert rot < 1.5e-4 and tr < 5e-5
        else:
            assert rot < 1e-3 and tr < 2e-4 and free["state"] == r.state


def test_icp_states_reached(oracle, c1):
    """the parametrisation above really exercises every exit of hasConverged"""
    seen = set()
    for kw in (dict(max_iterations=40, transformation_epsilon=1e-10), dict(max_iterations=40, euclidean_fitness_epsilon=1e-4),
               dict(max_iterations=40, abs_mse_threshold=1e-9), dict(max_iterations=3), dict(max_iterations=5, max_corr_dist=1e-5)):
        seen.add(oracle.icp(c1.target).align(c1.source, None, default_params(**kw))["result"].state)
    assert seen == {W.TRANSFORM, W.REL_MSE, W.ABS_MSE, W.ITERATIONS, W.NO_CORRESPONDENCES}


def test_icp_with_guess_nonfinite_points_and_double_sums(oracle, c1):
    src = c1.source.copy()
    src[::97, 0] = np.nan
    src[5, 2] = np.inf
    tgt = c1.target.copy()
    tgt[::131, 1] = np.nan
    rng = np.random.default_rng(4)
    guess = synth.perturb_pose(np.eye(4), rng, 2.0, 0.002).astype(np.float32)
    o, forced, _, icp = _run_pair(oracle, tgt, src, guess, max_iterations=15, abs_mse_threshold=-1.0)
    _check_lockstep(o, forced)
    # wide_accum (the oracle's double-sum mode, what the GPU is held to at 1e-5) against the witness in float64
    prm = default_params(max_iterations=15, abs_mse_threshold=-1.0)
    ow = oracle.icp(tgt, wide_accum=True).align(src, guess, prm, trace_cap=15)
    ww = icp.align(src, guess, max_iterations=15, abs_mse_threshold=-1.0, wide=True)
    rot, tr = pose_delta(ww["T"], ow["result"].matrix())
    assert rot < 2e-6 and tr < 2e-6
    for a, b in zip(ww["trace"]["inc"][:3], ow["trace_T"][:3]):
        assert np.abs(a - b).max() < 2e-7
    # getFitnessScore incl. the squared-vs-unsquared max_range quirk
    T = ow["result"].matrix()
    for max_range in (DBL_MAX, 1e-6, 2e-7):
        f_o, _ = oracle.icp(tgt).fitness(src, T, max_range)
        f_w = icp.fitness(src, T, max_range)
        assert f_o == f_w


@pytest.fixture(scope="module")
def scene(oracle):
    return synth.make_c2(scale=0.2, downsample=lambda p, leaf: oracle.voxel_grid(p, leaf)[0])


def test_point_to_plane_icp_in_lock_step(oracle, scene):
    p = scene
    nrm = oracle.normals(p.target, 12)
    nrm[::211, :3] = np.nan  # pairs with a non-finite normal stay out of the normal equations but count as correspondences
    src = p.source[:4000]
    o, forced, free, _ = _run_pair(oracle, p.target, src, p.guess.astype(np.float32), normals=nrm[:, :3].copy(),
                                   max_iterations=12, abs_mse_threshold=-1.0, estimator=1, max_corr_dist=0.02)
    _check_lockstep(o, forced)
    rot, tr = pose_delta(free["T"], o["result"].matrix())
    assert rot < 1e-5 and tr < 1e-5


def test_voxel_grid_witness(oracle, scene):
    p = scene
    raw = p.organized[::7]
    for leaf, min_pts in ((0.004, 0), (0.006, 3)):
        got, unchanged = oracle.voxel_grid(raw, leaf, min_pts)
        ref, unchanged_w = W.voxel_grid(raw, leaf, min_pts)
        assert unchanged == unchanged_w
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    got, unchanged = oracle.voxel_grid(raw, 1e-5)
    ref, unchanged_w = W.voxel_grid(raw, 1e-5)
    assert unchanged and unchanged_w and np.array_equal(got[:, :3], ref[:, :3], equal_nan=True)


def test_normals_witness(oracle, scene):
    pts = scene.target[:6000].copy()
    pts[17] = np.nan
    for k in (8, 30):
        got, nn_o = oracle.normals(pts, k, viewpoint=(0.1, -0.2, 0.0), want_nn=True)
        ref, nn_w = W.normals(pts, k, viewpoint=(0.1, -0.2, 0.0))
        ok = np.isfinite(ref[:, 0])
        assert np.array_equal(np.isfinite(got[:, 0]), ok)
        # same FLANN family, same L2_Simple distances: identical lists except where two neighbours are exactly equidistant
        same = (nn_o[ok] == nn_w[ok]).all(1)
        assert same.mean() > 0.995
        rows = np.flatnonzero(ok)[same]
        # libm vs numpy float trigonometry: a few ulps on the roots -> normals to ~1e-5, curvature to ~1e-6 relative
        cosang = np.abs(np.sum(got[rows, :3] * ref[rows, :3], 1))
        assert np.all(np.sum(got[rows, :3] * ref[rows, :3], 1) > 0)  # same flip
        assert np.degrees(np.arccos(np.clip(cosang, -1, 1))).max() < 0.05
        assert np.median(np.abs(got[rows, 4] - ref[rows, 3]) / np.maximum(ref[rows, 3], 1e-12)) < 1e-4
