"""Parity of the CUDA path (through the C ABI of libpe_b200.so) against the CPU oracle.

Bars (BASELINE.json north star): correspondence indices bit-exact except equidistant ties within
1e-6 m; final transforms within 1e-5 rad / 1e-5 m; fitness within 1e-6 relative; VoxelGrid
bit-exact; iteration counts and convergence states equal.
"""
import ctypes as C

import numpy as np
import pytest

from oracle import DBL_MAX, default_params
from pose_estimation_b200.testing import synth
from util import pose_delta, tie_ok

pytestmark = pytest.mark.gpu

ROT_TOL = 1e-5   # rad
TRANS_TOL = 1e-5  # m
FIT_RTOL = 1e-6


@pytest.fixture(scope="module")
def pcl():
    from pose_estimation_b200 import pcl as m

    return m


@pytest.fixture(scope="module")
def ctx(pcl):
    c = pcl.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def c1():
    return synth.make_c1(20000, seed=1)


@pytest.fixture(scope="module")
def scene_small(oracle):
    """C2 at 1/4 linear scale: ~12.5k-pt scene, ~3.1k-pt model (the oracle finishes in seconds)."""
    return synth.make_c2(scale=0.25, downsample=lambda p, leaf: oracle.voxel_grid(p, leaf)[0])


def _set_params(icp, prm):
    for name, _ in prm._fields_:
        setattr(icp.params, name, getattr(prm, name))


# PCL runs umeyama in float32.  Its result depends on the float summation order at
# the 3e-5 rad level: the oracle itself moves by that much when the same source points are merely
# permuted (test_icp_float_noise_floor_of_the_reference measures it).  The north-star tolerances
# (1e-5 rad / 1e-5 m / 1e-6 relative fitness) are therefore asserted against the oracle with the
# 3x3 moment sums widened to double (`wide_accum`, otherwise identical code), and against the
# float32 oracle with the looser, measured noise band below.
FLOAT_ROT_TOL = 1.5e-4
FLOAT_TRANS_TOL = 5e-5
FLOAT_FIT_RTOL = 1e-3


def _check_align(pcl, ctx, oracle, source, target, prm, guess=None, normals=None, cls=None, rot_tol=ROT_TOL,
                 trans_tol=TRANS_TOL):
    cls = cls or pcl.IterativeClosestPoint
    icp = cls(ctx)
    icp.setInputSource(source)
    icp.setInputTarget(target, normals)
    _set_params(icp, prm)
    aligned = icp.align(guess, want_correspondences=True)
    got = icp.result
    cap = max(prm.max_iterations, 1)
    ref = oracle.icp(target, normals, wide_accum=True).align(source, guess, prm, trace_cap=cap)
    r = ref["result"]
    if got.iterations != r.iterations:
        # only an MSE-difference threshold (|mse_k - mse_k-1| < 1e-12 ...) may fire an iteration or two
        # apart, and only after an equidistant tie put the two runs on separate trajectories
        assert got.state == r.state and got.state in (3, 4) and abs(got.iterations - r.iterations) <= 3
        rot, tr = pose_delta(icp.getFinalTransformation(), r.matrix())
        assert rot < rot_tol and tr < trans_tol, (rot, tr)
        assert abs(got.fitness - r.fitness) <= 1e-3 * abs(r.fitness)
        return icp, ref
    assert got.state == r.state
    assert got.converged == r.converged
    assert got.n_correspondences == r.n_correspondences
    rot, tr = pose_delta(icp.getFinalTransformation(), r.matrix())
    assert rot < rot_tol and tr < trans_tol, (rot, tr)
    # Same correspondences => same arithmetic => the per-iteration increments are bit-identical.
    # They stop being identical only after an equidistant tie was broken differently (FLANN: first
    # visited, here: lowest index — the north star's tie allowance); from then on the two runs are
    # two valid ICP trajectories a few 1e-6 rad apart and the fitness is compared more loosely.
    tg, to = icp.trace(), ref["trace_T"]
    assert tg.shape == to.shape
    lockstep = np.array_equal(tg.view(np.uint32), to.view(np.uint32))
    fit_rtol = FIT_RTOL if lockstep else 1e-3
    if r.fitness < DBL_MAX:
        assert abs(got.fitness - r.fitness) <= fit_rtol * abs(r.fitness), (got.fitness, r.fitness)
        # the fitness function itself, on the oracle's own final transform: 1e-6 relative always
        f_on_ref, _ = ctx.fitness_score(r.matrix(), prm.fitness_max_range)
        assert abs(f_on_ref - r.fitness) <= FIT_RTOL * abs(r.fitness), (f_on_ref, r.fitness)
    else:
        assert got.fitness == r.fitness
    if r.last_mse < DBL_MAX:
        assert abs(got.last_mse - r.last_mse) <= fit_rtol * abs(r.last_mse)
    for a, b in zip(tg, to):
        drot, dtr = pose_delta(a, b)
        assert drot < rot_tol and dtr < trans_tol
    # correspondences of the last iteration: bit-exact indices except equidistant ties within 1e-6 m
    idx, d2 = icp.correspondences
    assert tie_ok(target, _last_work(ref, source, guess), idx, ref["corr_idx"], tol=1e-6)
    same = idx == ref["corr_idx"]
    assert same.mean() > 0.999
    if lockstep:
        assert np.array_equal(d2[same], ref["corr_d2"][same])
    else:
        assert np.allclose(d2[same], ref["corr_d2"][same], rtol=0, atol=2e-6 * np.sqrt(np.maximum(d2[same], 1e-12)) + 1e-12)
    # output cloud = final * input
    assert np.nanmax(np.abs(aligned[:, :3] - ref["aligned"][:, :3])) < 1e-5
    assert np.array_equal(np.isnan(aligned), np.isnan(ref["aligned"]))
    # the float32 oracle (PCL's own arithmetic): same outcome within its summation-order noise
    if prm.estimator == 0:
        f = oracle.icp(target, normals).align(source, guess, prm)["result"]
        if f.iterations == got.iterations:  # an MSE-threshold stop may legitimately fire one iteration apart
            rot, tr = pose_delta(icp.getFinalTransformation(), f.matrix())
            assert rot < FLOAT_ROT_TOL and tr < FLOAT_TRANS_TOL, (rot, tr)
            if f.fitness < DBL_MAX:
                assert abs(got.fitness - f.fitness) <= FLOAT_FIT_RTOL * abs(f.fitness)
        else:
            # threshold crossings of a sequence that carries the float32 noise: a few iterations apart
            assert abs(f.iterations - got.iterations) <= max(6, got.iterations // 5)
    return icp, ref


def _last_work(ref, source, guess):
    """The oracle's working cloud at its last correspondence search (float64 is fine for tie checks)."""
    T = np.eye(4) if guess is None else np.asarray(guess, np.float64)
    incs = ref["trace_T"]
    n_apply = len(incs) - 1 if ref["result"].state != 5 else len(incs)
    for k in range(max(n_apply, 0)):
        T = incs[k].astype(np.float64) @ T
    return synth.apply_pose(T, np.asarray(source, np.float64)[:, :3])


# ---- nearest neighbour -------------------------------------------------------------------------
def test_nn_grid_equals_bruteforce_equals_oracle(ctx, oracle, c1):
    ctx.target_set(c1.target)
    gi, gd = ctx.nn_search(c1.source)
    bi, bd = ctx.nn_search(c1.source, bruteforce=True)
    oi, od = oracle.knn(c1.target, c1.source, 1)
    assert np.array_equal(gi, bi) and np.array_equal(gd, bd)      # validator: bit-exact, same tie rule
    assert np.array_equal(gd, od[:, 0])                           # distances bit-exact vs FLANN arithmetic
    assert tie_ok(c1.target, c1.source, gi, oi[:, 0])


def test_nn_golden_flann_vectors(ctx, golden):
    for case in ("a", "b"):
        tgt, qry = golden[f"{case}_target"], golden[f"{case}_query"]
        ctx.target_set(tgt)
        gi, gd = ctx.nn_search(qry)
        # OpenCV's bundled FLANN KDTreeSingleIndex (tests/golden/make_golden.py): distances bit-exact
        assert np.array_equal(gd, golden[f"{case}_d1"].reshape(-1))
        assert tie_ok(tgt, qry, gi, golden[f"{case}_idx1"].reshape(-1), tol=0.0)


@pytest.mark.parametrize("group", [1, 2, 4, 8, 16])
def test_nn_all_group_widths_and_far_queries(ctx, oracle, group):
    rng = np.random.default_rng(3)
    tgt = (rng.uniform(-1, 1, (5000, 3)) * [1.0, 0.6, 0.05]).astype(np.float32)
    tgt[::11] = np.nan
    q = np.concatenate([rng.uniform(-1, 1, (500, 3)), rng.uniform(-3, 3, (500, 3)),
                        rng.uniform(-1, 1, (64, 3)) * [1, 1, 0] + [0, 0, 40.0], tgt[1:50] + 1e-7]).astype(np.float32)
    q[7] = np.nan
    ctx.set_int("nn_group", group)
    try:
        ctx.target_set(tgt)
        gi, gd = ctx.nn_search(q)
        bi, bd = ctx.nn_search(q, bruteforce=True)
    finally:
        ctx.set_int("nn_group", 1)
    oi, od = oracle.nn_bruteforce(tgt, q)
    ok = np.isfinite(q).all(1)
    assert np.array_equal(gi[ok], oi[ok]) and np.array_equal(gd[ok], od[ok])
    assert np.array_equal(bi[ok], oi[ok]) and np.array_equal(bd[ok], od[ok])
    assert (gi[~ok] == -1).all() and (bi[~ok] == -1).all()


def test_nn_degenerate_targets(ctx, oracle):
    rng = np.random.default_rng(6)
    q = rng.uniform(-1, 1, (300, 3)).astype(np.float32)
    dup = np.repeat(rng.uniform(-1, 1, (20, 3)).astype(np.float32), 5, axis=0)
    line = np.zeros((500, 3), np.float32)
    line[:, 0] = np.linspace(-1, 1, 500)
    single = np.array([[0.3, 0.2, 0.1]], np.float32)
    for tgt in (dup, line, single):
        ctx.target_set(tgt)
        gi, gd = ctx.nn_search(q)
        oi, od = oracle.nn_bruteforce(tgt, q)
        assert np.array_equal(gi, oi) and np.array_equal(gd, od)
    ctx.target_set(np.full((5, 3), np.nan, np.float32))
    gi, gd = ctx.nn_search(q)
    assert (gi == -1).all() and np.isinf(gd).all()
    ctx.target_set(np.empty((0, 3), np.float32))
    gi, gd = ctx.nn_search(q)
    assert (gi == -1).all()
    gi, gd = ctx.nn_search(np.empty((0, 3), np.float32))
    assert gi.shape == (0,)


def test_cloud_strides(ctx, oracle, c1):
    """cv::Mat N x 3 (12 B), PointXYZ (16 B), N x 6 (24 B) and PointNormal (48 B) records."""
    t3 = np.ascontiguousarray(c1.target[:2000, :3])
    q3 = np.ascontiguousarray(c1.source[:500, :3])
    oi, od = oracle.nn_bruteforce(t3, q3)
    for width in (3, 4, 6, 12):
        t = np.zeros((len(t3), width), np.float32)
        t[:, :3] = t3
        t[:, 3:] = 123.0
        q = np.zeros((len(q3), width), np.float32)
        q[:, :3] = q3
        ctx.target_set(t)
        gi, gd = ctx.nn_search(q)
        assert np.array_equal(gi, oi) and np.array_equal(gd, od)


# ---- ICP -----------------------------------------------------------------------------------------
def test_icp_c1_fixed_30_iterations(pcl, ctx, oracle, c1):
    prm = default_params(max_iterations=30, abs_mse_threshold=-1.0)
    icp, ref = _check_align(pcl, ctx, oracle, c1.source, c1.target, prm)
    assert icp.nr_iterations_ == 30 and icp.hasConverged()
    rot, tr = pose_delta(icp.getFinalTransformation(), c1.gt_pose)
    # point-to-point ICP slides slowly along a smooth sheet: 3 deg / 17 mm -> < 0.7 deg / 1.5 mm in 30 iterations
    assert rot < 1.3e-2 and tr < 1.5e-3
    # per-iteration increments follow the oracle's trace
    tg = icp.trace()
    assert tg.shape == ref["trace_T"].shape
    for a, b in zip(tg, ref["trace_T"]):
        rot, tr = pose_delta(a, b)
        assert rot < ROT_TOL and tr < TRANS_TOL


def test_icp_float_noise_floor_of_the_reference(pcl, ctx, oracle, c1):
    """Quantifies why the float32 oracle cannot be matched to 1e-5 rad by anything, itself included:
    permuting the source points (same set, another float summation order) moves its answer by more
    than the CUDA path's distance to the double-accumulating oracle."""
    prm = default_params(max_iterations=30, abs_mse_threshold=-1.0)
    perm = np.random.default_rng(0).permutation(len(c1.source))
    a = oracle.icp(c1.target).align(c1.source, None, prm)["result"]
    b = oracle.icp(c1.target).align(c1.source[perm], None, prm)["result"]
    w = oracle.icp(c1.target, wide_accum=True).align(c1.source, None, prm)["result"]
    icp = pcl.IterativeClosestPoint(ctx)
    icp.setInputSource(c1.source)
    icp.setInputTarget(c1.target)
    _set_params(icp, prm)
    icp.align(want_output=False)
    icp_perm = pcl.IterativeClosestPoint(ctx)
    icp_perm.setInputSource(c1.source[perm])
    icp_perm.setInputTarget(c1.target)
    _set_params(icp_perm, prm)
    icp_perm.align(want_output=False)
    noise_rot, noise_tr = pose_delta(a.matrix(), b.matrix())
    gpu_rot, gpu_tr = pose_delta(icp.getFinalTransformation(), w.matrix())
    gpu_perm_rot, gpu_perm_tr = pose_delta(icp.getFinalTransformation(), icp_perm.getFinalTransformation())
    assert noise_rot > 1e-5                      # the reference arithmetic is order-dependent beyond the bar
    assert gpu_rot < 5e-6 and gpu_tr < 5e-6      # the CUDA path sits on the double-accumulated answer
    assert gpu_perm_rot < 5e-6 and gpu_perm_tr < 5e-6   # and does not depend on the point order
    assert noise_rot > 3 * max(gpu_rot, gpu_perm_rot)
    rot, tr = pose_delta(icp.getFinalTransformation(), a.matrix())
    assert rot < 3 * noise_rot + 1e-5 and tr < 3 * noise_tr + 1e-5


def test_icp_c1_pcl_defaults_state_machine(pcl, ctx, oracle, c1):
    _check_align(pcl, ctx, oracle, c1.source, c1.target, default_params())                       # 10 iterations
    _check_align(pcl, ctx, oracle, c1.source, c1.target, default_params(max_iterations=200))     # ABS_MSE stop
    _check_align(pcl, ctx, oracle, c1.source, c1.target,
                 default_params(max_iterations=100, transformation_epsilon=1e-9))               # TRANSFORM stop
    _check_align(pcl, ctx, oracle, c1.source, c1.target,
                 default_params(max_iterations=100, euclidean_fitness_epsilon=1e-3, abs_mse_threshold=-1.0))  # REL_MSE
    _check_align(pcl, ctx, oracle, c1.source, c1.target, default_params(max_iterations=1))


def test_icp_threshold_rejector_guess_and_no_correspondences(pcl, ctx, oracle, c1):
    guess = synth.make_pose(synth.rotation_about([0, 1, 0], np.deg2rad(0.5)), [0.001, 0.0, -0.001])
    _check_align(pcl, ctx, oracle, c1.source, c1.target,
                 default_params(max_iterations=15, max_corr_dist=0.004, abs_mse_threshold=-1.0), guess=guess)
    _check_align(pcl, ctx, oracle, c1.source, c1.target,
                 default_params(max_iterations=15, rejector_max_dist=0.003, abs_mse_threshold=-1.0))
    # a threshold nobody passes: NO_CORRESPONDENCES, not converged, zero iterations
    icp, ref = _check_align(pcl, ctx, oracle, c1.source, c1.target, default_params(max_corr_dist=1e-7))
    assert icp.result.state == 5 and not icp.hasConverged() and icp.nr_iterations_ == 0
    assert (icp.correspondences[0] == -1).sum() >= len(c1.source) - 2


def test_icp_nonfinite_points(pcl, ctx, oracle, c1):
    src = c1.source[:4000].copy()
    tgt = c1.target.copy()
    src[::13, 0] = np.nan
    tgt[::17, 2] = np.inf
    _check_align(pcl, ctx, oracle, src, tgt, default_params(max_iterations=12, abs_mse_threshold=-1.0))


def test_icp_scene_config_small(pcl, ctx, oracle, scene_small):
    p = scene_small
    prm = default_params(max_iterations=30, abs_mse_threshold=-1.0)
    icp, _ = _check_align(pcl, ctx, oracle, p.source, p.target, prm, guess=p.guess)
    rot, tr = pose_delta(icp.getFinalTransformation(), p.gt_pose)
    assert rot < 5e-3 and tr < 1e-3


def test_icp_point_to_plane(pcl, ctx, oracle, scene_small):
    p = scene_small
    normals = oracle.normals(p.target, 30)
    prm = default_params(max_iterations=30, abs_mse_threshold=-1.0, estimator=1)
    icp, _ = _check_align(pcl, ctx, oracle, p.source, p.target, prm, guess=p.guess, normals=normals,
                          cls=pcl.IterativeClosestPointWithNormals)
    rot, tr = pose_delta(icp.getFinalTransformation(), p.gt_pose)
    assert rot < 5e-3 and tr < 1e-3
    # PointNormal-style single array (xyz1 | nx ny nz 0 | curvature ...)
    pn = np.zeros((len(p.target), 12), np.float32)
    pn[:, :4] = p.target
    pn[:, 4:8] = normals[:, :4]
    icp2 = pcl.IterativeClosestPointWithNormals(ctx)
    icp2.setInputSource(p.source)
    icp2.setInputTarget(pn)
    _set_params(icp2, prm)
    icp2.align(p.guess)
    assert np.array_equal(icp2.getFinalTransformation(), icp.getFinalTransformation())


def test_icp_batch_equals_single_and_oracle(pcl, ctx, oracle, scene_small):
    p = scene_small
    rng = np.random.default_rng(77)
    guesses = np.stack([synth.perturb_pose(p.gt_pose, rng, 5.0, 0.006) for _ in range(12)] + [p.guess])
    prm = default_params(max_iterations=20, max_corr_dist=0.02)
    icp = pcl.IterativeClosestPoint(ctx)
    icp.setInputSource(p.source)
    icp.setInputTarget(p.target)
    _set_params(icp, prm)
    batch = icp.alignBatch(guesses)
    ref = oracle.icp(p.target, wide_accum=True).align_batch(p.source, guesses, prm)
    for h, (g, r) in enumerate(zip(batch, ref)):
        icp.align(guesses[h], want_output=False)
        s = icp.result
        assert bytes(s.T) == bytes(g.T) and s.fitness == g.fitness and s.iterations == g.iterations  # bit-identical
        assert g.iterations == r.iterations and g.state == r.state and g.converged == r.converged
        rot, tr = pose_delta(pcl.result_matrix(g), r.matrix())
        assert rot < ROT_TOL and tr < TRANS_TOL
        assert abs(g.fitness - r.fitness) <= FIT_RTOL * r.fitness


@pytest.mark.parametrize("option,value", [("warm_start", 0), ("cert_margin_x1000", 300), ("nn_group", 8), ("anchor_seed", 0),
                                          ("pdl", 0), ("batch_streams", 1),
                                          # first iteration of a batch: per-lane verification instead of the warp-cooperative one,
                                          # a row limit that makes most patches fall back, a guard that sends many points to the
                                          # next patch's anchor or to a cold search
                                          ("coop_max_rows", 0), ("coop_max_rows", 40), ("seed_guard_x10", 15),
                                          # whole-grid dependencies between the launches instead of per-hypothesis flags
                                          ("flag_deps", 0),
                                          # the candidate cache of the warm launches (nn_cache.cuh): from launch 1, 3 or 10
                                          # on, with balls small enough that most certificates fail and large ones
                                          ("nn_cache_from", 1), ("nn_cache_from", 3), ("nn_cache_from", 10),
                                          ("nn_cache_from+r", (2, 20)), ("nn_cache_from+r", (2, 200))])
def test_speed_options_never_change_results(pcl, oracle, scene_small, option, value):
    """warm start, search-skipping certificates, the cold lane-group width and the cooperative first iteration
    (packed-arithmetic filter + exact re-evaluation) are exactness-preserving: every combination must give the
    bit-identical answer."""
    p = scene_small
    rng = np.random.default_rng(5)
    guesses = np.stack([synth.perturb_pose(p.gt_pose, rng, 5.0, 0.006) for _ in range(400)])  # enough for anchor seeding
    out = []
    for use in (False, True):
        c = pcl.Context(0)
        if use and option == "nn_cache_from+r":
            c.set_int("nn_cache_from", value[0])
            c.set_int("nn_cache_r_x100", value[1])
        elif use:
            c.set_int(option, value)
        elif option.startswith("nn_cache"):
            c.set_int("nn_cache_from", 0)  # (the comparison run: the plain warm search, whatever the default is)
        icp = pcl.IterativeClosestPoint(c)
        icp.setInputSource(p.source)
        icp.setInputTarget(p.target)
        _set_params(icp, default_params(max_iterations=25, max_corr_dist=0.02, abs_mse_threshold=-1.0))
        res = icp.alignBatch(guesses)
        icp.align(p.guess, want_correspondences=True)
        out.append(([bytes(r.T) + bytes(np.float64(r.fitness)) for r in res], bytes(icp.result.T), icp.result.fitness,
                    icp.correspondences[0].copy(), icp.correspondences[1].copy()))
        c.close()
    a, b = out
    assert a[0] == b[0] and a[1] == b[1] and a[2] == b[2]
    assert np.array_equal(a[3], b[3]) and np.array_equal(a[4], b[4])


@pytest.mark.parametrize("options,H", [({"blocks_factor_cold": 128}, 400), ({"blocks_factor_cold": 8, "blocks_factor": 64}, 400),
                                       ({"nn_group": 8, "batch_streams": 2}, 32), ({"nn_group": 8, "batch_streams": 4}, 96)])
def test_cold_and_warm_launches_with_different_block_counts_do_not_share_partial_records(pcl, scene_small, options, H):
    """The per-block partial sums of a hypothesis live at a stride that does not depend on the launch: launch 0 (cold) and
    the warm launches may use different numbers of blocks per hypothesis (blocks_factor_cold, or a lane-group width > 1
    in launch 0 only), and chains / per-hypothesis launch dependencies let a warm launch of one hypothesis run while the
    cold launch of another still sums.  A different block count changes the ORDER of the double sums (last bits), so the
    check is against single aligns at a tolerance far below any corruption."""
    p = scene_small
    rng = np.random.default_rng(17)
    guesses = np.stack([synth.perturb_pose(p.gt_pose, rng, 5.0, 0.006) for _ in range(H)])
    prm = default_params(max_iterations=12, max_corr_dist=0.02, abs_mse_threshold=-1.0)
    c = pcl.Context(0)
    for k, v in options.items():
        c.set_int(k, v)
    icp = pcl.IterativeClosestPoint(c)
    icp.setInputSource(p.source)
    icp.setInputTarget(p.target)
    _set_params(icp, prm)
    runs = [[bytes(r) for r in icp.alignBatch(guesses)] for _ in range(3)]
    assert runs[0] == runs[1] == runs[2]  # deterministic (an overlap of partial records would depend on timing)
    res = icp.alignBatch(guesses)
    plain = pcl.Context(0)
    ref = pcl.IterativeClosestPoint(plain)
    ref.setInputSource(p.source)
    ref.setInputTarget(p.target)
    _set_params(ref, prm)
    for h in range(0, H, max(1, H // 24)):
        ref.align(guesses[h], want_output=False)
        rot, tr = pose_delta(pcl.result_matrix(res[h]), pcl.result_matrix(ref.result))
        assert rot < 2e-6 and tr < 2e-6, (h, rot, tr)
        assert res[h].iterations == ref.result.iterations and res[h].state == ref.result.state
        assert abs(res[h].fitness - ref.result.fitness) <= 1e-5 * ref.result.fitness
    c.close()
    plain.close()


def test_batch_with_hypotheses_that_stop_at_different_iterations(pcl, scene_small):
    """Convergence criteria on: hypotheses stop after different numbers of iterations, some never get a correspondence.
    The launches of a batch depend on each other per hypothesis (epoch flags, a sentinel when a hypothesis stops): the
    records must equal those of whole-grid dependencies and those of single aligns, bit for bit."""
    p = scene_small
    rng = np.random.default_rng(8)
    guesses = [synth.perturb_pose(p.gt_pose, rng, a, t) for a, t in [(0.2, 0.0003)] * 24 + [(2.0, 0.003)] * 24 + [(6.0, 0.008)] * 24]
    far = np.array(p.gt_pose, np.float64)
    far[:3, 3] += [0.5, 0.5, 0.5]  # nothing within the correspondence distance: PEB_NO_CORRESPONDENCES at iteration 0
    guesses = np.stack(guesses + [far] * 8)
    prm = default_params(max_iterations=40, max_corr_dist=0.01, transformation_epsilon=1e-9)
    out = []
    for deps in (1, 0):
        c = pcl.Context(0)
        c.set_int("flag_deps", deps)
        icp = pcl.IterativeClosestPoint(c)
        icp.setInputSource(p.source)
        icp.setInputTarget(p.target)
        _set_params(icp, prm)
        res = icp.alignBatch(guesses)
        out.append([bytes(r) for r in res])
        if deps:
            its = np.array([r.iterations for r in res])
            states = np.array([r.state for r in res])
            assert len(set(its[:72].tolist())) > 3 and (states[72:] == 5).all() and (its[72:] == 0).all()
            for h in (0, 30, 60, 75):
                icp.align(guesses[h], want_output=False)
                assert bytes(icp.result.T) == bytes(res[h].T) and icp.result.iterations == res[h].iterations
                assert icp.result.state == res[h].state
        c.close()
    assert out[0] == out[1]


def test_icp_edge_cases_and_errors(pcl, ctx, oracle, c1):
    fresh = pcl.Context(0)
    try:
        icp = pcl.IterativeClosestPoint(fresh)
        with pytest.raises(pcl.PebError) as e:
            icp.align()
        assert e.value.code == -2  # NO_TARGET
        icp.setInputTarget(c1.target[:100])
        with pytest.raises(pcl.PebError) as e:
            icp.align()
        assert e.value.code == -3  # NO_SOURCE
        # empty source: no correspondences
        icp.setInputSource(np.empty((0, 4), np.float32))
        icp.align()
        assert icp.result.state == 5 and icp.result.iterations == 0
        # empty target
        icp.setInputSource(c1.source[:100])
        icp.setInputTarget(np.empty((0, 4), np.float32))
        icp.align()
        assert icp.result.state == 5 and icp.result.fitness == DBL_MAX
        # point-to-plane without normals is an argument error, unknown estimator is unsupported
        icp.setInputTarget(c1.target[:100])
        icp.params.estimator = 1
        with pytest.raises(pcl.PebError) as e:
            icp.align()
        assert e.value.code == -1
        icp.params.estimator = 7
        with pytest.raises(pcl.PebError) as e:
            icp.align()
        assert e.value.code == -6
        with pytest.raises(pcl.PebError):
            icp.setUseReciprocalCorrespondences(True)
        bad = np.zeros((10, 3), np.float32)
        rc = pcl.lib.peb_source_set(fresh.handle, bad.ctypes.data, 10, 10)
        assert rc == -1 and b"stride" in pcl.lib.peb_last_error(fresh.handle)
    finally:
        fresh.close()


def test_fitness_score_matches_oracle(pcl, ctx, oracle, c1):
    ctx.target_set(c1.target)
    ctx.source_set(c1.source)
    o = oracle.icp(c1.target)
    T = synth.make_pose(synth.rotation_about([1, 0, 1], 0.01), [0.001, 0.002, 0.0])
    for max_range in (DBL_MAX, 1e-5, 4e-6):
        f, n = ctx.fitness_score(T, max_range)
        rf, rn = o.fitness(c1.source, T, max_range)
        assert n == rn and abs(f - rf) <= 1e-12 * abs(rf)


# ---- VoxelGrid -------------------------------------------------------------------------------------
def test_voxel_grid_bit_exact(pcl, ctx, oracle):
    rng = np.random.default_rng(21)
    surf = synth.Surface(21)
    scene = synth.render_scene(surf, synth.default_gt_pose(rng), rng, 486, 300)
    for leaf, min_pts in ((0.002, 0), (0.004, 3), ((0.003, 0.002, 0.005), 2)):
        vg = pcl.VoxelGrid(ctx)
        vg.setInputCloud(scene)
        if np.isscalar(leaf):
            vg.setLeafSize(leaf)
        else:
            vg.setLeafSize(*leaf)
        vg.setMinimumPointsNumberPerVoxel(min_pts)
        out = vg.filter()
        ref, unchanged = oracle.voxel_grid(scene, leaf, min_pts)
        assert not unchanged
        assert out.shape == ref.shape
        assert np.array_equal(out.view(np.uint32), ref.view(np.uint32))  # bit-exact, same order


def test_voxel_grid_edge_cases(pcl, ctx, oracle):
    vg = pcl.VoxelGrid(ctx)
    vg.setLeafSize(0.01)
    vg.setInputCloud(np.empty((0, 4), np.float32))
    assert vg.filter().shape == (0, 4)
    vg.setInputCloud(np.full((7, 4), np.nan, np.float32))
    assert vg.filter().shape == (0, 4)
    one = np.array([[0.1, -0.2, 0.3, 1.0]], np.float32)
    vg.setInputCloud(one)
    assert np.array_equal(vg.filter(), one)
    # overflow guard: PCL warns and returns the input unchanged
    rng = np.random.default_rng(1)
    big = np.ones((100, 4), np.float32)
    big[:, :3] = rng.uniform(-100, 100, (100, 3))
    vg.setLeafSize(1e-3)
    vg.setInputCloud(big)
    out = vg.filter()
    ref, unchanged = oracle.voxel_grid(big, 1e-3)
    assert unchanged and np.array_equal(out, ref)
    # negative coordinates and duplicates
    pts = np.ones((5000, 4), np.float32)
    pts[:, :3] = rng.normal(0, 0.05, (5000, 3))
    pts[100:200] = pts[0]
    vg.setLeafSize(0.01)
    vg.setInputCloud(pts)
    ref, _ = oracle.voxel_grid(pts, 0.01)
    assert np.array_equal(vg.filter().view(np.uint32), ref.view(np.uint32))
    with pytest.raises(pcl.PebError):
        vg.setLeafSize(0.0)
        vg.filter()


# ---- NormalEstimation --------------------------------------------------------------------------------
@pytest.mark.parametrize("k", [8, 30, 60])
def test_normals_match_oracle(pcl, ctx, oracle, scene_small, k):
    pts = scene_small.target
    ne = pcl.NormalEstimation(ctx)
    ne.setInputCloud(pts)
    ne.setKSearch(k)
    out, nn = ne.compute(return_neighbours=True)
    ref, rnn = oracle.normals(pts, k, want_nn=True)
    # neighbour lists: same sets in the same order up to equidistant ties
    assert tie_ok(pts, pts, nn.reshape(-1), rnn.reshape(-1))
    same = (nn == rnn).all(1)
    assert same.mean() > 0.999
    # where the neighbour order is identical the float arithmetic is identical up to libm vs CUDA
    # atan2f / cosf / sinf (<= 2 ulp), which PCL's closed-form eigen33 amplifies slightly
    d = np.abs(out[same] - ref[same])
    assert np.nanmax(d[:, :3]) < 2e-4 and np.nanmedian(d[:, :3]) < 1e-6
    assert np.nanmax(d[:, 4]) < 1e-4
    ang = np.degrees(np.arccos(np.clip((out[same, :3] * ref[same, :3]).sum(1), -1, 1)))
    assert np.nanmax(ang) < 0.05
    assert np.allclose(np.linalg.norm(out[:, :3], axis=1), 1.0, atol=1e-5)
    assert (out[:, 2] <= 1e-6).mean() > 0.99  # flipped towards the viewpoint at the origin


def test_normals_edge_cases(pcl, ctx, oracle):
    ne = pcl.NormalEstimation(ctx)
    ne.setKSearch(10)
    pts = np.ones((50, 4), np.float32)
    rng = np.random.default_rng(2)
    pts[:, :3] = rng.normal(0, 0.01, (50, 3)) + [0, 0, 0.7]
    pts[5, :3] = np.nan
    ne.setInputCloud(pts)
    ne.setViewPoint(0.0, 0.0, 2.0)
    out = ne.compute()
    ref = oracle.normals(pts, 10, viewpoint=(0, 0, 2.0))
    assert np.isnan(out[5, :3]).all() and np.isnan(out[5, 4])
    ok = np.arange(50) != 5
    assert np.abs(out[ok] - ref[ok]).max() < 1e-3
    # fewer than 3 points: NaN normals; k larger than the cloud: clamped
    ne.setInputCloud(pts[:2])
    assert np.isnan(ne.compute()[:, :3]).all()
    ne.setKSearch(100)
    ne.setInputCloud(pts[10:30])
    out = ne.compute()
    ref = oracle.normals(pts[10:30], 100, viewpoint=(0, 0, 2.0))
    assert np.abs(out - ref).max() < 1e-3
    ne.setInputCloud(np.empty((0, 4), np.float32))
    assert ne.compute().shape == (0, 8)
    ne.setKSearch(500)
    ne.setInputCloud(pts)
    with pytest.raises(pcl.PebError) as e:
        ne.compute()
    assert e.value.code == -6


# ---- scene pre-filter ----------------------------------------------------------
def test_scene_prefilter_bit_exact(pcl, ctx, oracle):
    from oracle import prefilter_params

    rng = np.random.default_rng(31)
    surf = synth.Surface(31)
    scene = synth.render_scene(surf, synth.default_gt_pose(rng), rng, 486, 300)
    scene[5, :3] = 0.0
    plane = (0.0, -0.0872, 0.9962, -0.747)
    for sphere, remove_inliers, planes in ((None, False, []), ((0.0, 0.0, 0.7, 0.08), False, []),
                                           ((0.0, 0.0, 0.7, 0.08), True, [plane]), (None, False, [plane, (1.0, 0.0, 0.0, 0.1)])):
        pf = pcl.ScenePrefilter(ctx)
        pf.setInputCloud(scene)
        if sphere is not None:
            pf.setSphereFilter(sphere[:3], sphere[3], "inliers" if remove_inliers else "outliers")
        for pl in planes:
            pf.addPlane(*pl)
        out = pf.filter()
        ref = oracle.scene_prefilter(scene, prefilter_params(sphere, remove_inliers, planes, 0.005))
        assert out.shape == ref.shape and len(out) > 0
        assert np.array_equal(out.view(np.uint32), ref.view(np.uint32))  # same survivors, original order, same bits
    # edge cases: empty, all NaN, everything filtered out; 12-byte records
    pf = pcl.ScenePrefilter(ctx)
    pf.setInputCloud(np.empty((0, 4), np.float32))
    assert pf.filter().shape == (0, 4)
    pf.setInputCloud(np.full((9, 4), np.nan, np.float32))
    assert pf.filter().shape == (0, 4)
    pf.setSphereFilter((10.0, 10.0, 10.0), 0.01)
    pf.setInputCloud(scene)
    assert pf.filter().shape == (0, 4)
    pf3 = pcl.ScenePrefilter(ctx)
    pf3.setInputCloud(np.ascontiguousarray(scene[:, :3]))
    assert np.array_equal(pf3.filter(), oracle.scene_prefilter(scene, prefilter_params()))
    # the chain the node would run: prefilter -> VoxelGrid
    vg = pcl.VoxelGrid(ctx)
    vg.setInputCloud(pf3.filter())
    vg.setLeafSize(0.003)
    ref_ds, _ = oracle.voxel_grid(oracle.scene_prefilter(scene, prefilter_params()), 0.003)
    assert np.array_equal(vg.filter().view(np.uint32), ref_ds.view(np.uint32))
