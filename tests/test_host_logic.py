"""CPU checks of the device-side arithmetic (core_math.cuh compiled for the host by
tests/host/host_check.cu) against the oracle.  These run without a GPU; the real parity tests
(through the C ABI, on the B200) are in test_gpu_parity.py."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

from oracle import default_params, IcpParams
from pose_estimation_b200.testing import synth
from util import tie_ok, pose_delta

HOST = Path(__file__).resolve().parent / "host"


@pytest.fixture(scope="module")
def hc():
    subprocess.run(["make", "-C", str(HOST)], check=True, capture_output=True)
    L = C.CDLL(str(HOST / "libpe_hostcheck.so"))
    vp, sz = C.c_void_p, C.c_size_t
    L.hc_grid_nn.argtypes = [vp, sz, sz, vp, sz, sz, C.c_float, C.c_float, C.c_float, vp, vp, vp]
    L.hc_grid_nn_warm.argtypes = [vp, sz, sz, vp, sz, sz, C.c_float, vp, C.c_float, C.c_float, vp, vp, vp]
    L.hc_grid_nn_seeded.argtypes = [vp, sz, sz, vp, vp, sz, C.c_float, C.c_float, vp, vp]
    L.hc_grid_nn_warm_graph.argtypes = [vp, sz, sz, vp, sz, sz, C.c_float, vp, C.c_float, vp, vp, C.c_int]
    L.hc_rotation_paths.argtypes = [vp, vp, vp]
    L.hc_umeyama_pairs.argtypes = [vp, vp, sz, vp]
    L.hc_lls_pairs.argtypes = [vp, vp, vp, sz, vp]
    L.hc_criteria_script.argtypes = [vp, vp, vp, vp, sz, vp, vp, vp]
    L.hc_normal_from_neighbours.argtypes = [vp, C.c_int, vp, vp, vp]
    L.hc_transforms.argtypes = [vp] * 4
    L.hc_mat4_mul.argtypes = [vp] * 3
    return L


def grid_nn(hc, tgt, q, occupancy=2.0, h=0.0, stop=np.inf):
    tgt = np.ascontiguousarray(tgt, np.float32)
    q = np.ascontiguousarray(q, np.float32)
    idx = np.empty(len(q), np.int32)
    d2 = np.empty(len(q), np.float32)
    rings = np.empty(len(q), np.int32)
    hc.hc_grid_nn(tgt.ctypes.data, len(tgt), tgt.strides[0], q.ctypes.data, len(q), q.strides[0], occupancy, h, stop,
                  idx.ctypes.data, d2.ctypes.data, rings.ctypes.data)
    return idx, d2, rings


def test_grid_search_is_exact_on_surface_cloud(hc, oracle):
    prob = synth.make_c1(4000, seed=11)
    bi, bd = oracle.nn_bruteforce(prob.target, prob.source)
    gi, gd, rings = grid_nn(hc, prob.target, prob.source)
    assert np.array_equal(gd, bd)
    assert tie_ok(prob.target, prob.source, gi, bi)
    assert rings.max() < 40


@pytest.mark.parametrize("occupancy", [0.5, 2.0, 16.0])
def test_grid_search_far_and_outside_queries(hc, oracle, occupancy):
    rng = np.random.default_rng(5)
    tgt = rng.uniform(-1, 1, (3000, 3)).astype(np.float32) * np.array([1.0, 0.6, 0.05], np.float32)
    q = np.concatenate([
        rng.uniform(-1, 1, (300, 3)),            # inside the box
        rng.uniform(-3, 3, (300, 3)),            # mostly outside
        rng.uniform(-1, 1, (50, 3)) * [1, 1, 0] + [0, 0, 5.0],   # far above a thin slab
        tgt[:50] + 1e-7,                         # on top of target points
    ]).astype(np.float32)
    bi, bd = oracle.nn_bruteforce(tgt, q)
    gi, gd, _ = grid_nn(hc, tgt, q, occupancy=occupancy)
    assert np.array_equal(gd, bd)
    assert tie_ok(tgt, q, gi, bi)


def test_grid_search_degenerate_targets(hc, oracle):
    rng = np.random.default_rng(6)
    q = rng.uniform(-1, 1, (200, 3)).astype(np.float32)
    # duplicates: lowest index must win; a line; a single point; with non-finite entries
    dup = np.repeat(rng.uniform(-1, 1, (20, 3)).astype(np.float32), 5, axis=0)
    line = np.zeros((500, 3), np.float32)
    line[:, 0] = np.linspace(-1, 1, 500)
    single = np.array([[0.3, 0.2, 0.1]], np.float32)
    holes = rng.uniform(-1, 1, (400, 3)).astype(np.float32)
    holes[::7] = np.nan
    holes[3, 1] = np.inf
    for tgt in (dup, line, single, holes):
        bi, bd = oracle.nn_bruteforce(tgt, q)
        gi, gd, _ = grid_nn(hc, tgt, q)
        assert np.array_equal(gd, bd)
        assert np.array_equal(gi, bi)  # brute force = lowest index on exact ties, and so must the grid
    gi, gd, _ = grid_nn(hc, np.full((4, 3), np.nan, np.float32), q)
    assert (gi == -1).all() and np.isinf(gd).all()


def test_grid_search_stop_distance_only_cuts_rejected_matches(hc, oracle):
    prob = synth.make_c1(3000, seed=12)
    src = prob.source[:, :3] + np.float32(0.02)
    bi, bd = oracle.nn_bruteforce(prob.target, src)
    stop = np.float32(0.01 ** 2)
    gi, gd, _ = grid_nn(hc, prob.target, src, stop=float(stop))
    acc = bd <= stop
    assert acc.any() and (~acc).any()
    assert np.array_equal(gd[acc], bd[acc]) and tie_ok(prob.target, src[acc], gi[acc], bi[acc])
    assert (gd[~acc] > stop).all()  # whatever it returns there is rejected by the caller anyway


def grid_nn_warm(hc, tgt, q, prev, occupancy=2.0, limit=np.inf, margin=-1.0, want_slack=False):
    """margin < 0: the plain warm search; >= 0: the certificate-producing variant"""
    tgt = np.ascontiguousarray(tgt, np.float32)
    q = np.ascontiguousarray(q, np.float32)
    prev = np.ascontiguousarray(prev, np.int32)
    idx = np.empty(len(q), np.int32)
    d2 = np.empty(len(q), np.float32)
    slack = np.empty(len(q), np.float32)
    hc.hc_grid_nn_warm(tgt.ctypes.data, len(tgt), tgt.strides[0], q.ctypes.data, len(q), q.strides[0], occupancy,
                       prev.ctypes.data, limit, margin, idx.ctypes.data, d2.ctypes.data, slack.ctypes.data)
    return (idx, d2, slack) if want_slack else (idx, d2)


@pytest.mark.parametrize("occupancy", [0.5, 2.0, 8.0])
def test_warm_started_search_is_exact_for_any_candidate(hc, oracle, occupancy):
    """The ball search must return the exact nearest neighbour whatever candidate it is seeded
    with: the true answer, a near miss, or an arbitrary far point."""
    rng = np.random.default_rng(15)
    prob = synth.make_c1(4000, seed=16)
    tgt = prob.target.copy()
    tgt[100:110] = tgt[100]  # exact duplicates: lowest index must win
    q = np.concatenate([prob.source[:, :3], tgt[100:101, :3], tgt[:200, :3] + np.float32(3e-4),
                        rng.uniform(-0.2, 0.2, (200, 3)).astype(np.float32) + tgt[:200, :3]])
    bi, bd = oracle.nn_bruteforce(tgt, q)
    seeds = {
        "true": bi,
        "near": np.clip(bi + rng.integers(-3, 4, len(bi)), 0, len(tgt) - 1),
        "random": rng.integers(0, len(tgt), len(bi)),
    }
    for name, prev in seeds.items():
        gi, gd = grid_nn_warm(hc, tgt, q, prev, occupancy)
        assert np.array_equal(gd, bd), name
        assert np.array_equal(gi, bi), name
    # with a rejection limit: everything the caller would accept is still exact
    lim = np.float32(1e-6)
    gi, gd = grid_nn_warm(hc, tgt, q, seeds["random"], occupancy, float(lim))
    acc = bd <= lim
    assert acc.any() and (~acc).any()
    assert np.array_equal(gd[acc], bd[acc]) and np.array_equal(gi[acc], bi[acc])
    assert (gd[~acc] > lim).all()


@pytest.mark.parametrize("offset,occupancy", [(0.0, 0.25), (8.0, 0.25), (40.0, 0.5), (-300.0, 2.0)])
def test_ball_search_in_cell_units_survives_large_coordinates(hc, oracle, offset, occupancy):
    """The ball search works in cell units ((q - origin) / h, computed like the cell assignment of the target): clouds
    far from the origin and grids with thousands of cells per axis stress the rounding its margins must cover."""
    rng = np.random.default_rng(23)
    prob = synth.make_c1(6000, seed=24)
    tgt = np.ascontiguousarray(prob.target[:, :3]) + np.float32(offset)
    tgt[50:55] = tgt[50]
    n = 2000
    base = prob.source[:n, :3] + np.float32(offset)
    step = rng.normal(size=(n, 3)).astype(np.float32)
    step *= (rng.uniform(0.0, 0.004, n) / np.linalg.norm(step, axis=1))[:, None].astype(np.float32)
    q = np.ascontiguousarray(np.concatenate([base + step, tgt[:300] + np.float32(1e-5), tgt[50:51]]))
    bi, bd = oracle.nn_bruteforce(tgt, q)
    for name, prev in (("true", bi), ("random", rng.integers(0, len(tgt), len(bi)))):
        gi, gd = grid_nn_warm(hc, tgt, q, prev, occupancy)
        assert np.array_equal(gd, bd), (name, offset)
        assert np.array_equal(gi, bi), (name, offset)


@pytest.mark.parametrize("occupancy", [0.5, 2.0, 8.0])
def test_seeded_and_large_ball_searches_are_exact(hc, oracle, occupancy):
    """Far queries (balls spanning dozens of cells) seeded with the match of
    a neighbouring query, and random far seeds through the plain warm search."""
    rng = np.random.default_rng(19)
    prob = synth.make_c1(6000, seed=20)
    tgt = np.ascontiguousarray(prob.target[:, :3])
    n = 1500
    base = prob.source[:n, :3]
    offsets = rng.normal(size=(n, 3)).astype(np.float32)
    offsets *= (rng.uniform(0.0, 0.02, n) / np.linalg.norm(offsets, axis=1))[:, None].astype(np.float32)
    q = np.ascontiguousarray(base + offsets)                                  # up to 20 mm off the surface
    sq = np.ascontiguousarray(q + rng.normal(0, 1.5e-3, q.shape).astype(np.float32))   # the "anchor" of each query
    bi, bd = oracle.nn_bruteforce(tgt, q)
    idx = np.empty(n, np.int32)
    d2 = np.empty(n, np.float32)
    for limit in (np.inf, 4e-4):
        hc.hc_grid_nn_seeded(tgt.ctypes.data, len(tgt), tgt.strides[0], q.ctypes.data, sq.ctypes.data, n, occupancy,
                             limit, idx.ctypes.data, d2.ctypes.data)
        acc = bd <= limit
        assert np.array_equal(d2[acc], bd[acc]) and np.array_equal(idx[acc], bi[acc])
        assert (d2[~acc] > limit).all()
    gi, gd = grid_nn_warm(hc, tgt, q, rng.integers(0, len(tgt), n), occupancy)   # arbitrary seeds: huge balls
    assert np.array_equal(gd, bd) and np.array_equal(gi, bi)


@pytest.mark.parametrize("margin", [0.0, 2e-4, 1e-3])
def test_search_certificate_is_a_true_lower_bound(hc, oracle, margin):
    """slack must never exceed the real gap between the runner-up and the winner, and a query moved
    by less than slack / 2 must keep its nearest neighbour (what lets icp.cu skip searches)."""
    rng = np.random.default_rng(17)
    prob = synth.make_c1(4000, seed=18)
    tgt = prob.target[:, :3].copy()
    tgt[50:53] = tgt[50]
    q = np.concatenate([prob.source[:, :3], tgt[50:51], tgt[:100] + np.float32(2e-4)]).astype(np.float32)
    bi, bd = oracle.nn_bruteforce(tgt, q)
    for prev in (bi, rng.integers(0, len(tgt), len(bi))):
        gi, gd, slack = grid_nn_warm(hc, tgt, q, prev, 2.0, np.inf, margin, want_slack=True)
        assert np.array_equal(gi, bi) and np.array_equal(gd, bd)
        d = np.linalg.norm(tgt[None, :, :].astype(np.float64) - q[:, None, :].astype(np.float64), axis=2)
        d[np.arange(len(q)), bi] = np.inf
        gap = d.min(1) - np.sqrt(bd.astype(np.float64))
        assert (slack <= np.maximum(gap, 0) + 1e-9).all()
        assert (slack <= margin + 1e-9).all()
        if margin > 0:
            assert (slack > 0).mean() > 0.5          # most queries get a usable certificate
            assert slack[len(prob.source)] <= 0      # the query on top of the duplicated target point does not
        # move every certified query by just under slack / 2 in a random direction: same neighbour
        ok = slack > 0
        step = rng.normal(size=(ok.sum(), 3))
        step *= (0.499 * slack[ok] / np.linalg.norm(step, axis=1))[:, None]
        q2 = (q[ok].astype(np.float64) + step).astype(np.float32)
        bi2, _ = oracle.nn_bruteforce(tgt, q2)
        assert np.array_equal(bi2, bi[ok])


def test_umeyama_from_sums_matches_oracle(hc, oracle):
    rng = np.random.default_rng(7)
    prob = synth.make_c1(5000, seed=13)
    s = prob.source[:, :3].copy()
    t = prob.target[:, :3].copy()
    T = np.zeros(16, np.float32)
    hc.hc_umeyama_pairs(s.ctypes.data, t.ctypes.data, len(s), T.ctypes.data)
    Tg = T.reshape(4, 4).T
    To = np.zeros(16, np.float32)
    oracle.L.orc_umeyama(s.ctypes.data, t.ctypes.data, len(s), 1, To.ctypes.data)  # double umeyama on demeaned data
    rot, tr = pose_delta(Tg, To.reshape(4, 4).T)
    assert rot < 1e-7 and tr < 1e-7
    oracle.L.orc_umeyama(s.ctypes.data, t.ctypes.data, len(s), 0, To.ctypes.data)  # PCL's float umeyama
    rot, tr = pose_delta(Tg, To.reshape(4, 4).T)
    assert rot < 5e-6 and tr < 5e-6
    # reflection / planar case keeps det(R) = +1
    s2 = rng.normal(size=(50, 3)).astype(np.float32)
    s2[:, 2] = 0
    t2 = (s2 * np.array([1, 1, 1], np.float32)) + np.float32(0.1)
    hc.hc_umeyama_pairs(s2.ctypes.data, t2.ctypes.data, len(s2), T.ctypes.data)
    assert abs(np.linalg.det(T.reshape(4, 4).T[:3, :3].astype(np.float64)) - 1) < 1e-5


def test_polar_rotation_equals_svd_rotation(hc):
    """The fast path of the solve (scaled Newton polar decomposition) must give umeyama's U V^T to double
    precision whenever it accepts the matrix, and must decline reflections / rank-deficient matrices."""
    rng = np.random.default_rng(23)
    Rp = np.zeros(9)
    Rs = np.zeros(9)
    accepted = 0
    worst = 0.0
    for k in range(400):
        U, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        V, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        if np.linalg.det(U) * np.linalg.det(V) < 0:
            V[:, 0] *= -1
        sv = np.sort(10.0 ** rng.uniform(-6, 0, 3))[::-1] * 10.0 ** rng.uniform(-8, 2)   # cond up to 1e6, any scale
        A = np.ascontiguousarray(U @ np.diag(sv) @ V.T)
        ok = hc.hc_rotation_paths(A.ctypes.data, Rp.ctypes.data, Rs.ctypes.data)
        ref = (U @ V.T).reshape(-1)
        assert np.abs(Rs - ref).max() < 1e-9
        if ok:
            accepted += 1
            worst = max(worst, np.abs(Rp - Rs).max())
            assert np.abs(Rp.reshape(3, 3) @ Rp.reshape(3, 3).T - np.eye(3)).max() < 1e-14
    assert accepted > 380 and worst < 1e-10 * 1e3   # agreement limited by conditioning: 1e-16 * cond
    # ICP-like covariances (well conditioned): agreement at the 1e-15 level
    for k in range(200):
        p = rng.normal(size=(500, 3)) * [0.05, 0.03, 0.004]
        Rt = synth.rotation_about(synth.random_unit(rng), rng.uniform(0, 0.3))
        q = p @ Rt.T + rng.normal(0, 1e-3, p.shape)
        A = np.ascontiguousarray((q - q.mean(0)).T @ (p - p.mean(0)) / len(p))
        assert hc.hc_rotation_paths(A.ctypes.data, Rp.ctypes.data, Rs.ctypes.data) == 1
        assert np.abs(Rp - Rs).max() < 5e-14
    # reflection (det < 0), rank 2 and zero matrices are declined
    for A in (np.diag([1.0, 1.0, -1.0]), np.diag([1.0, 0.5, 0.0]), np.zeros((3, 3)), np.diag([1.0, 1e-9, 1e-9])):
        A = np.ascontiguousarray(A)
        assert hc.hc_rotation_paths(A.ctypes.data, Rp.ctypes.data, Rs.ctypes.data) == 0


def test_lls_from_sums_matches_oracle_bitwise(hc, oracle):
    rng = np.random.default_rng(8)
    n = 4000
    s = rng.uniform(-0.1, 0.1, (n, 3)).astype(np.float32) + np.array([0, 0, 0.7], np.float32)
    nr = rng.normal(size=(n, 3))
    nr = (nr / np.linalg.norm(nr, axis=1, keepdims=True)).astype(np.float32)
    R = synth.rotation_about([1, 2, 3], np.deg2rad(1.0))
    d = (s @ R.T + np.array([0.002, -0.001, 0.003])).astype(np.float32)
    Tg = np.zeros(16, np.float32)
    To = np.zeros(16, np.float32)
    hc.hc_lls_pairs(s.ctypes.data, d.ctypes.data, nr.ctypes.data, n, Tg.ctypes.data)
    oracle.L.orc_point_to_plane_lls(s.ctypes.data, d.ctypes.data, nr.ctypes.data, n, To.ctypes.data)
    # same accumulation order here (sequential), LU vs explicit inverse: agree to float rounding
    assert np.allclose(Tg, To, rtol=0, atol=2e-7)


def test_criteria_state_machine_matches_oracle(hc, oracle):
    rng = np.random.default_rng(9)

    def run(prm, incs, mses, ncorr):
        n = len(mses)
        incs_c = np.ascontiguousarray(np.asarray(incs, np.float32).transpose(0, 2, 1)).reshape(n, 16)
        m = np.asarray(mses, np.float64)
        nc = np.asarray(ncorr, np.int32)
        st = np.zeros(n, np.int32)
        cv = np.zeros(n, np.int32)
        it = np.zeros(n, np.int32)
        hc.hc_criteria_script(C.byref(prm), incs_c.ctypes.data, m.ctypes.data, nc.ctypes.data, n, st.ctypes.data,
                              cv.ctypes.data, it.ctypes.data)
        os_ = np.zeros(n, np.int32)
        oret = np.zeros(n, np.int32)
        # the oracle's script feeds float(mse) through one pseudo correspondence
        oracle.L.orc_criteria_script(C.byref(prm), incs_c.ctypes.data, m.astype(np.float32).astype(np.float64).ctypes.data,
                                     n, os_.ctypes.data, oret.ctypes.data)
        return st, cv, it, os_, oret

    big = [synth.make_pose(synth.rotation_about(synth.random_unit(rng), 0.05), [0.01, 0, 0]) for _ in range(12)]
    tiny = [np.eye(4) for _ in range(12)]
    mses = [float(np.float32(1e-3 / (i + 1))) for i in range(12)]
    cases = [
        (default_params(max_iterations=5), big, mses),                                   # ITERATIONS
        (default_params(max_iterations=50), big, [float(np.float32(1e-4))] * 12),        # ABS_MSE at step 2
        (default_params(max_iterations=50, transformation_epsilon=1e-8), tiny, mses),     # TRANSFORM
        (default_params(max_iterations=50, euclidean_fitness_epsilon=0.6, abs_mse_threshold=-1.0), big, mses),  # REL_MSE
        (default_params(max_iterations=50, transformation_epsilon=1e-8, max_iterations_similar=3), tiny, mses),
    ]
    for prm, incs, ms in cases:
        st, cv, it, os_, oret = run(prm, incs, ms, [100] * 12)
        stop = int(np.argmax(os_ != 0)) if (os_ != 0).any() else len(os_) - 1
        assert np.array_equal(st[: stop + 1], os_[: stop + 1])
        assert np.array_equal(cv[: stop + 1], oret[: stop + 1])
        assert it[stop] == stop + 1
        assert (st[stop:] == st[stop]).all()  # the device state stays frozen once it left NOT_CONVERGED
    # too few correspondences: state 5, not converged, iteration not counted
    st, cv, it, _, _ = run(default_params(max_iterations=50), big, mses, [100, 100, 2] + [100] * 9)
    assert st[2] == 5 and cv[2] == 0 and it[2] == 2 and (st[2:] == 5).all()


def test_normal_from_neighbours_matches_oracle(hc, oracle):
    prob = synth.make_c1(3000, seed=14)
    pts = prob.target
    k = 20
    ref, nn = oracle.normals(pts, k, want_nn=True)
    out = np.zeros(8, np.float32)
    vp = np.zeros(3, np.float32)
    worst = 0.0
    for i in range(0, len(pts), 37):
        nb = np.ascontiguousarray(pts[nn[i], :3])
        p = np.ascontiguousarray(pts[i, :3])
        hc.hc_normal_from_neighbours(nb.ctypes.data, k, p.ctypes.data, vp.ctypes.data, out.ctypes.data)
        worst = max(worst, float(np.abs(out - ref[i]).max()))
    assert worst < 1e-5  # same float arithmetic; libm vs CUDA's host math may differ in the last ulp


def test_transform_and_product_orders_match_oracle(hc, oracle):
    rng = np.random.default_rng(10)
    for _ in range(50):
        T = synth.make_pose(synth.rotation_about(synth.random_unit(rng), rng.uniform(0, 1)), rng.normal(size=3))
        Tc = np.ascontiguousarray(T.astype(np.float32).T).reshape(16)
        p = rng.normal(size=3).astype(np.float32)
        a = np.zeros(3, np.float32)
        b = np.zeros(3, np.float32)
        oa = np.zeros(3, np.float32)
        ob = np.zeros(3, np.float32)
        hc.hc_transforms(Tc.ctypes.data, p.ctypes.data, a.ctypes.data, b.ctypes.data)
        oracle.L.orc_transform_icp(Tc.ctypes.data, p.ctypes.data, oa.ctypes.data)
        oracle.L.orc_transform_tpc(Tc.ctypes.data, p.ctypes.data, ob.ctypes.data)
        assert np.array_equal(a, oa) and np.array_equal(b, ob)
        U = np.ascontiguousarray(rng.normal(size=(4, 4)).astype(np.float32)).reshape(16)
        c = np.zeros(16, np.float32)
        oc = np.zeros(16, np.float32)
        hc.hc_mat4_mul(Tc.ctypes.data, U.ctypes.data, c.ctypes.data)
        oracle.L.orc_mul4(Tc.ctypes.data, U.ctypes.data, oc.ctypes.data)
        assert np.array_equal(c, oc)
