"""Adversarial inputs for the grid searches that replace pcl::KdTreeFLANN::nearestKSearch: the
__host__ __device__ search code of core_math.cuh, compiled for the CPU (tests/host/host_check.cu), against the oracle's
brute force (same float L2_Simple arithmetic, lowest index on exact ties) on clouds built to sit on the edges of what
the margins of the searches cover: points and queries exactly on cell boundaries, extents of a few ulps, clouds a
kilometre from the origin, lattices full of exact ties, sheets, lines, duplicates, clusters with empty space between.
Runs without a GPU; the kernels execute the same source."""
import numpy as np
import pytest

from test_host_logic import grid_nn, grid_nn_warm, hc  # noqa: F401  (hc is the fixture)

F = np.float32


def _lattice(rng, n, step, dims=3):
    """points on multiples of `step` (a power of two): every coordinate difference is exact, ties are exact"""
    k = rng.integers(-12, 13, (n, 3)).astype(F)
    k[:, dims:] = 0
    return k * F(step)


def clouds(rng):
    """name -> (n, 3) float32 target"""
    out = {}
    out["lattice"] = _lattice(rng, 300, 2.0 ** -6)
    out["lattice_sheet"] = _lattice(rng, 300, 2.0 ** -8, dims=2) + F([0, 0, 0.5])
    out["lattice_line"] = _lattice(rng, 120, 2.0 ** -5, dims=1)
    out["lattice_far"] = _lattice(rng, 300, 2.0 ** -4) + F([1024.0, -2048.0, 512.0])   # spacing = 512 ulps of x
    out["tiny_extent"] = (F(0.7) + rng.integers(0, 40, (200, 3)).astype(F) * np.spacing(F(0.7))).astype(F)  # a few ulps wide
    out["anisotropic"] = (rng.uniform(0, 1, (400, 3)) * [1.0, 1e-3, 1e-6]).astype(F)
    two = rng.normal(0, 0.01, (300, 3))
    two[150:] += [5.0, -3.0, 2.0]
    out["two_clusters"] = two.astype(F)
    dup = rng.uniform(-1, 1, (40, 3)).astype(F)
    out["duplicates"] = np.concatenate([dup, dup[::-1], dup])
    out["single"] = F([[0.25, -0.5, 0.125]])
    out["pair"] = F([[0.0, 0.0, 0.0], [1.0, 0.0, 0.0]])
    big = rng.uniform(-1, 1, (400, 3))
    out["kilometre_off"] = (big * 0.05 + [1000.0, 1000.0, 1000.0]).astype(F)
    out["negative_far"] = (big * 3.0 - [5e3, 0.0, 2e3]).astype(F)
    out["huge_coords"] = (rng.uniform(-5, 5, (300, 3)) + [1e5, -1e5, 3e4]).astype(F)   # coordinates quantised to 1/128
    mixed = rng.normal(0, 1e-4, (300, 3))
    mixed[:6] = rng.uniform(-100, 100, (6, 3))                                        # a dense knot + outliers 100 m away
    out["mixed_scale"] = mixed.astype(F)
    out["with_nonfinite"] = np.where(rng.uniform(size=(300, 3)) < 0.05, np.nan, rng.uniform(-1, 1, (300, 3))).astype(F)
    surf = rng.uniform(-0.06, 0.06, (600, 2))
    out["surface"] = np.column_stack([surf, 0.7 + 0.02 * np.sin(40 * surf[:, 0]) * np.cos(35 * surf[:, 1])]).astype(F)
    return out


def queries(rng, tgt):
    tgt = tgt[np.isfinite(tgt).all(1)]
    n = len(tgt)
    lo, hi = tgt.min(0), tgt.max(0)
    ext = np.maximum(hi - lo, np.spacing(np.abs(hi).max().astype(F)) * 4)
    pick = rng.integers(0, n, 160)
    a, b = tgt[rng.integers(0, n, 160)], tgt[rng.integers(0, n, 160)]
    qs = [
        tgt[pick],                                                        # on top of target points
        ((a.astype(np.float64) + b) * 0.5).astype(F),                     # midpoints: exact ties on lattices
        np.nextafter(tgt[pick], F(np.inf)), np.nextafter(tgt[pick], F(-np.inf)),   # one ulp off a target point
        (lo + rng.uniform(0, 1, (160, 3)) * ext).astype(F),               # inside the box
        (lo + rng.uniform(-2, 3, (160, 3)) * ext).astype(F),              # around the box
        (lo + rng.uniform(-40, 41, (60, 3)) * ext).astype(F),             # far outside
        np.array([lo, hi, (lo + hi) / 2, [lo[0], hi[1], lo[2]], [hi[0], lo[1], hi[2]]], F),  # corners
        (lo + rng.integers(0, 9, (100, 3)) * (ext / 8)).astype(F),        # a coarse lattice over the box (cell edges for some h)
    ]
    return np.ascontiguousarray(np.concatenate(qs).astype(F))


OCCUPANCIES = [0.25, 1.0, 3.5, 16.0]


def grid_nn_seeded(hc, tgt, q, sq, occupancy, limit=np.inf):  # noqa: F811
    tgt = np.ascontiguousarray(tgt, F)
    idx = np.empty(len(q), np.int32)
    d2 = np.empty(len(q), F)
    hc.hc_grid_nn_seeded(tgt.ctypes.data, len(tgt), tgt.strides[0], q.ctypes.data, sq.ctypes.data, len(q), occupancy, limit,
                         idx.ctypes.data, d2.ctypes.data)
    return idx, d2


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_ring_search_is_exact_on_adversarial_clouds(hc, oracle, seed):  # noqa: F811
    rng = np.random.default_rng(1000 + seed)
    for name, tgt in clouds(rng).items():
        q = queries(rng, tgt)
        bi, bd = oracle.nn_bruteforce(tgt, q)
        for occ in OCCUPANCIES:
            gi, gd, _ = grid_nn(hc, tgt, q, occupancy=occ)
            assert np.array_equal(gd, bd), (name, occ, np.flatnonzero(gd != bd)[:5])
            assert np.array_equal(gi, bi), (name, occ, np.flatnonzero(gi != bi)[:5])


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_warm_ball_search_is_exact_on_adversarial_clouds(hc, oracle, seed):  # noqa: F811
    """whatever the candidate: the true match, a neighbour in index order, a random point"""
    rng = np.random.default_rng(2000 + seed)
    for name, tgt in clouds(rng).items():
        q = queries(rng, tgt)
        bi, bd = oracle.nn_bruteforce(tgt, q)
        ok = np.flatnonzero(np.isfinite(tgt).all(1))  # candidates are previous matches: always finite points
        for occ in OCCUPANCIES:
            for kind, prev in (("true", bi), ("next", ok[(np.searchsorted(ok, bi) + 1) % len(ok)]), ("random", rng.choice(ok, len(q)))):
                gi, gd = grid_nn_warm(hc, tgt, q, prev.astype(np.int32), occ)
                assert np.array_equal(gd, bd), (name, occ, kind, np.flatnonzero(gd != bd)[:5])
                assert np.array_equal(gi, bi), (name, occ, kind, np.flatnonzero(gi != bi)[:5])


def grid_nn_warm_graph(hc, tgt, q, prev, occupancy, limit=np.inf, mode=0):  # noqa: F811
    """mode 0: every row scanned; 1: rows that cannot certify skipped (the warm launches of a batch); 2: greedy descent
    from prev, then the ball search (launch 0's candidate and its verification); 3: mode 1 with the flatness certificate
    (and the look at four neighbours for every other query)"""
    tgt = np.ascontiguousarray(tgt, F)
    idx = np.empty(len(q), np.int32)
    d2 = np.empty(len(q), F)
    hc.hc_grid_nn_warm_graph(tgt.ctypes.data, len(tgt), tgt.strides[0], q.ctypes.data, len(q), q.strides[0], occupancy,
                             prev.ctypes.data, limit, idx.ctypes.data, d2.ctypes.data, mode)
    return idx, d2


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_knn_graph_warm_search_is_exact_on_adversarial_clouds(hc, oracle, seed):  # noqa: F811
    """csrc/nn_graph.cuh: certificate from the candidate's k-NN row, greedy steps, grid walk as the last resort — exact
    whatever the candidate (lattices full of ties, duplicates: more equal points than a row holds, clouds smaller than a row)"""
    rng = np.random.default_rng(7000 + seed)
    for name, tgt in clouds(rng).items():
        q = queries(rng, tgt)
        bi, bd = oracle.nn_bruteforce(tgt, q)
        ok = np.flatnonzero(np.isfinite(tgt).all(1))
        for occ in (1.0, 5.0):
            for kind, prev in (("true", bi), ("next", ok[(np.searchsorted(ok, bi) + 1) % len(ok)]), ("random", rng.choice(ok, len(q)))):
                for mode in (0, 1, 2, 3):
                    gi, gd = grid_nn_warm_graph(hc, tgt, q, np.ascontiguousarray(prev, np.int32), occ, mode=mode)
                    assert np.array_equal(gd, bd), (name, occ, kind, mode, np.flatnonzero(gd != bd)[:5])
                    assert np.array_equal(gi, bi), (name, occ, kind, mode, np.flatnonzero(gi != bi)[:5])


@pytest.mark.parametrize("seed", [0, 1])
def test_flatness_certificate_is_exact_on_sheets_steps_and_creases(hc, oracle, seed):  # noqa: F811
    """csrc/nn_graph.cuh : knn_aux_of / flat() — the geometry the certificate is made for and the geometry that must stop
    it: a noisy sheet, a sheet with a second one a few spacings behind half of it (depth step), a crease, a sheet with a
    lone point hovering above it; queries hover 0.1-4 spacings off the sheet around true, neighbouring and random matches"""
    rng = np.random.default_rng(9100 + seed)
    a = 1e-3  # spacing
    n = 45
    gx, gy = np.meshgrid(np.arange(n), np.arange(n))
    base = np.column_stack([gx.ravel(), gy.ravel(), np.zeros(n * n)]).astype(np.float64) * a
    base[:, :2] += rng.uniform(-0.3, 0.3, (n * n, 2)) * a
    sheets = {}
    noisy = base.copy()
    noisy[:, 2] += rng.normal(0, 0.05 * a, n * n)
    sheets["noisy_sheet"] = noisy
    step = np.concatenate([noisy, noisy[noisy[:, 0] > 0.5 * n * a] + [0.0, 0.0, -2.5 * a]])
    sheets["depth_step"] = step
    crease = noisy.copy()
    crease[:, 2] += np.abs(crease[:, 0] - 0.5 * n * a) * 0.8
    sheets["crease"] = crease
    sheets["hovering_point"] = np.concatenate([noisy, [[0.5 * n * a, 0.5 * n * a, 1.3 * a]]])
    for name, cloud in sheets.items():
        tgt = (cloud + [0.1, -0.2, 0.7]).astype(F)
        pick = rng.integers(0, len(tgt), 1500)
        off = np.column_stack([rng.normal(0, 0.4 * a, (1500, 2)), rng.uniform(-4 * a, 4 * a, 1500) * rng.choice([0.03, 0.3, 1.0], 1500)])
        q = np.ascontiguousarray((tgt[pick].astype(np.float64) + off).astype(F))
        bi, bd = oracle.nn_bruteforce(tgt, q)
        for occ in (2.0, 5.0):
            for kind, prev in (("true", bi), ("picked", pick), ("random", rng.integers(0, len(tgt), len(q)))):
                gi, gd = grid_nn_warm_graph(hc, tgt, q, np.ascontiguousarray(prev, np.int32), occ, mode=3)
                assert np.array_equal(gd, bd), (name, occ, kind, np.flatnonzero(gd != bd)[:5])
                assert np.array_equal(gi, bi), (name, occ, kind, np.flatnonzero(gi != bi)[:5])


@pytest.mark.parametrize("seed", [0, 1])
def test_warm_search_with_a_rejection_limit_is_exact_where_it_matters(hc, oracle, seed):  # noqa: F811
    """limit_d2 (the max-correspondence-distance cut): matches within the limit are exact, the others stay beyond it"""
    rng = np.random.default_rng(3000 + seed)
    for name, tgt in clouds(rng).items():
        q = queries(rng, tgt)
        bi, bd = oracle.nn_bruteforce(tgt, q)
        finite = bd[np.isfinite(bd) & (bd > 0)]
        if finite.size == 0:
            continue
        lim = F(np.median(finite))
        prev = rng.choice(np.flatnonzero(np.isfinite(tgt).all(1)), len(q)).astype(np.int32)
        for occ in (1.0, 3.5):
            gi, gd = grid_nn_warm(hc, tgt, q, prev, occ, float(lim))
            acc = bd <= lim
            assert np.array_equal(gd[acc], bd[acc]) and np.array_equal(gi[acc], bi[acc]), (name, occ)
            assert (gd[~acc] > lim).all(), (name, occ)


@pytest.mark.parametrize("seed", [0, 1])
def test_seeded_search_is_exact_on_adversarial_clouds(hc, oracle, seed):  # noqa: F811
    """the first-iteration scheme: a query seeded with the match of a nearby (or not so nearby) query"""
    rng = np.random.default_rng(4000 + seed)
    for name, tgt in clouds(rng).items():
        q = queries(rng, tgt)
        ext = np.maximum(np.nanmax(tgt, 0) - np.nanmin(tgt, 0), F(1e-6))
        for spread in (1e-3, 0.05, 2.0):
            sq = np.ascontiguousarray((q + rng.normal(0, 1, q.shape) * ext * spread).astype(F))
            bi, bd = oracle.nn_bruteforce(tgt, q)
            for occ in (0.25, 3.5):
                gi, gd = grid_nn_seeded(hc, tgt, q, sq, occ)
                assert np.array_equal(gd, bd), (name, occ, spread, np.flatnonzero(gd != bd)[:5])
                assert np.array_equal(gi, bi), (name, occ, spread, np.flatnonzero(gi != bi)[:5])


def grid_nn_warm_upfront(hc, tgt, q, prev, occupancy, limit=np.inf, rows3=0):  # noqa: F811
    import ctypes as C

    tgt = np.ascontiguousarray(tgt, F)
    prev = np.ascontiguousarray(prev, np.int32)
    idx = np.empty(len(q), np.int32)
    d2 = np.empty(len(q), F)
    vp, sz = C.c_void_p, C.c_size_t
    hc.hc_grid_nn_warm_upfront.argtypes = [vp, sz, sz, vp, sz, sz, C.c_float, vp, C.c_float, vp, vp, C.c_int]
    hc.hc_grid_nn_warm_upfront(tgt.ctypes.data, len(tgt), tgt.strides[0], q.ctypes.data, len(q), q.strides[0], occupancy,
                               prev.ctypes.data, limit, idx.ctypes.data, d2.ctypes.data, rows3)
    return idx, d2


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_staged_upfront_warm_search_is_exact_too(hc, oracle, seed):  # noqa: F811
    """csrc/nn_upfront.cuh (staged for the next round, not in any kernel yet): row bounds of the initial ball fetched up
    front, chords not narrowed — must return what the row-after-row walk returns, on the same adversarial inputs and on
    an ICP-like surface case where most balls are small (the 2 x 2 path) and some are not (the fallback)."""
    from pose_estimation_b200.testing import synth

    rng = np.random.default_rng(5000 + seed)
    cases = dict(clouds(rng))
    prob = synth.make_c1(5000, seed=40 + seed)
    cases["icp_like"] = np.ascontiguousarray(prob.target[:, :3])
    for name, tgt in cases.items():
        q = queries(rng, tgt)
        if name == "icp_like":
            q = np.ascontiguousarray(np.concatenate([q, prob.source[:, :3]]))
        bi, bd = oracle.nn_bruteforce(tgt, q)
        ok = np.flatnonzero(np.isfinite(tgt).all(1))
        for occ in OCCUPANCIES:
            for kind, prev in (("true", bi), ("next", ok[(np.searchsorted(ok, bi) + 1) % len(ok)]), ("random", rng.choice(ok, len(q)))):
                for rows3 in (0, 1):  # boxes up to 2 x 2 rows, up to 3 x 3 rows
                    gi, gd = grid_nn_warm_upfront(hc, tgt, q, prev, occ, rows3=rows3)
                    assert np.array_equal(gd, bd), (name, occ, kind, rows3, np.flatnonzero(gd != bd)[:5])
                    assert np.array_equal(gi, bi), (name, occ, kind, rows3, np.flatnonzero(gi != bi)[:5])
        lim = F(np.median(bd[np.isfinite(bd)])) if np.isfinite(bd).any() else F(1.0)
        gi, gd = grid_nn_warm_upfront(hc, tgt, q, rng.choice(ok, len(q)), 3.5, float(lim))
        acc = bd <= lim
        assert np.array_equal(gd[acc], bd[acc]) and np.array_equal(gi[acc], bi[acc]), name
        assert (gd[~acc] > lim).all(), name


def _bounded(hc, tgt, q, bound_d2, occupancy, limit=np.inf):
    import ctypes as C

    tgt = np.ascontiguousarray(tgt, F)
    bound_d2 = np.ascontiguousarray(bound_d2, F)
    idx = np.empty(len(q), np.int32)
    d2 = np.empty(len(q), F)
    vp, sz = C.c_void_p, C.c_size_t
    hc.hc_grid_nn_bounded.argtypes = [vp, sz, sz, vp, sz, sz, C.c_float, vp, C.c_float, vp, vp]
    hc.hc_grid_nn_bounded(tgt.ctypes.data, len(tgt), tgt.strides[0], q.ctypes.data, len(q), q.strides[0], occupancy,
                          bound_d2.ctypes.data, limit, idx.ctypes.data, d2.ctypes.data)
    return idx, d2


@pytest.mark.parametrize("seed", [0, 1])
def test_staged_bound_only_search_is_exact_for_any_true_bound(hc, oracle, seed):  # noqa: F811
    """csrc/nn_upfront.cuh : grid_nn_bounded_upfront (staged): no candidate, only an upper bound of the nearest
    neighbour's squared distance — exact for the tightest possible bound (the distance itself) and for loose ones;
    a bound cut by the rejection limit reports 'nothing' exactly for the queries brute force rejects."""
    rng = np.random.default_rng(6000 + seed)
    for name, tgt in clouds(rng).items():
        q = queries(rng, tgt)
        bi, bd = oracle.nn_bruteforce(tgt, q)
        for occ in (0.25, 3.5, 16.0):
            for k in (0.0, 1e-6, 0.5, 10.0):
                bound = (bd.astype(np.float64) * (1.0 + k)).astype(F)
                bound = np.maximum(bound, bd)  # (rounding of the product must not undercut the distance)
                gi, gd = _bounded(hc, tgt, q, bound, occ)
                assert np.array_equal(gd, bd), (name, occ, k, np.flatnonzero(gd != bd)[:5])
                assert np.array_equal(gi, bi), (name, occ, k, np.flatnonzero(gi != bi)[:5])
        pos = bd[bd > 0]
        if pos.size:
            lim = F(np.median(pos))
            gi, gd = _bounded(hc, tgt, q, np.full(len(q), np.inf, F), 3.5, float(lim))
            acc = bd <= lim
            assert np.array_equal(gi[acc], bi[acc]) and np.array_equal(gd[acc], bd[acc]), name
            assert (gi[~acc] == -1).all(), name


@pytest.mark.parametrize("limit", [np.inf, 0.02 ** 2, 0.002 ** 2])
def test_staged_bound_only_search_along_an_icp_trajectory(hc, oracle, limit):  # noqa: F811
    """The invariant the kernel would rely on: d_old + |movement| (inflated by warm_bound_d2) bounds the new nearest
    distance.  Query sets = the working cloud of a simulated ICP (float32 coordinates, as the kernel stores them),
    including a large first step; every warm search must equal brute force wherever the match is accepted."""
    import ctypes as C

    from scipy.spatial import cKDTree

    from pose_estimation_b200.testing import synth

    prob = synth.make_c1(6000, seed=61)
    tgt = np.ascontiguousarray(prob.target[:, :3], F)
    tree = cKDTree(tgt.astype(np.float64))
    work = prob.source[:, :3].astype(np.float64)
    sets = []
    for it in range(12):
        sets.append(work.astype(F))
        d, idx = tree.query(work)
        src, dst = work, tgt[idx].astype(np.float64)
        cs, cd = src.mean(0), dst.mean(0)
        U, _, Vt = np.linalg.svd((dst - cd).T @ (src - cs))
        R = U @ np.diag([1.0, 1.0, np.sign(np.linalg.det(U @ Vt))]) @ Vt
        work = work @ R.T + (cd - R @ cs)
    q = np.ascontiguousarray(np.stack(sets))
    steps, nq = q.shape[0], q.shape[1]
    out_idx = np.empty((steps, nq), np.int32)
    out_d2 = np.empty((steps, nq), F)
    vp, sz = C.c_void_p, C.c_size_t
    hc.hc_bounded_trajectory.argtypes = [vp, sz, sz, vp, sz, sz, C.c_float, C.c_float, vp, vp]
    hc.hc_bounded_trajectory(tgt.ctypes.data, len(tgt), tgt.strides[0], q.ctypes.data, nq, steps, 3.5, limit,
                             out_idx.ctypes.data, out_d2.ctypes.data)
    accepted_warm = 0
    for k in range(steps):
        bi, bd = oracle.nn_bruteforce(tgt, q[k])
        acc = bd <= limit
        assert np.array_equal(out_idx[k][acc], bi[acc]) and np.array_equal(out_d2[k][acc], bd[acc]), k
        assert ((out_idx[k][~acc] == -1) | (out_d2[k][~acc] > limit)).all(), k
        accepted_warm += int(acc.sum()) if k else 0
    assert accepted_warm > 0


def test_staged_device_code_compiles_for_sm_100a(tmp_path):
    """grid_nn_bounded_upfront<2 / 3> + warm_bound_d2 are templates nothing in the library instantiates yet: compile them
    as device code so that the next round starts from something nvcc accepts (no spills expected at 48 registers)."""
    import subprocess
    from pathlib import Path

    src = Path(__file__).resolve().parent / "host" / "staged_compile.cu"
    r = subprocess.run(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O3", "-fmad=false",
                        "--expt-relaxed-constexpr", "-Xptxas", "-v", "-c", str(src), "-o", str(tmp_path / "staged.o")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert r.stderr.count("0 bytes spill stores") == 2, r.stderr
