"""Plane RANSAC (pcl::SACSegmentation, SACMODEL_PLANE + SAC_RANSAC; the plane fit of the reference's remove_planes,
pose_estimation/src/pose_estimation.cpp:285-297).

CPU part: the oracle's Mersenne twister against known answers, the oracle's sequential loop against an independent
numpy restatement (sample stream from numpy's own MT19937, float32 arithmetic spelled out), the optimised
coefficients against a float64 PCA.  GPU part: the library (all samples drawn up front, one counting pass, replayed
loop) against the oracle's sequential loop, through the C ABI."""
import numpy as np
import pytest

from oracle import sac_params
from pose_estimation_b200.testing import synth

F = np.float32


@pytest.fixture(scope="module")
def pcl():
    from pose_estimation_b200 import pcl as m

    return m


@pytest.fixture(scope="module")
def ctx(pcl):
    c = pcl.Context(0)
    yield c
    c.close()


def plane_cloud(n_plane=1500, n_other=700, seed=0, noise=2e-4, nan=0):
    """A tilted plane patch ~0.75 m in front of the camera + clutter, float32 (n, 4)."""
    rng = np.random.default_rng(seed)
    nrm = synth.rotation_about([1.0, 0.3, 0.0], np.deg2rad(7.0)) @ np.array([0.0, 0.0, 1.0])
    u = np.cross(nrm, [1.0, 0.0, 0.0])
    u /= np.linalg.norm(u)
    v = np.cross(nrm, u)
    ab = rng.uniform(-0.3, 0.3, (n_plane, 2))
    plane = np.array([0.0, 0.0, 0.75]) + ab[:, :1] * u + ab[:, 1:] * v + rng.normal(0, noise, (n_plane, 1)) * nrm
    other = rng.uniform([-0.3, -0.3, 0.55], [0.3, 0.3, 0.72], (n_other, 3))
    pts = np.concatenate([plane, other], 0)
    pts = pts[rng.permutation(len(pts))]
    out = synth.xyz4(pts.astype(np.float32))
    if nan:
        out[rng.choice(len(out), nan, replace=False), :3] = np.nan
    return out


def numpy_ransac(pts, threshold, max_iterations, probability=0.99, seed=12345):
    """Independent restatement of RandomSampleConsensus::computeModel for the plane model (no optimisation):
    -> (found, coefficients, iterations)."""
    n = len(pts)
    raw = iter(np.random.RandomState(seed).randint(0, 2**32, size=40000, dtype=np.uint64))  # genrand_int32 stream
    shuffled = np.arange(n)
    xyz = pts[:, :3].astype(F)

    def dot4(mc, p):
        return F(F(F(mc[0] * p[0]) + F(mc[1] * p[1])) + F(mc[2] * p[2])) + F(mc[3] * F(1.0))

    iterations, n_best, k, best = 0, -(2**31 - 1), 1.0, None
    log_p = np.log(1.0 - probability)
    while iterations < k:
        good = False
        for _ in range(1000):
            for i in range(3):
                r = int(next(raw)) >> 1
                j = i + r % (n - i)
                shuffled[i], shuffled[j] = shuffled[j], shuffled[i]
            p0, p1, p2 = xyz[shuffled[0]], xyz[shuffled[1]], xyz[shuffled[2]]
            with np.errstate(all="ignore"):
                q = (p1 - p0) / (p2 - p0)
            if not (q[0] == q[1] and q[2] == q[1]):
                good = True
                break
        if not good:
            break
        a, b = p1 - p0, p2 - p0
        mc = np.array([F(a[1] * b[2]) - F(a[2] * b[1]), F(a[2] * b[0]) - F(a[0] * b[2]), F(a[0] * b[1]) - F(a[1] * b[0]), 0], F)
        nn = np.sqrt(F(F(F(mc[0] * mc[0]) + F(mc[1] * mc[1])) + F(mc[2] * mc[2])) + F(mc[3] * mc[3]))
        mc = (mc / nn).astype(F)
        mc[3] = F(-1.0) * dot4(mc, p0)
        d = (mc[0] * xyz[:, 0] + mc[1] * xyz[:, 1]) + mc[2] * xyz[:, 2] + mc[3] * F(1.0)  # float32 array ops, same order
        cnt = int(np.count_nonzero(np.abs(d).astype(np.float64) < threshold))
        if cnt > n_best:
            n_best, best = cnt, mc.copy()
            w = n_best / n
            pno = min(max(1.0 - w**3, np.finfo(np.float64).eps), 1.0 - np.finfo(np.float64).eps)
            k = log_p / np.log(pno)
        iterations += 1
        if iterations > max_iterations:
            break
    return best is not None, best, iterations


def test_mersenne_twister_known_answers(oracle):
    assert oracle.mt19937_nth(5489, 10000) == 4123659995  # the value the C++ standard requires of std::mt19937
    raw = np.random.RandomState(12345).randint(0, 2**32, size=700, dtype=np.uint64)
    assert [oracle.mt19937_nth(12345, i) for i in (1, 2, 3, 624, 625, 700)] == [int(raw[i - 1]) for i in (1, 2, 3, 624, 625, 700)]


@pytest.mark.parametrize("seed,threshold,max_it", [(0, 4e-4, 100), (1, 1e-4, 100), (2, 1e-3, 7), (3, 2e-4, 50)])
def test_oracle_loop_against_numpy_restatement(oracle, seed, threshold, max_it):
    pts = plane_cloud(seed=seed)
    found, coeff, inl, its = oracle.sac_plane(pts, sac_params(threshold, max_it, optimize=False))
    ok, ref, ref_its = numpy_ransac(pts, threshold, max_it)
    assert found and ok and its == ref_its
    assert coeff.tobytes() == ref.astype(F).tobytes()
    d = (coeff[0] * pts[:, 0] + coeff[1] * pts[:, 1]) + coeff[2] * pts[:, 2] + coeff[3]
    assert np.array_equal(inl, np.flatnonzero(np.abs(d).astype(np.float64) < threshold))


def test_oracle_optimised_coefficients_are_the_pca_plane(oracle):
    pts = plane_cloud(seed=5, n_plane=4000)
    found, coeff, inl, _ = oracle.sac_plane(pts, sac_params(4e-4, 100, optimize=True), wide_accum=True)
    assert found and len(inl) > 1000
    # least-squares plane of the inliers of the UNoptimised model (float64)
    _, c0, inl0, _ = oracle.sac_plane(pts, sac_params(4e-4, 100, optimize=False))
    q = pts[inl0, :3].astype(np.float64)
    cen = q.mean(0)
    w, v = np.linalg.eigh(np.cov((q - cen).T))
    nrm = v[:, 0] * np.sign(v[:, 0] @ coeff[:3])
    assert np.degrees(np.arccos(np.clip(nrm @ coeff[:3], -1, 1))) < 2e-3  # eigen33 in float on a thin slab
    assert abs(coeff[3] + nrm @ cen) < 2e-5
    # PCL 1.10's float sums stay close (its noise is what DESIGN.md documents)
    _, cf, _, _ = oracle.sac_plane(pts, sac_params(4e-4, 100, optimize=True), wide_accum=False)
    assert np.degrees(np.arccos(np.clip(cf[:3] @ coeff[:3], -1, 1))) < 0.05


def test_oracle_degenerate_inputs(oracle):
    found, coeff, inl, its = oracle.sac_plane(np.zeros((2, 4), F), sac_params(1e-3, 10))
    assert not found and not coeff.any() and len(inl) == 0 and its == 0
    line = np.zeros((50, 4), F)
    line[:, 0] = np.arange(50)  # y = z = 0 everywhere: every sample is collinear (0/0 compares false -> "good" ...)
    line[:, 3] = 1
    found, coeff, inl, its = oracle.sac_plane(line, sac_params(1e-3, 10, optimize=False))
    # (p1-p0)/(p2-p0) = (r, nan, nan): nan != nan, so PCL accepts the sample and fits a NaN plane with no inliers
    assert found and np.isnan(coeff).all() and len(inl) == 0


# ------------------------------------------------------------------------------------------------------------
# GPU: the library against the oracle
# ------------------------------------------------------------------------------------------------------------
def _segment(pcl, ctx, pts, threshold, max_it, optimize, probability=0.99):
    seg = pcl.SACSegmentation(ctx)
    seg.setModelType(pcl.SACSegmentation.SACMODEL_PLANE)
    seg.setMethodType(pcl.SACSegmentation.SAC_RANSAC)
    seg.setOptimizeCoefficients(optimize)
    seg.setDistanceThreshold(threshold)
    seg.setMaxIterations(max_it)
    seg.setProbability(probability)
    seg.setInputCloud(pts)
    inl, coeff = seg.segment()
    return inl, coeff, seg.iterations_


@pytest.mark.gpu
@pytest.mark.parametrize("seed,threshold,max_it,nan", [(0, 4e-4, 100, 0), (1, 1e-4, 100, 0), (2, 1e-3, 7, 0), (3, 2e-4, 50, 60),
                                                      (4, 1e-4, 0, 0), (5, 5e-4, 400, 0)])
def test_plane_ransac_matches_the_sequential_loop(pcl, ctx, oracle, seed, threshold, max_it, nan):
    pts = plane_cloud(seed=seed, nan=nan)
    # without optimisation every number is PCL's float arithmetic: bit-exact
    inl, coeff, its = _segment(pcl, ctx, pts, threshold, max_it, False)
    found, rc, rinl, rits = oracle.sac_plane(pts, sac_params(threshold, max_it, optimize=False))
    if max_it == 0:  # max_skip = 10 * max_iterations = 0: PCL's loop never runs, there is no model
        assert not found and len(inl) == 0 and len(coeff) == 0 and its == rits == 0
        return
    assert found and its == rits and coeff.tobytes() == rc.tobytes() and np.array_equal(inl, rinl)
    # with optimisation the inlier moments are double sums in a different order: same plane to ~1e-7
    inl, coeff, its = _segment(pcl, ctx, pts, threshold, max_it, True)
    found, rc, rinl, rits = oracle.sac_plane(pts, sac_params(threshold, max_it, optimize=True), wide_accum=True)
    assert found and its == rits
    assert np.allclose(coeff, rc, rtol=0, atol=2e-6)
    assert len(np.setxor1d(inl, rinl)) <= max(2, len(rinl) // 2000)  # only points within an ulp of the threshold may flip


@pytest.mark.gpu
def test_plane_ransac_edge_cases(pcl, ctx, oracle):
    inl, coeff, its = _segment(pcl, ctx, np.zeros((2, 4), F), 1e-3, 10, True)
    assert len(inl) == 0 and len(coeff) == 0 and its == 0
    inl, coeff, its = _segment(pcl, ctx, np.zeros((0, 4), F), 1e-3, 10, True)
    assert len(inl) == 0 and len(coeff) == 0
    # threshold 0: no point is ever an inlier, the first sample's plane is returned un-optimised
    pts = plane_cloud(seed=9)
    inl, coeff, its = _segment(pcl, ctx, pts, 0.0, 20, True)
    found, rc, rinl, rits = oracle.sac_plane(pts, sac_params(0.0, 20, optimize=True))
    assert found and len(inl) == 0 == len(rinl) and coeff.tobytes() == rc.tobytes() and its == rits
    # exact duplicates of one point: (p1 - p0) / (p2 - p0) is 0/0 in every lane, NaN never compares equal, so PCL's
    # collinearity test lets the sample through and the "model" is a NaN plane without inliers — reproduced as is
    dup = np.tile(np.array([[0.1, 0.2, 0.7, 1.0]], F), (100, 1))
    inl, coeff, its = _segment(pcl, ctx, dup, 1e-3, 10, True)
    found, rc, rinl, rits = oracle.sac_plane(dup, sac_params(1e-3, 10))
    assert found and np.isnan(rc).all() and np.isnan(coeff).all() and len(inl) == 0 == len(rinl) and its == rits
    # two distinct points repeated: p1 - p0 = 0 against p2 - p0 != 0 gives 0 in every lane -> truly "collinear":
    # 1000 draws without a good sample, no model
    two = np.tile(np.array([[0.1, 0.2, 0.7, 1.0], [0.1, 0.2, 0.7, 1.0], [0.3, 0.1, 0.8, 1.0]], F), (40, 1))
    inl, coeff, its = _segment(pcl, ctx, two, 1e-3, 10, True)
    found, rc, rinl, rits = oracle.sac_plane(two, sac_params(1e-3, 10))
    assert found == (len(coeff) == 4) and its == rits and len(inl) == len(rinl)
    seg = pcl.SACSegmentation(ctx)
    with pytest.raises(pcl.PebError) as e:
        seg.setModelType(5)  # SACMODEL_SPHERE
    assert e.value.code == -6
    with pytest.raises(pcl.PebError):
        seg.setMethodType(1)  # SAC_LMEDS


@pytest.mark.gpu
def test_plane_ransac_full_scene_and_band_removal(pcl, ctx, oracle):
    """The reference's remove_planes on the NaN-free 1944 x 1200 scene: threshold 0.0001, 100 iterations, optimised
    coefficients, then the 5 mm band removal with those coefficients (pose_estimation.cpp:285-333)."""
    rng = np.random.default_rng(11)
    surf = synth.Surface(11)
    scene = synth.render_scene(surf, synth.default_gt_pose(rng), rng)
    pf = pcl.ScenePrefilter(ctx)
    pf.setInputCloud(scene)
    cloud = pf.filter()  # NaN removal
    assert len(cloud) > 2_000_000
    inl, coeff, its = _segment(pcl, ctx, cloud, 1e-4, 100, True)
    found, rc, rinl, rits = oracle.sac_plane(cloud, sac_params(1e-4, 100, optimize=True), wide_accum=True)
    assert found and its == rits and np.allclose(coeff, rc, rtol=0, atol=2e-6)
    assert len(np.setxor1d(inl, rinl)) <= len(rinl) // 2000
    # the background plane of the synthetic scene: z = 0.75 at the optical axis, tilted 5 degrees
    n_true = synth.rotation_about([1.0, 0.3, 0.0], np.deg2rad(5.0)) @ np.array([0.0, 0.0, 1.0])
    assert np.degrees(np.arccos(abs(float(coeff[:3] @ n_true)))) < 0.05
    pf2 = pcl.ScenePrefilter(ctx)
    pf2.setInputCloud(cloud)
    pf2.addPlane(*[float(v) for v in coeff])
    kept = pf2.filter()
    from oracle import prefilter_params
    ref = oracle.scene_prefilter(cloud, prefilter_params(planes=[rc]))
    assert abs(len(kept) - len(ref)) <= 2 and len(kept) < 0.5 * len(cloud)


@pytest.mark.gpu
def test_scene_preparation_in_one_call_equals_the_steps(pcl, ctx):
    """peb_scene_prepare (create_surface_match_pc in one call) against the same stages called one after the other."""
    rng = np.random.default_rng(12)
    surf = synth.Surface(12)
    gt = synth.default_gt_pose(rng)
    scene = synth.render_scene(surf, gt, rng, 972, 600)
    centre, radius = gt[:3, 3], 0.35
    got, planes = pcl.create_surface_match_pc(scene, ctx, filter_pose=centre, filter_radius=radius, num_planes=2, leaf=0.002)
    pf = pcl.ScenePrefilter(ctx)
    pf.setInputCloud(scene)
    pf.setSphereFilter(centre, radius)
    cloud = pf.filter()
    ref_planes = []
    for _ in range(2):
        inl, coeff, _ = _segment(pcl, ctx, cloud, 1e-4, 100, True)
        ref_planes.append(coeff)
        band = pcl.ScenePrefilter(ctx)
        band.setInputCloud(cloud)
        band.addPlane(*[float(v) for v in coeff])
        cloud = band.filter()
    vg = pcl.VoxelGrid(ctx)
    vg.setInputCloud(cloud)
    vg.setLeafSize(0.002)
    ref = vg.filter()
    assert np.array_equal(planes, np.stack(ref_planes)) and got.tobytes() == ref.tobytes() and 1000 < len(got) < len(scene)
    # no planes, no sphere, no VoxelGrid: plain NaN removal
    only, none = pcl.create_surface_match_pc(scene, ctx)
    assert len(none) == 0 and np.array_equal(only, scene[np.isfinite(scene[:, :3]).all(1)])
