"""The binned warm launches of a batch (pose_estimation_b200/csrc/icp.cu : icp_iteration_binned_kernel, knob `warm_bin`):
the queries of a block are sorted by the number of grid rows their ball search will walk before the warps search them.
Queries, searches and the order of every thread's double sums are those of the plain kernel, so every record must be
byte-identical with the knob on and off — point-to-point and point-to-plane, hypotheses that stop at different
iterations, clouds with non-finite points, sources smaller than a tile, rejector on.
"""
import numpy as np
import pytest

from pose_estimation_b200.testing import synth

pytestmark = [pytest.mark.gpu]


@pytest.fixture(scope="module")
def pcl():
    from pose_estimation_b200 import pcl as m

    return m


@pytest.fixture(scope="module")
def scene_small(oracle):
    return synth.make_c2(scale=0.25, downsample=lambda p, leaf: oracle.voxel_grid(p, leaf)[0])


def _run(pcl, source, target, cls, normals, guesses, warm_bin, opts=(), **params):
    from oracle import default_params

    c = pcl.Context(0)
    c.set_int("warm_bin", warm_bin)
    for k, v in opts:
        c.set_int(k, v)
    icp = cls(c)
    icp.setInputSource(source)
    icp.setInputTarget(target, normals)
    prm = default_params(**params)
    for name, _ in prm._fields_:
        if name != "estimator":
            setattr(icp.params, name, getattr(prm, name))
    res = icp.alignBatch(guesses)
    out = [bytes(r) for r in res]
    c.close()
    return out


def test_binned_warm_launches_never_change_results(pcl, scene_small):
    p = scene_small
    rng = np.random.default_rng(15)
    guesses = np.stack([synth.perturb_pose(p.gt_pose, rng, 5.0, 0.006) for _ in range(200)])
    kw = dict(max_iterations=25, max_corr_dist=0.02, abs_mse_threshold=-1.0)
    a = _run(pcl, p.source, p.target, pcl.IterativeClosestPoint, None, guesses, 0, **kw)
    b = _run(pcl, p.source, p.target, pcl.IterativeClosestPoint, None, guesses, 1, **kw)
    assert a == b
    # other launch shapes: one chain without per-hypothesis dependencies, many small blocks
    for opts in ((("batch_streams", 1), ("flag_deps", 0)), (("blocks_factor", 96),), (("blocks_factor", 4),)):
        assert _run(pcl, p.source, p.target, pcl.IterativeClosestPoint, None, guesses, 1, opts, **kw) == \
               _run(pcl, p.source, p.target, pcl.IterativeClosestPoint, None, guesses, 0, opts, **kw)


def test_binned_point_to_plane_rejector_and_criteria(pcl, scene_small):
    p = scene_small
    c = pcl.Context(0)
    ne = pcl.NormalEstimation(c)
    ne.setInputCloud(p.target)
    ne.setKSearch(12)
    normals = ne.compute()
    c.close()
    rng = np.random.default_rng(16)
    guesses = np.stack([synth.perturb_pose(p.gt_pose, rng, 4.0, 0.004) for _ in range(40)])
    kw = dict(max_iterations=30, max_corr_dist=0.01, transformation_epsilon=1e-9, rejector_max_dist=0.008)
    a = _run(pcl, p.source, p.target, pcl.IterativeClosestPointWithNormals, normals, guesses, 0, **kw)
    b = _run(pcl, p.source, p.target, pcl.IterativeClosestPointWithNormals, normals, guesses, 1, **kw)
    assert a == b
    assert len({r[-32:] for r in a}) > 1  # (the records differ between hypotheses: the comparison is not vacuous)


def test_binned_nonfinite_points_far_hypotheses_and_tiny_sources(pcl, scene_small):
    p = scene_small
    rng = np.random.default_rng(17)
    src = p.source.copy()
    src[rng.choice(len(src), 300, replace=False), rng.integers(0, 3, 300)] = np.nan
    src[rng.choice(len(src), 50, replace=False), 0] = np.inf
    # a third of the hypotheses start far away: their points have no match within max_corr_dist and search cold
    guesses = []
    for k in range(48):
        g = synth.perturb_pose(p.gt_pose, rng, 5.0, 0.006)
        if k % 3 == 0:
            g = g.copy()
            g[:3, 3] += (0.03 + 0.05 * rng.random()) * np.array([1.0, -1.0, 0.5])
        guesses.append(g)
    guesses = np.stack(guesses)
    kw = dict(max_iterations=12, max_corr_dist=0.02, abs_mse_threshold=-1.0)
    assert _run(pcl, src, p.target, pcl.IterativeClosestPoint, None, guesses, 0, **kw) == \
           _run(pcl, src, p.target, pcl.IterativeClosestPoint, None, guesses, 1, **kw)
    for n in (1, 31, 129, 700):  # less than a warp, a block, a tile
        a = _run(pcl, p.source[:n], p.target, pcl.IterativeClosestPoint, None, guesses[:20], 0, **kw)
        b = _run(pcl, p.source[:n], p.target, pcl.IterativeClosestPoint, None, guesses[:20], 1, **kw)
        assert a == b
