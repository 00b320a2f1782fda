"""The `warm_upfront` option of the library (pose_estimation_b200/csrc/nn_upfront.cuh): the warm searches of iterations
>= 1 fetch all row bounds of their ball before the first scan.  Same bit-identity checks as
test_gpu_parity.py::test_speed_options_never_change_results.

Status (round 2, B200, profiles/README.md): results bit-identical, but SLOWER than the narrowing walk on C4 (2 x 2 rows:
11 050 against 11 490 hypotheses/s; 3 x 3: 9 260), so it stays off; the option is kept as a recorded experiment and these
tests keep it honest.
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


def _run(pcl, p, cls, normals, guesses, upfront, **params):
    """upfront: 0 = the default kernels, 1 = balls up to 2 x 2 grid rows take the up-front path, 3 = up to 3 x 3"""
    from oracle import default_params

    c = pcl.Context(0)
    if upfront:
        c.set_int("warm_upfront", upfront)
        c.set_int("warm_upfront_from", 1)  # every warm launch, also the first one with its large balls
    icp = cls(c)
    icp.setInputSource(p.source)
    icp.setInputTarget(p.target, normals)
    prm = default_params(**params)
    for name, _ in prm._fields_:
        if name != "estimator":
            setattr(icp.params, name, getattr(prm, name))
    res = icp.alignBatch(guesses)
    icp.align(p.guess, want_correspondences=True)
    out = ([bytes(r) for r in res], bytes(icp.result), icp.correspondences[0].copy(), icp.correspondences[1].copy())
    c.close()
    return out


@pytest.mark.parametrize("rows", [1, 3])
def test_upfront_warm_search_never_changes_results(pcl, scene_small, rows):
    p = scene_small
    rng = np.random.default_rng(5)
    guesses = np.stack([synth.perturb_pose(p.gt_pose, rng, 5.0, 0.006) for _ in range(200)])
    a = _run(pcl, p, pcl.IterativeClosestPoint, None, guesses, 0, max_iterations=25, max_corr_dist=0.02, abs_mse_threshold=-1.0)
    b = _run(pcl, p, pcl.IterativeClosestPoint, None, guesses, rows, max_iterations=25, max_corr_dist=0.02, abs_mse_threshold=-1.0)
    assert a[0] == b[0] and a[1] == b[1]
    assert np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3])


@pytest.mark.parametrize("rows", [1, 3])
def test_upfront_warm_search_point_to_plane_and_criteria(pcl, scene_small, rows):
    p = scene_small
    c = pcl.Context(0)
    ne = pcl.NormalEstimation(c)
    ne.setInputCloud(p.target)
    ne.setKSearch(12)
    normals = ne.compute()
    c.close()
    rng = np.random.default_rng(6)
    guesses = np.stack([synth.perturb_pose(p.gt_pose, rng, 4.0, 0.004) for _ in range(40)])
    kw = dict(max_iterations=30, max_corr_dist=0.01, transformation_epsilon=1e-9)  # hypotheses stop at different iterations
    a = _run(pcl, p, pcl.IterativeClosestPointWithNormals, normals, guesses, 0, **kw)
    b = _run(pcl, p, pcl.IterativeClosestPointWithNormals, normals, guesses, rows, **kw)
    assert a[0] == b[0] and a[1] == b[1]
    assert np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3])
