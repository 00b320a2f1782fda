"""cv::ppf_match_3d::ICP::registerModelToScene — what the reference runs in the refinement slot
(pose_estimation/src/opencv_surface_match.cpp:85-94).  The oracle restates the algorithm from recollection (PARITY
UNPINNED: opencv_contrib is in neither the reference tree nor this image); the CPU part checks that the restatement does
what an ICP must (it converges to the ground truth, the pyramid and the rejection behave), the GPU part checks the
library against it through the C ABI."""
import numpy as np
import pytest

from oracle import cvicp_params
from pose_estimation_b200.testing import synth


make_case = synth.make_cvicp_case


def test_oracle_converges_to_the_ground_truth(oracle):
    model, scene, poses, gt = make_case()
    P, res = oracle.cvicp_register(model, scene, poses, cvicp_params())
    for before, after, r in zip(poses, P, res):
        rot0, tr0 = synth.pose_error(before, gt)
        rot, tr = synth.pose_error(after, gt)
        assert rot < 5e-4 and tr < 5e-5 and rot < 0.1 * rot0 and 0 < r < 1e-3
    # the same optimum from every start
    assert max(synth.pose_error(P[0], Q)[0] for Q in P[1:]) < 1e-5


def test_oracle_rejection_and_pyramid(oracle):
    model, scene, poses, gt = make_case(seed=1, clutter=3000)
    P, _ = oracle.cvicp_register(model, scene, poses, cvicp_params())
    assert max(synth.pose_error(Q, gt)[0] for Q in P) < 1e-3  # robust rejection copes with 25 % clutter
    # one level, one iteration: a single linearised step from the start pose, still an improvement
    P1, _ = oracle.cvicp_register(model, scene, poses[:2], cvicp_params(iterations=1, num_levels=1))
    for before, after in zip(poses[:2], P1):
        assert synth.pose_error(after, gt)[1] < synth.pose_error(before, gt)[1]
    # an identity start on an already aligned pair stays put
    aligned = np.stack([gt])
    P2, _ = oracle.cvicp_register(model, scene, aligned, cvicp_params())
    assert synth.pose_error(P2[0], gt)[0] < 5e-4


def test_oracle_degenerate_inputs(oracle):
    model, scene, poses, gt = make_case(seed=4, n_model=40, n_scene=90, n_poses=2)
    # 8 levels on 40 points: the coarse levels hold fewer than 6 pairs (or no point at all) and must leave the pose alone
    P, res = oracle.cvicp_register(model, scene, poses, cvicp_params())
    assert np.isfinite(P).all() and np.isfinite(res).all()
    # fewer than 6 points: no level can solve, the poses come back unchanged and the residual is the initial sentinel or 0
    P5, _ = oracle.cvicp_register(model[:5], scene, poses, cvicp_params())
    assert np.allclose(P5, poses, atol=1e-12)
    # no iterations allowed: unchanged as well
    P0, _ = oracle.cvicp_register(model, scene, poses, cvicp_params(iterations=0))
    assert np.allclose(P0, poses, atol=1e-12)


@pytest.fixture(scope="module")
def pcl():
    from pose_estimation_b200 import pcl as m

    return m


@pytest.fixture(scope="module")
def ctx(pcl):
    c = pcl.Context(0)
    yield c
    c.close()


@pytest.mark.gpu
@pytest.mark.parametrize("seed,clutter,kw", [(0, 0, {}), (1, 3000, {}), (2, 0, dict(rejection_scale=0.0)),
                                             (3, 500, dict(iterations=40, num_levels=3, tolerance=0.02))])
def test_library_matches_the_oracle(pcl, ctx, oracle, seed, clutter, kw):
    model, scene, poses, gt = make_case(seed=seed, clutter=clutter)
    prm = cvicp_params(**kw)
    ref, ref_res = oracle.cvicp_register(model, scene, poses, prm)
    icp = pcl.CvIcp(prm.iterations, prm.tolerance, prm.rejection_scale, prm.num_levels, ctx=ctx)
    got, res = icp.registerModelToScene(model, scene, poses)
    for a, b, ra, rb in zip(got, ref, res, ref_res):
        rot, tr = synth.pose_error(a, b)
        # the nearest neighbour is searched in the scene's own coordinates (the oracle: in the normalised frame) and the
        # inlier sums run in another order: same iterates up to rounding, a flipped near-tie moves the pose by ~1e-6
        assert rot < 2e-5 and tr < 2e-6, (rot, tr)
        assert abs(ra - rb) <= 1e-3 * rb
    assert max(synth.pose_error(Q, gt)[0] for Q in got) < 1e-3


@pytest.mark.gpu
def test_reference_sized_call(pcl, ctx, oracle):
    """The shape of the reference's call: ICP(250, 0.005f, 2.5f, 8) on a 50k-point model against a ~200k-point scene with
    6 poses; the oracle runs the same (a few seconds)."""
    model, scene, poses, gt = make_case(seed=5, n_model=50000, n_scene=200000, clutter=20000)
    icp = pcl.CvIcp(250, 0.005, 2.5, 8, ctx=ctx)
    got, res = icp.registerModelToScene(model, scene, poses)
    ref, ref_res = oracle.cvicp_register(model, scene, poses, cvicp_params())
    for a, b in zip(got, ref):
        rot, tr = synth.pose_error(a, b)
        assert rot < 5e-5 and tr < 5e-6, (rot, tr)
    assert max(synth.pose_error(Q, gt)[0] for Q in got) < 1e-3


@pytest.mark.gpu
def test_argument_errors(pcl, ctx):
    icp = pcl.CvIcp(ctx=ctx)
    with pytest.raises(pcl.PebError):
        icp.registerModelToScene(np.zeros((10, 3), np.float32), np.zeros((10, 6), np.float32), np.eye(4)[None])
    P, r = icp.registerModelToScene(np.zeros((10, 6), np.float32), np.zeros((10, 6), np.float32), np.zeros((0, 4, 4)))
    assert P.shape == (0, 4, 4)
