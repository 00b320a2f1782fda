"""BASELINE.json's configurations at FULL size on the B200, against the oracle where it finishes in
seconds (it does for single aligns) and through size-independent properties otherwise."""
import numpy as np
import pytest

from oracle import DBL_MAX, default_params
from pose_estimation_b200.testing import synth
from util import pose_delta, tie_ok

pytestmark = pytest.mark.gpu


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
def c2(pcl, ctx):
    """configs[1]: 1944 x 1200 organized scene -> VoxelGrid (CUDA path) -> ~200k target, 50k model."""

    def ds(points, leaf):
        vg = pcl.VoxelGrid(ctx)
        vg.setInputCloud(points)
        vg.setLeafSize(leaf)
        return vg.filter()

    return synth.make_c2(downsample=ds)


def _params(icp, prm):
    for name, _ in prm._fields_:
        setattr(icp.params, name, getattr(prm, name))


def test_c5_voxel_grid_2m3_points_bit_exact(pcl, ctx, oracle, c2):
    assert c2.organized.shape[0] == 1944 * 1200
    assert 195000 <= len(c2.target) <= 205000
    ref, unchanged = oracle.voxel_grid(c2.organized, c2.leaf)
    assert not unchanged
    assert np.array_equal(c2.target.view(np.uint32), ref.view(np.uint32))
    # properties that do not need the oracle: ascending, unique voxel ids (PCL's output order)
    inv = np.float32(1.0) / np.float32(c2.leaf)
    fin = c2.organized[np.isfinite(c2.organized[:, :3]).all(1), :3]
    mn = np.floor(fin.min(0) * inv).astype(np.int64)
    dims = np.floor(fin.max(0) * inv).astype(np.int64) - mn + 1
    ijk = np.floor(c2.target[:, :3] * inv).astype(np.int64) - mn
    key = ijk[:, 0] + dims[0] * (ijk[:, 1] + dims[1] * ijk[:, 2])
    assert (np.diff(key) > 0).all()


def test_c2_grid_search_equals_validator_at_full_size(pcl, ctx, c2):
    ctx.target_set(c2.target)
    q = synth.apply_pose(c2.guess, c2.source[:, :3].astype(np.float64)).astype(np.float32)
    gi, gd = ctx.nn_search(q)
    bi, bd = ctx.nn_search(q, bruteforce=True)          # 50k x 200k FP32 brute force on the device
    assert np.array_equal(gi, bi) and np.array_equal(gd, bd)


def test_c2_single_align_full_size_vs_oracle(pcl, ctx, oracle, c2):
    prm = default_params(max_iterations=30, abs_mse_threshold=-1.0)
    icp = pcl.IterativeClosestPoint(ctx)
    icp.setInputSource(c2.source)
    icp.setInputTarget(c2.target)
    _params(icp, prm)
    icp.align(c2.guess, want_correspondences=True, want_output=False)
    ref = oracle.icp(c2.target, wide_accum=True).align(c2.source, c2.guess, prm, trace_cap=30)
    r = ref["result"]
    assert icp.result.iterations == r.iterations == 30 and icp.result.state == r.state
    rot, tr = pose_delta(icp.getFinalTransformation(), r.matrix())
    assert rot < 1e-5 and tr < 1e-5, (rot, tr)
    lockstep = np.array_equal(icp.trace().view(np.uint32), ref["trace_T"].view(np.uint32))
    assert abs(icp.result.fitness - r.fitness) <= (1e-6 if lockstep else 1e-3) * r.fitness
    f_on_ref, _ = ctx.fitness_score(r.matrix())
    assert abs(f_on_ref - r.fitness) <= 1e-6 * r.fitness
    same = icp.correspondences[0] == ref["corr_idx"]
    assert same.mean() > 0.999
    rot, tr = pose_delta(icp.getFinalTransformation(), c2.gt_pose)
    assert rot < 1e-2 and tr < 2e-3   # 30 point-to-point iterations from the 2 deg / 3 mm guess (slides slowly)


def test_c3_point_to_plane_with_gpu_normals_full_size_vs_oracle(pcl, ctx, oracle, c2):
    ne = pcl.NormalEstimation(ctx)
    ne.setInputCloud(c2.target)
    ne.setKSearch(30)
    normals, nn = ne.compute(return_neighbours=True)
    ref_n, ref_nn = oracle.normals(c2.target, 30, threads=0, want_nn=True)
    assert tie_ok(c2.target, c2.target, nn.reshape(-1), ref_nn.reshape(-1))
    same = (nn == ref_nn).all(1)
    assert same.mean() > 0.999
    ang = np.degrees(np.arccos(np.clip((normals[same, :3] * ref_n[same, :3]).sum(1), -1, 1)))
    assert np.nanmax(ang) < 0.1 and np.nanmedian(ang) < 1e-3
    # the align uses the GPU's own normals on both sides, so that it tests the estimator alone
    prm = default_params(max_iterations=30, abs_mse_threshold=-1.0, estimator=1)
    icp = pcl.IterativeClosestPointWithNormals(ctx)
    icp.setInputSource(c2.source)
    icp.setInputTarget(c2.target, normals)
    _params(icp, prm)
    icp.align(c2.guess, want_output=False)
    r = oracle.icp(c2.target, normals).align(c2.source, c2.guess, prm)["result"]
    assert icp.result.iterations == r.iterations == 30
    rot, tr = pose_delta(icp.getFinalTransformation(), r.matrix())
    assert rot < 1e-5 and tr < 1e-5, (rot, tr)
    assert abs(icp.result.fitness - r.fitness) <= 1e-3 * r.fitness
    rot, tr = pose_delta(icp.getFinalTransformation(), c2.gt_pose)
    assert rot < 2e-3 and tr < 5e-4   # point-to-plane gets closer to the truth than point-to-point


def test_c5_end_to_end_50_iterations(pcl, ctx, oracle, c2):
    """2.3M-pt scene -> VoxelGrid -> normals(k=30) -> 50-iteration point-to-plane, all on the device path;
    the oracle runs the same chain on the CPU."""
    vg = pcl.VoxelGrid(ctx)
    vg.setInputCloud(c2.organized)
    vg.setLeafSize(c2.leaf)
    tgt = vg.filter()
    ne = pcl.NormalEstimation(ctx)
    ne.setInputCloud(tgt)
    ne.setKSearch(30)
    nrm = ne.compute()
    icp = pcl.IterativeClosestPointWithNormals(ctx)
    icp.setInputSource(c2.source)
    icp.setInputTarget(tgt, nrm)
    icp.setMaximumIterations(50)
    icp.getConvergeCriteria().setAbsoluteMSE(-1.0)
    icp.align(c2.guess, want_output=False)
    o_tgt, _ = oracle.voxel_grid(c2.organized, c2.leaf)
    o_nrm = oracle.normals(o_tgt, 30, threads=0)
    prm = default_params(max_iterations=50, abs_mse_threshold=-1.0, estimator=1)
    r = oracle.icp(o_tgt, o_nrm).align(c2.source, c2.guess, prm)["result"]
    assert icp.result.iterations == r.iterations == 50
    rot, tr = pose_delta(icp.getFinalTransformation(), r.matrix())
    # the two chains differ by the libm-vs-CUDA ulps of the normals (<= 0.1 deg on a few points)
    assert rot < 2e-5 and tr < 2e-5, (rot, tr)
    assert abs(icp.result.fitness - r.fitness) <= 1e-3 * r.fitness


def test_c4_batch_sample_full_size_vs_oracle(pcl, ctx, oracle):
    """configs[3] at full size (500k scene): a sample of the 1024 hypotheses against the oracle, and the
    whole batch through properties (all run 30 iterations, all converge to the same pose)."""

    def ds(points, leaf):
        vg = pcl.VoxelGrid(ctx)
        vg.setInputCloud(points)
        vg.setLeafSize(leaf)
        return vg.filter()

    c4 = synth.make_c4(n_guesses=1024, downsample=ds)
    assert 490000 <= len(c4.target) <= 510000
    prm = default_params(max_iterations=30, abs_mse_threshold=-1.0, max_corr_dist=0.02)
    icp = pcl.IterativeClosestPoint(ctx)
    icp.setInputSource(c4.source)
    icp.setInputTarget(c4.target)
    _params(icp, prm)
    res = icp.alignBatch(c4.guess)
    sample = [int(h) for h in np.linspace(0, 1023, 96).round()]  # every 11th hypothesis: ~3 s of the CPU port on 16 cores
    ref = oracle.icp(c4.target, wide_accum=True).align_batch(c4.source, c4.guess[sample], prm, threads=0)
    worst = [0.0, 0.0, 0.0]
    for h, r in zip(sample, ref):
        g = res[h]
        assert g.iterations == r.iterations == 30 and g.state == r.state
        rot, tr = pose_delta(pcl.result_matrix(g), r.matrix())
        assert rot < 1e-5 and tr < 1e-5, (h, rot, tr)
        assert abs(g.fitness - r.fitness) <= 1e-3 * r.fitness
        worst = [max(worst[0], rot), max(worst[1], tr), max(worst[2], abs(g.fitness - r.fitness) / r.fitness)]
    print(f"C4 full size, {len(sample)} of 1024 hypotheses vs the oracle: worst rotation {worst[0]:.2e} rad, "
          f"translation {worst[1]:.2e} m, fitness {worst[2]:.2e} relative")
    its = np.array([r.iterations for r in res])
    fit = np.array([r.fitness for r in res])
    assert (its == 30).all() and np.isfinite(fit).all()
    # refinement property: 30 point-to-point iterations bring (almost) every hypothesis closer to the truth
    before = np.array([pose_delta(g, c4.gt_pose) for g in c4.guess])
    after = np.array([pose_delta(pcl.result_matrix(r), c4.gt_pose) for r in res])
    assert ((after[:, 0] < before[:, 0]) | (after[:, 1] < before[:, 1])).mean() > 0.98
    assert np.median(after[:, 1]) < 0.5 * np.median(before[:, 1])
    # batch == single, bit for bit, also at this size
    icp.align(c4.guess[511], want_output=False)
    assert bytes(icp.result.T) == bytes(res[511].T) and icp.result.fitness == res[511].fitness
