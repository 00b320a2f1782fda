"""Variants of the warm launches of a batch (pose_estimation_b200/csrc/icp.cu), each against the plain warm kernel:

* `warm_graph` (default ON for batches of >= 32 hypotheses): the search over the target's k-nearest-neighbour graph
  (csrc/nn_graph.cuh) — certificate from the previous match's row, greedy steps, grid walk as the last resort;
* `warm_graph_queue` (off: measured slower): in a graph launch every warp queues the unproven queries of a tile of 8 / 16 passes
  and walks the grid for them 32 at a time (icp_iteration_graphq_kernel);
* `warm_bin` (off: measured slower): the queries of a block sorted by the number of grid rows their ball search walks
  (icp_iteration_binned_kernel).

Queries, exact searches with the index tie rule and the order of every thread's double sums are the same in all of them,
so every record must be byte-identical — point-to-point and point-to-plane, hypotheses that stop at different
iterations, clouds with non-finite points, far hypotheses that search cold, sources smaller than a tile, rejector on.
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


VARIANTS = {"plain": (("warm_graph", 0), ("warm_bin", 0)),
            # every warm launch over the graph / hypotheses switch to it when their MSE says the certificate will hold
            "graph": (("warm_graph", 1), ("warm_graph_min_hyp", 2), ("warm_graph_kappa_x100", 0), ("warm_bin", 0),
                      ("warm_graph_queue", 0)),
            "graph_auto": (("warm_graph", 1), ("warm_graph_min_hyp", 2), ("warm_bin", 0), ("warm_graph_queue", 0)),
            # ... with the unproven queries of a tile of 8 / 16 passes queued per warp and walked 32 at a time (off: measured slower)
            "graphq": (("warm_graph", 1), ("warm_graph_min_hyp", 2), ("warm_graph_kappa_x100", 0), ("warm_bin", 0),
                       ("warm_graph_queue", 8)),
            "graphq_auto": (("warm_graph", 1), ("warm_graph_min_hyp", 2), ("warm_bin", 0), ("warm_graph_queue", 8)),
            "graphq16": (("warm_graph", 1), ("warm_graph_min_hyp", 2), ("warm_graph_kappa_x100", 0), ("warm_bin", 0),
                        ("warm_graph_queue", 16)),
            # launch 0 takes its candidates from a greedy descent on the graph instead of the 3 x 3 x 3 probe
            "cold_graph": (("warm_graph", 1), ("warm_graph_min_hyp", 2), ("warm_bin", 0), ("warm_graph_queue", 0),
                           ("cold_graph", 1)),
            # every graph launch peeks at the four nearest neighbours of the previous match before a walk / none does
            "peek_all": (("warm_graph", 1), ("warm_graph_min_hyp", 2), ("warm_bin", 0), ("warm_graph_queue", 0),
                         ("warm_graph_peek", 1000)),
            "peek_none": (("warm_graph", 1), ("warm_graph_min_hyp", 2), ("warm_bin", 0), ("warm_graph_queue", 0),
                          ("warm_graph_peek", 0)),
            # without the flatness certificate (only the triangle inequality proves a match)
            "flat_off": (("warm_graph", 1), ("warm_graph_min_hyp", 2), ("warm_bin", 0), ("warm_graph_queue", 0),
                         ("warm_graph_flat", 0)),
            "flat_all": (("warm_graph", 1), ("warm_graph_min_hyp", 2), ("warm_bin", 0), ("warm_graph_queue", 0),
                         ("warm_graph_flat", 1), ("warm_graph_flat_from", 1), ("warm_graph_flat_until", 1000)),
            "bin": (("warm_graph", 0), ("warm_bin", 1))}


def _run(pcl, source, target, cls, normals, guesses, variant, opts=(), **params):
    from oracle import default_params

    c = pcl.Context(0)
    for k, v in VARIANTS[variant]:
        c.set_int(k, v)
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


@pytest.mark.parametrize("variant", ["graph", "graph_auto", "graphq", "graphq_auto", "graphq16", "cold_graph", "peek_all", "peek_none", "flat_off", "flat_all", "bin"])
def test_warm_variants_never_change_results(pcl, scene_small, variant):
    p = scene_small
    rng = np.random.default_rng(15)
    guesses = np.stack([synth.perturb_pose(p.gt_pose, rng, 5.0, 0.006) for _ in range(200)])
    kw = dict(max_iterations=25, max_corr_dist=0.02, abs_mse_threshold=-1.0)
    a = _run(pcl, p.source, p.target, pcl.IterativeClosestPoint, None, guesses, "plain", **kw)
    b = _run(pcl, p.source, p.target, pcl.IterativeClosestPoint, None, guesses, variant, **kw)
    assert a == b
    # other launch shapes: one chain without per-hypothesis dependencies, many small blocks
    for opts in ((("batch_streams", 1), ("flag_deps", 0)), (("blocks_factor", 96),), (("blocks_factor", 4),)):
        assert _run(pcl, p.source, p.target, pcl.IterativeClosestPoint, None, guesses, variant, opts, **kw) == \
               _run(pcl, p.source, p.target, pcl.IterativeClosestPoint, None, guesses, "plain", opts, **kw)


@pytest.mark.parametrize("variant", ["graph", "graph_auto", "graphq", "graphq_auto", "cold_graph", "peek_all", "peek_none", "flat_off", "flat_all", "bin"])
def test_warm_variants_point_to_plane_rejector_and_criteria(pcl, scene_small, variant):
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
    a = _run(pcl, p.source, p.target, pcl.IterativeClosestPointWithNormals, normals, guesses, "plain", **kw)
    b = _run(pcl, p.source, p.target, pcl.IterativeClosestPointWithNormals, normals, guesses, variant, **kw)
    assert a == b
    assert len({r[-32:] for r in a}) > 1  # (the records differ between hypotheses: the comparison is not vacuous)


@pytest.mark.parametrize("variant", ["graph", "graph_auto", "graphq", "graphq_auto", "cold_graph", "peek_all", "peek_none", "flat_off", "flat_all", "bin"])
def test_warm_variants_nonfinite_points_far_hypotheses_and_tiny_sources(pcl, scene_small, variant):
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
    assert _run(pcl, src, p.target, pcl.IterativeClosestPoint, None, guesses, "plain", **kw) == \
           _run(pcl, src, p.target, pcl.IterativeClosestPoint, None, guesses, variant, **kw)
    # targets smaller than a row of the graph, and a target that is one point repeated
    for tgt in (p.target[:1], p.target[:7], p.target[:16], np.repeat(p.target[:1], 40, 0)):
        assert _run(pcl, src[:400], tgt, pcl.IterativeClosestPoint, None, guesses, "plain", **kw) == \
               _run(pcl, src[:400], tgt, pcl.IterativeClosestPoint, None, guesses, variant, **kw)
    for n in (1, 31, 129, 700):  # less than a warp, a block, a tile
        a = _run(pcl, p.source[:n], p.target, pcl.IterativeClosestPoint, None, guesses[:20], "plain", **kw)
        b = _run(pcl, p.source[:n], p.target, pcl.IterativeClosestPoint, None, guesses[:20], variant, **kw)
        assert a == b
