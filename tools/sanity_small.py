"""Small run through every kernel of the library (for compute-sanitizer): prefilter, VoxelGrid, normals,
grid NN + brute force, single p2p / p2plane aligns with all outputs, a batch with anchors and sub-streams."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from pose_estimation_b200 import pcl  # noqa: E402
from pose_estimation_b200.testing import synth  # noqa: E402

ctx = pcl.Context(0)
rng = np.random.default_rng(1)
surf = synth.Surface(1)
gt = synth.default_gt_pose(rng)
scene = synth.render_scene(surf, gt, rng, 243, 150)
pf = pcl.ScenePrefilter(ctx)
pf.setInputCloud(scene)
pf.setSphereFilter((0, 0, 0.7), 0.5)
pf.addPlane(1.0, 0.0, 0.0, 5.0)
kept = pf.filter()
vg = pcl.VoxelGrid(ctx)
vg.setInputCloud(kept)
vg.setLeafSize(0.004)
vg.setMinimumPointsNumberPerVoxel(2)
tgt = vg.filter()
ne = pcl.NormalEstimation(ctx)
ne.setInputCloud(tgt)
ne.setKSearch(12)
nrm, nn = ne.compute(return_neighbours=True)
model, _ = surf.sample(3000, rng)
model = synth.xyz4(model.astype(np.float32))
model[7, 0] = np.nan
guess = synth.perturb_pose(gt, rng, 2.0, 0.003, exact=True)
ctx.target_set(tgt)
q = synth.apply_pose(guess, model[:, :3].astype(np.float64)).astype(np.float32)
gi, gd = ctx.nn_search(q)
bi, bd = ctx.nn_search(q, bruteforce=True)
ok = np.isfinite(q).all(1)
assert np.array_equal(gi[ok], bi[ok])
for cls, normals in ((pcl.IterativeClosestPoint, None), (pcl.IterativeClosestPointWithNormals, nrm)):
    icp = cls(ctx)
    icp.setInputSource(model)
    icp.setInputTarget(tgt, normals)
    icp.setMaximumIterations(12)
    icp.setMaxCorrespondenceDistance(0.02)
    icp.align(guess, want_correspondences=True)
    print(cls.__name__, icp.nr_iterations_, icp.result.state, f"{icp.getFitnessScore():.3e}", len(icp.trace()))
    print("  fitness(range)", icp.getFitnessScore(1e-5))
guesses = np.stack([synth.perturb_pose(gt, rng, 5.0, 0.006) for _ in range(400)])
icp = pcl.IterativeClosestPoint(ctx)
icp.setInputSource(model)
icp.setInputTarget(tgt)
icp.setMaximumIterations(8)
icp.setMaxCorrespondenceDistance(0.02)
res = icp.alignBatch(guesses)
print("batch", len(res), sum(r.iterations for r in res), f"{np.median([r.fitness for r in res]):.3e}")
ctx.set_int("cert_margin_x1000", 300)
res2 = icp.alignBatch(guesses[:40])
ctx.set_int("cert_margin_x1000", 0)
ctx.set_int("nn_group", 8)
res3 = icp.alignBatch(guesses[:40])
print("variants ok", all(bytes(a.T) == bytes(b.T) for a, b in zip(res2, res3)))
ctx.close()
print("sanity ok, launches")
