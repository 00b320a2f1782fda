"""Randomised consistency run (development): batches of many sizes (1 .. 70 hypotheses, including the thresholds
where chains / per-hypothesis launch dependencies / anchors switch on), source sizes and iteration counts; every
record of a batch must equal the single align of the same pose bit for bit, with convergence criteria on and off."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from pose_estimation_b200 import pcl  # noqa: E402
from pose_estimation_b200.testing import synth  # noqa: E402

ctx = pcl.Context(0)


def _downsample(points, leaf):
    vg = pcl.VoxelGrid(ctx)
    vg.setInputCloud(points)
    vg.setLeafSize(leaf)
    return vg.filter()


prob = synth.make_c2(scale=0.3, downsample=_downsample)
rng = np.random.default_rng(0)
bad = 0
cases = 0
for trial in range(40):
    H = int(rng.choice([1, 2, 3, 7, 15, 16, 17, 31, 32, 33, 47, 64, 70, 300]))
    n = int(rng.choice([40, 500, len(prob.source)]))
    its = int(rng.choice([1, 2, 5, 30]))
    crit = bool(rng.integers(2))
    src = prob.source[rng.choice(len(prob.source), n, replace=False)]
    guesses = np.stack([synth.perturb_pose(prob.gt_pose, rng, float(rng.uniform(0.1, 6)), float(rng.uniform(0, 0.008))) for _ in range(H)])
    icp = pcl.IterativeClosestPoint(ctx)
    icp.setInputSource(src)
    icp.setInputTarget(prob.target)
    icp.setMaximumIterations(its)
    icp.setMaxCorrespondenceDistance(0.02)
    icp.getConvergeCriteria().setAbsoluteMSE(-1.0 if not crit else 1e-12)
    icp.setTransformationEpsilon(1e-9 if crit else 0.0)
    res = icp.alignBatch(guesses)
    for h in rng.choice(H, min(H, 6), replace=False):
        icp.align(guesses[h], want_output=False)
        s = icp.result
        cases += 1
        if not (bytes(s.T) == bytes(res[h].T) and s.fitness == res[h].fitness and s.iterations == res[h].iterations and
                s.state == res[h].state):
            bad += 1
            print(f"MISMATCH trial {trial} H {H} n {n} its {its} crit {crit} h {h}: {s.iterations}/{res[h].iterations} "
                  f"{s.state}/{res[h].state} {s.fitness}/{res[h].fitness}")
print(f"stress: {cases} comparisons, {bad} mismatches")
ctx.close()
sys.exit(1 if bad else 0)
