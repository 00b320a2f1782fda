"""Profiling driver (not part of the bench contract): one pass of the C5 chain on configs[1]/[4]
data — VoxelGrid of the 2.33M-pt scene, k = 30 normals, target grid build, one 30-iteration
point-to-point align and one point-to-plane align — for `ncu` captures of the non-batched kernels."""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from pose_estimation_b200 import pcl  # noqa: E402
from pose_estimation_b200.testing import synth  # noqa: E402


def main():
    ctx = pcl.Context(0)

    def ds(points, leaf):
        vg = pcl.VoxelGrid(ctx)
        vg.setInputCloud(points)
        vg.setLeafSize(leaf)
        return vg.filter()

    prob = synth.make_c2(downsample=ds)
    for rep in range(2):
        t0 = time.perf_counter()
        tgt = ds(prob.organized, prob.leaf)
        ne = pcl.NormalEstimation(ctx)
        ne.setInputCloud(tgt)
        ne.setKSearch(30)
        nrm = ne.compute()
        icp = pcl.IterativeClosestPoint(ctx)
        icp.setInputSource(prob.source)
        icp.setInputTarget(tgt)
        icp.setMaximumIterations(30)
        icp.getConvergeCriteria().setAbsoluteMSE(-1.0)
        icp.align(prob.guess, want_output=False)
        f1 = icp.getFitnessScore()
        if rep == 1:  # the brute-force validator on the model's points (target = the down-sampled scene)
            ctx.nn_search(prob.source, bruteforce=True)
        icpn = pcl.IterativeClosestPointWithNormals(ctx)
        icpn.setInputSource(prob.source)
        icpn.setInputTarget(tgt, nrm)
        icpn.setMaximumIterations(30)
        icpn.getConvergeCriteria().setAbsoluteMSE(-1.0)
        icpn.align(prob.guess, want_output=False)
        print(f"pass {rep}: {len(tgt)} pts, p2p fitness {f1:.4e}, p2plane fitness {icpn.getFitnessScore():.4e}, "
              f"{1e3 * (time.perf_counter() - t0):.1f} ms host wall, launches {ctx.launch_count}")


if __name__ == "__main__":
    main()
