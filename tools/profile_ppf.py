"""Profiling driver (not part of the bench contract): PPF3DDetector train + match on the C2 clouds, for ncu launch lists."""
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

    def with_normals(cloud, viewpoint):
        ne = pcl.NormalEstimation(ctx)
        ne.setInputCloud(cloud)
        ne.setKSearch(20)
        ne.setViewPoint(*viewpoint)
        nrm = ne.compute()
        ok = np.isfinite(nrm[:, :3]).all(1) & np.isfinite(cloud[:, :3]).all(1)
        return np.ascontiguousarray(np.concatenate([cloud[ok, :3], nrm[ok, :3]], 1), np.float32)

    prob = synth.make_c2(downsample=ds)
    scene6 = with_normals(prob.target[prob.target[:, 2] < 0.735], (0.0, 0.0, 0.0))
    model6 = with_normals(prob.source, (0.0, 0.0, 1.0))
    det = pcl.PPF3DDetector(0.03, 0.03, 40, ctx=ctx)
    for rep in range(2):
        t0 = time.perf_counter()
        det.trainModel(model6)
        t1 = time.perf_counter()
        res = det.match(scene6, 1.0, 0.03)
        t2 = time.perf_counter()
        _, raw = det.match(scene6, 1.0, 0.03, return_raw=True)
        print(f"pass {rep}: train {1e3 * (t1 - t0):.2f} ms, match {1e3 * (t2 - t1):.2f} ms, {len(raw)} reference points, "
              f"{len(res)} clusters, top votes {res[0].num_votes}, error {synth.pose_error(res[0].matrix, prob.gt_pose)}")


if __name__ == "__main__":
    main()
