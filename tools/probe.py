"""Development probe (not part of the product or the bench contract): times the single align and
a small batch at full C2 size for every nn_group width and a few grid occupancies, with the
library's own per-launch CUDA-event profile."""
import ctypes as C
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from pose_estimation_b200 import pcl  # noqa: E402
from pose_estimation_b200.pcl import lib  # noqa: E402
from pose_estimation_b200.testing import synth  # noqa: E402


def profile(ctx):
    buf = np.zeros(256, np.float32)
    n = C.c_size_t(0)
    ctx.check(lib.peb_profile_read(ctx.handle, buf.ctypes.data, 256, C.byref(n)))
    return buf[: n.value]


def main():
    scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
    ctx = pcl.Context(0)

    def ds(points, leaf):
        vg = pcl.VoxelGrid(ctx)
        vg.setInputCloud(points)
        vg.setLeafSize(leaf)
        return vg.filter()

    t0 = time.perf_counter()
    prob = synth.make_c2(scale=scale, downsample=ds)
    print(f"C2 scale {scale}: scene {prob.organized.shape[0]} px -> target {len(prob.target)} pts (leaf {prob.leaf:.5f}), "
          f"model {len(prob.source)} pts, generated in {time.perf_counter() - t0:.1f} s", flush=True)
    for _ in range(3):
        t0 = time.perf_counter()
        ds(prob.organized, prob.leaf)
        print(f"  voxel grid (host buffers, e2e): {1e3 * (time.perf_counter() - t0):.2f} ms")

    icp = pcl.IterativeClosestPoint(ctx)
    icp.setMaximumIterations(30)
    icp.getConvergeCriteria().setAbsoluteMSE(-1.0)
    ctx.set_int("profile", 2)
    icp.setInputTarget(prob.target)
    rng = np.random.default_rng(0)
    guesses = np.stack([synth.perturb_pose(prob.gt_pose, rng, 6.0, 0.008) for _ in range(128)])
    icp.setInputSource(prob.source)
    icp.setMaxCorrespondenceDistance(0.02)
    for rows, guard in ((25, 40), (25, 60), (25, 100), (81, 60), (225, 60), (1000000, 60), (1000000, 100), (25, 0)):
        ctx.set_int("ball_direct_rows", rows)
        ctx.set_int("seed_guard_x10", guard)
        best = 1e9
        for rep in range(3):
            t0 = time.perf_counter()
            res = icp.alignBatch(guesses)
            best = min(best, time.perf_counter() - t0)
        pr = profile(ctx)
        fit = np.array([r.fitness for r in res])
        print(f"rows {rows:7d} guard {guard / 10}: batch H=128 iters(ms) {np.round(pr[:8], 2).tolist()} wall {1e3 * best:.2f} ms = {128 / best:.0f} hyp/s | "
              f"median {np.median(pr[:-1]):.3f} ms | fitness median {np.median(fit):.3e}", flush=True)
    # normals
    ne = pcl.NormalEstimation(ctx)
    ne.setInputCloud(prob.target)
    ne.setKSearch(30)
    for _ in range(3):
        t0 = time.perf_counter()
        ne.compute()
        print(f"normals k=30 on {len(prob.target)} pts (host buffers, e2e): {1e3 * (time.perf_counter() - t0):.2f} ms")
    print("launches:", ctx.launch_count)


if __name__ == "__main__":
    main()
