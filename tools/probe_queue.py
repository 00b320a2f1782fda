"""Development probe: the batched align on C4 data, one launch per iteration (sub-stream chains) against
the work-queue kernel, for full (1024) and shard-sized (128, 256) batches.  Checks that every variant
returns bit-identical records and prints wall-clock throughput of the host-buffer call (uploads of the
poses + read-back included; the scene and the model stay resident)."""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from pose_estimation_b200 import pcl  # noqa: E402
from pose_estimation_b200.testing import synth  # noqa: E402


def main():
    scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
    sizes = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1024, 128]
    ctx = pcl.Context(0)

    def ds(points, leaf):
        vg = pcl.VoxelGrid(ctx)
        vg.setInputCloud(points)
        vg.setLeafSize(leaf)
        return vg.filter()

    prob = synth.make_c4(scale=scale, n_guesses=max(sizes), downsample=ds)
    print(f"C4 scale {scale}: target {len(prob.target)} pts, model {len(prob.source)} pts, {len(prob.guess)} poses", flush=True)
    icp = pcl.IterativeClosestPoint(ctx)
    icp.setMaximumIterations(30)
    icp.getConvergeCriteria().setAbsoluteMSE(-1.0)
    icp.setInputTarget(prob.target)
    icp.setInputSource(prob.source)
    icp.setMaxCorrespondenceDistance(0.02)

    def run(H, reps=3):
        best = 1e9
        res = None
        for _ in range(reps):
            t0 = time.perf_counter()
            res = icp.alignBatch(prob.guess[:H])
            best = min(best, time.perf_counter() - t0)
        return best, b"".join(bytes(r) for r in res)

    for H in sizes:
        ctx.set_int("work_queue", 0)
        t_ref, ref = run(H)
        print(f"H={H}: per-iteration launches {1e3 * t_ref:8.2f} ms = {H / t_ref:8.0f} hyp/s", flush=True)
        ctx.set_int("work_queue", 1)
        for qpt in (4, 8, 16, 32, 64):
            ctx.set_int("queue_queries_per_thread", qpt)
            t, out = run(H)
            a = np.frombuffer(ref, np.float32).reshape(H, 24)[:, :16]
            b = np.frombuffer(out, np.float32).reshape(H, 24)[:, :16]
            print(f"H={H}: queue, {qpt:3d} patches per item      {1e3 * t:8.2f} ms = {H / t:8.0f} hyp/s  "
                  f"{'bit-identical' if out == ref else f'max |dT| {np.abs(a - b).max():.2e}'}", flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
