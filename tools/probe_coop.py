"""Development probe: first-iteration verification per lane against the warp-cooperative one
(coop_max_rows), C4 data; checks bit-identical records and prints wall-clock of the host-buffer call
and the device time of launch 0 (profile level 2)."""
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
    sizes = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1024, 128]
    ctx = pcl.Context(0)

    def ds(points, leaf):
        vg = pcl.VoxelGrid(ctx)
        vg.setInputCloud(points)
        vg.setLeafSize(leaf)
        return vg.filter()

    prob = synth.make_c4(scale=scale, n_guesses=max(sizes), downsample=ds)
    print(f"C4 scale {scale}: target {len(prob.target)} pts, model {len(prob.source)} pts", flush=True)
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
        ref = None
        import os
        for rows in ((0,) if os.environ.get("PEB_LIB_VARIANT") else (0, 1024)):
            if not os.environ.get("PEB_LIB_VARIANT"):
                ctx.set_int("coop_max_rows", rows)
            t, out = run(H)
            ctx.set_int("profile", 2)
            run(H, 1)
            pr = profile(ctx)
            ctx.set_int("profile", 0)
            if ref is None:
                ref = out
            print(f"H={H}: coop_max_rows {rows:5d}: {1e3 * t:8.2f} ms = {H / t:8.0f} hyp/s, launch 0 {pr[0]:7.3f} ms  "
                  f"{'bit-identical' if out == ref else 'RESULTS DIFFER'}", flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
