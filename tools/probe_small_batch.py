"""Development probe: where a shard-sized batch (128 hypotheses, the 8-GPU share of C4) loses efficiency
against the full batch: per-iteration device times (profile level 2, single chain) and wall-clock of the
host-buffer call for several chain counts / blocks-per-SM factors."""
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
    ctx = pcl.Context(0)

    def ds(points, leaf):
        vg = pcl.VoxelGrid(ctx)
        vg.setInputCloud(points)
        vg.setLeafSize(leaf)
        return vg.filter()

    prob = synth.make_c4(scale=1.0, n_guesses=1024, downsample=ds)
    icp = pcl.IterativeClosestPoint(ctx)
    icp.setMaximumIterations(30)
    icp.getConvergeCriteria().setAbsoluteMSE(-1.0)
    icp.setInputTarget(prob.target)
    icp.setInputSource(prob.source)
    icp.setMaxCorrespondenceDistance(0.02)

    def wall(H, reps=4):
        best = 1e9
        for _ in range(reps):
            t0 = time.perf_counter()
            icp.alignBatch(prob.guess[:H])
            best = min(best, time.perf_counter() - t0)
        return 1e3 * best

    for H in (128, 1024):
        ctx.set_int("profile", 2)
        wall(H, 2)
        pr = profile(ctx)
        ctx.set_int("profile", 0)
        print(f"H={H} single chain per-launch ms: first {np.round(pr[:4], 3).tolist()} last {np.round(pr[-4:], 3).tolist()} "
              f"sum {pr.sum():.2f} (iterations {pr[:-1].sum():.2f}, fitness {pr[-1]:.3f})", flush=True)
        for deps in (0, 1):
            for streams in (1, 2, 4):
                ctx.set_int("flag_deps", deps)
                ctx.set_int("batch_streams", streams)
                print(f"H={H} flag_deps {deps} chains {streams}: wall {wall(H):7.2f} ms", flush=True)
        ctx.set_int("flag_deps", 1)
        ctx.set_int("batch_streams", 0)
        ctx.set_int("blocks_factor", 0)
        print(f"H={H} defaults: wall {wall(H):7.2f} ms", flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
