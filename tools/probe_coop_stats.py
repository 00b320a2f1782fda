"""Development probe (needs the PEB_COOP_STATS build, PEB_LIB_VARIANT=stats): how many grid rows and
staged points the warp-cooperative first-iteration verification touches per 32-point patch."""
import ctypes as C
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from pose_estimation_b200 import pcl  # noqa: E402
from pose_estimation_b200.pcl import lib  # noqa: E402
from pose_estimation_b200.testing import synth  # noqa: E402

ctx = pcl.Context(0)


def ds(points, leaf):
    vg = pcl.VoxelGrid(ctx)
    vg.setInputCloud(points)
    vg.setLeafSize(leaf)
    return vg.filter()


prob = synth.make_c4(scale=1.0, n_guesses=64, downsample=ds)
icp = pcl.IterativeClosestPoint(ctx)
icp.setMaximumIterations(2)
icp.getConvergeCriteria().setAbsoluteMSE(-1.0)
icp.setInputTarget(prob.target)
icp.setInputSource(prob.source)
icp.setMaxCorrespondenceDistance(0.02)
out = (C.c_ulonglong * 8)()
lib.peb_debug_coop_stats(out, 1)
for rows in (1024,):
    ctx.set_int("coop_max_rows", rows)
    icp.alignBatch(prob.guess[:64])
    lib.peb_debug_coop_stats(out, 1)
    v = np.array(list(out), np.float64)
    done = v[0] - v[1]
    print(f"coop_max_rows {rows}: patches {int(v[0])}, fallbacks {int(v[1])} ({100 * v[1] / v[0]:.1f} %), per verified patch: "
          f"rows {v[2] / done:.1f}, non-empty rows {v[3] / done:.1f}, staged points {v[4] / done:.1f}, lanes verifying {v[5] / done:.1f}; cold lanes {int(v[6])} in {int(v[7])} warps")
info = ctx.grid_info() if hasattr(ctx, "grid_info") else None
print(info)
