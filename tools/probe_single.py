"""Development probe: where the time of ONE single-align iteration launch goes (C2 data)."""
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


prob = synth.make_c2(downsample=ds)
icp = pcl.IterativeClosestPoint(ctx)
icp.setInputSource(prob.source)
icp.setInputTarget(prob.target)
icp.setMaximumIterations(30)
icp.getConvergeCriteria().setAbsoluteMSE(-1.0)
ctx.set_int("debug_timers", 1)
lib.peb_debug_timers_read.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
for rep in range(3):
    icp.align(prob.guess, want_output=False)
buf = np.zeros((64, 8), np.uint64)
n = C.c_size_t(0)
ctx.check(lib.peb_debug_timers_read(ctx.handle, buf.ctypes.data, 64, C.byref(n)))
t = buf[: n.value, :5].astype(np.int64)
inner = buf[: n.value, 5:8].astype(np.int64)
print("inside the solve (us): call+entry, state load, estimator+criteria, store+return")
for k in (1, 2, 10, 20, 29):
    print(f"  it {k}: entry {(inner[k,0]-t[k,3])/1e3:5.2f} load {(inner[k,1]-inner[k,0])/1e3:5.2f} math {(inner[k,2]-inner[k,1])/1e3:5.2f} tail {(t[k,4]-inner[k,2])/1e3:5.2f}")
print("per launch (us): last-block start -> NN loop done -> ticket won -> partials reduced -> solved ; gap to next launch's last-block start")
for k in range(n.value):
    d = np.diff(t[k]) / 1e3
    gap = (t[k + 1, 0] - t[k, 4]) / 1e3 if k + 1 < n.value else float("nan")
    print(f"  it {k:2d}: nn {d[0]:6.2f}  reduce+ticket(wait for all blocks) {d[1]:6.2f}  partials {d[2]:6.2f}  solve {d[3]:6.2f}  | gap {gap:6.2f}")
print("total first start -> last solve: %.1f us" % ((t[-1, 4] - t[0, 0]) / 1e3))
