"""Development probe: where the time of the reference-shaped cv ICP call goes (fixed cost vs iterations)."""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from pose_estimation_b200 import pcl  # noqa: E402
from pose_estimation_b200.testing.synth import make_cvicp_case as make_case  # noqa: E402

ctx = pcl.Context(0)
model, scene, poses, gt = make_case(seed=5, n_model=50000, n_scene=200000, clutter=20000)


def run(label, H=6, **kw):
    args = dict(iterations=250, tolerance=0.005, rejection_scale=2.5, num_levels=8)
    args.update(kw)
    icp = pcl.CvIcp(args["iterations"], args["tolerance"], args["rejection_scale"], args["num_levels"], ctx=ctx)
    icp.registerModelToScene(model, scene, poses[:H])
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        icp.registerModelToScene(model, scene, poses[:H])
        ts.append(1e3 * (time.perf_counter() - t0))
    print(f"{label:40s} {min(ts):8.2f} ms", flush=True)


run("reference call (6 poses, 8 levels)")
run("no iterations (set-up + 8 grids)", iterations=0)
run("1 level, no iterations", iterations=0, num_levels=1)
run("1 pose", H=1)
run("1 level", num_levels=1)
run("no robust rejection", rejection_scale=0.0)
run("tolerance 0.05", tolerance=0.05)
ctx.close()
