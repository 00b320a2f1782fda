import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from oracle import Oracle, default_params
from pose_estimation_b200 import pcl
from pose_estimation_b200.testing import synth
from util import pose_delta
o = Oracle()
ctx = pcl.Context(0)
p = synth.make_c1(20000, seed=1)
for nit in (1, 2, 30):
    prm = default_params(max_iterations=nit, abs_mse_threshold=-1.0)
    icp = pcl.IterativeClosestPoint(ctx); icp.setInputSource(p.source); icp.setInputTarget(p.target)
    for name, _ in prm._fields_: setattr(icp.params, name, getattr(prm, name))
    icp.align(want_correspondences=True)
    ref = o.icp(p.target, wide_accum=True).align(p.source, None, prm, trace_cap=nit)
    tg = icp.trace(); to = ref["trace_T"]
    idx, d2 = icp.correspondences
    print("nit", nit, "corr idx equal", (idx == ref["corr_idx"]).mean(), "d2 equal", (d2 == ref["corr_d2"]).mean(),
          "mse", icp.result.last_mse, ref["result"].last_mse)
    for k in range(min(len(tg), len(to))):
        bits = (tg[k].view(np.uint32) != to[k].view(np.uint32)).sum()
        print("  it", k, "T entries differing", bits, "max abs diff", np.abs(tg[k] - to[k]).max(), pose_delta(tg[k], to[k]))
        if k > 6 and nit == 30 and k < 27: continue
