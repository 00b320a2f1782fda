import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import Oracle
from pose_estimation_b200 import pcl
from pose_estimation_b200.testing import synth
o = Oracle()
ctx = pcl.Context(0)
rng = np.random.default_rng(21)
surf = synth.Surface(21)
scene = synth.render_scene(surf, synth.default_gt_pose(rng), rng, 486, 300)
class P: pass
prob = P(); prob.target = scene
for leaf in (0.004, 0.002):
    vg = pcl.VoxelGrid(ctx); vg.setInputCloud(prob.target); vg.setLeafSize(leaf)
    a = vg.filter(); b, _ = o.voxel_grid(prob.target, leaf)
    print("leaf", leaf, a.shape, b.shape)
    sa = {tuple(r) for r in a.view(np.uint32).tolist()}; sb = {tuple(r) for r in b.view(np.uint32).tolist()}
    only_a = [np.array(r, np.uint32).view(np.float32) for r in (sa - sb)]
    only_b = [np.array(r, np.uint32).view(np.float32) for r in (sb - sa)]
    print(" only gpu:", len(only_a), only_a[:4]); print(" only oracle:", len(only_b), only_b[:4])
    n = min(len(a), len(b))
    d = np.flatnonzero((a[:n].view(np.uint32) != b[:n].view(np.uint32)).any(1))
    print(" first differing rows:", d[:5])
    for i in d[:2]:
        print("  gpu", a[i], "orc", b[i])
    # which input points fall in the voxels around the first difference
    if len(only_a) or len(only_b):
        inv = np.float32(1.0) / np.float32(leaf)
        p = prob.target[:, :3]
        with np.errstate(invalid="ignore"):
            ijk = np.nan_to_num(np.floor(p * inv), nan=1e9).astype(np.int64)
        for r in (only_a[:2] + only_b[:2]):
            c = np.floor(r[:3] * inv).astype(np.int64)
            m = (ijk == c).all(1)
            print("   voxel", c, "pts", np.flatnonzero(m), p[m])
