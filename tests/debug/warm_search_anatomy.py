"""Development analysis (CPU only): what a warm nearest-neighbour query of the batched ICP does, counted with the
library's own search loop compiled for the host (tests/host/host_check.cu : hc_warm_stats).

An ICP trajectory is simulated with scipy (cKDTree + float64 Kabsch; only the GEOMETRY of the queries matters here, not
bit-exactness) from a few of the C4 start poses; at every iteration >= 1 each query searches the ball of its previous
match.  Printed per iteration: grid rows in the ball's bounding box, rows that pass the slab test (each costs two
dependent cell_start loads and a point scan), points scanned, how often the match changes.  The question it answers:
how many DEPENDENT L2 round trips sit in the chain work[i] -> pts[j_prev] -> row bounds -> points, and what a variant
that issues all row-bound loads up front could save.

    python tests/debug/warm_search_anatomy.py [scale] [occupancy]
"""
import ctypes as C
import subprocess
import sys
from pathlib import Path

import numpy as np
from scipy.spatial import cKDTree

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from pose_estimation_b200.testing import synth  # noqa: E402

HOST = ROOT / "tests" / "host"


def kabsch(src, dst):
    cs, cd = src.mean(0), dst.mean(0)
    U, _, Vt = np.linalg.svd((dst - cd).T @ (src - cs))
    S = np.diag([1.0, 1.0, np.sign(np.linalg.det(U @ Vt))])
    R = U @ S @ Vt
    return R, cd - R @ cs


def numpy_voxel(points, leaf):
    p = points[np.isfinite(points[:, :3]).all(1), :3].astype(np.float64)
    key = np.floor(p / leaf).astype(np.int64)
    _, inv = np.unique(key, axis=0, return_inverse=True)
    inv = inv.reshape(-1)
    cnt = np.bincount(inv)
    out = np.stack([np.bincount(inv, p[:, a]) / cnt for a in range(3)], 1)
    return np.concatenate([out, np.ones((len(out), 1))], 1).astype(np.float32)


def main():
    scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.5
    occupancy = float(sys.argv[2]) if len(sys.argv) > 2 else 3.5
    subprocess.run(["make", "-C", str(HOST)], check=True, capture_output=True)
    L = C.CDLL(str(HOST / "libpe_hostcheck.so"))
    vp, sz = C.c_void_p, C.c_size_t
    L.hc_warm_stats.argtypes = [vp, sz, sz, vp, sz, sz, C.c_float, vp, C.c_float, vp, vp, vp]
    prob = synth.make_c4(scale=scale, n_guesses=4, downsample=numpy_voxel)
    tgt = np.ascontiguousarray(prob.target[:, :3], np.float32)
    src = prob.source[:, :3].astype(np.float64)
    tree = cKDTree(tgt.astype(np.float64))
    # warps: the library stores the source in 32-point patches (cells of its own sort grid, ~32 points each, x fastest)
    ext = np.sort(src.max(0) - src.min(0))[::-1]
    hs = np.sqrt(32.0 * ext[0] * ext[1] / len(src))
    cell = np.floor((src - src.min(0)) / hs).astype(np.int64)
    order = np.lexsort((cell[:, 0], cell[:, 1], cell[:, 2]))
    n_warp = len(src) // 32
    warp_of = order[: n_warp * 32].reshape(n_warp, 32)
    warp_big, warp_huge, warp_n = np.zeros(30, np.int64), np.zeros(30, np.int64), np.zeros(30, np.int64)
    print(f"scene {len(tgt)} points, model {len(src)} points, occupancy {occupancy}")
    limit = np.float32(0.02 ** 2)
    rows_hist = np.zeros(12, np.int64)
    box_big, box_huge, box_n = np.zeros(30, np.int64), np.zeros(30, np.int64), np.zeros(30, np.int64)
    total = np.zeros(5)
    for h in range(len(prob.guess)):
        T = np.asarray(prob.guess[h], np.float64)
        work = src @ T[:3, :3].T + T[:3, 3]
        prev = None
        for it in range(30):
            d, idx = tree.query(work)
            keep = d <= 0.02
            if it >= 1:
                q = np.ascontiguousarray(work, np.float32)
                out_idx = np.empty(len(q), np.int32)
                st = np.empty((len(q), 4), np.int32)
                cell = C.c_float(0)
                L.hc_warm_stats(tgt.ctypes.data, len(tgt), tgt.strides[0], q.ctypes.data, len(q), q.strides[0], occupancy,
                                prev.ctypes.data, limit, out_idx.ctypes.data, st.ctypes.data, C.byref(cell))
                changed = float((out_idx != prev).mean())
                if h == 0 and it in (1, 2, 5, 10, 20, 29):
                    r_cells = np.sqrt(np.minimum(((q - tgt[prev]) ** 2).sum(1), limit)) / cell.value
                    # the bound-only variant (nn_upfront.cuh : grid_nn_bounded_upfront) searches d_old + |movement| instead
                    r_bound = np.minimum(prev_d + np.linalg.norm(work - prev_work, axis=1), 0.02) / cell.value
                    print(f"         bound-only ball: {np.median(r_bound):.2f} cells median ({np.median(r_bound / np.maximum(r_cells, 1e-9)):.2f} x the "
                          f"candidate's ball), 2 x 2 rows suffice for {100 * np.mean(r_bound < 0.5):.0f} % (candidate: {100 * np.mean(r_cells < 0.5):.0f} %)")
                    print(f"  it {it:2d}: ball radius {np.median(r_cells):.2f} cells (median; cell {1e3 * cell.value:.2f} mm)  rows in box "
                          f"{st[:, 0].mean():.2f}  rows scanned {st[:, 1].mean():.2f}  points scanned {st[:, 2].mean():.1f}  "
                          f"match changed {100 * changed:.1f} %")
                box_big[it] += int((~np.isin(st[:, 0], (1, 2, 4))).sum())   # 3 x 1, 3 x 2, ...: more than 2 x 2 rows in the box
                box_huge[it] += int((st[:, 0] > 9).sum())                       # more than 3 x 3
                box_n[it] += len(q)
                warp_big[it] += int((~np.isin(st[:, 0], (1, 2, 4)))[warp_of].any(1).sum())
                warp_huge[it] += int((st[:, 0] > 9)[warp_of].any(1).sum())
                warp_n[it] += n_warp
                rows_hist += np.bincount(np.minimum(st[:, 1], 11), minlength=12)
                total += [len(q), st[:, 0].sum(), st[:, 1].sum(), st[:, 2].sum(), changed * len(q)]
            prev = idx.astype(np.int32)
            prev_d, prev_work = d.copy(), work.copy()
            R, t = kabsch(work[keep], tgt[idx[keep]].astype(np.float64))
            work = work @ R.T + t
    n = total[0]
    print(f"all warm queries ({int(n)}): rows in box {total[1] / n:.2f}, rows scanned {total[2] / n:.2f}, points scanned "
          f"{total[3] / n:.1f}, match changed {100 * total[4] / n:.1f} %")
    print("rows scanned per query, share of the queries: " + "  ".join(f"{k}{'+' if k == 11 else ''}: {100 * v / n:.1f} %"
                                                                       for k, v in enumerate(rows_hist) if v))
    print("queries whose box exceeds 2 x 2 rows (the up-front variant's fallback), by iteration: " +
          "  ".join(f"{it}: {100 * box_big[it] / box_n[it]:.0f} %" for it in (1, 2, 3, 5, 10, 20, 29)) +
          f"  all: {100 * box_big.sum() / box_n.sum():.1f} %;  exceeding 3 x 3: {100 * box_huge.sum() / box_n.sum():.2f} %")
    print("WARPS (32-point source patches) with at least one lane beyond 2 x 2 rows, by iteration: " +
          "  ".join(f"{it}: {100 * warp_big[it] / warp_n[it]:.0f} %" for it in (1, 2, 3, 5, 10, 20, 29)) +
          f"  all: {100 * warp_big.sum() / warp_n.sum():.0f} %;  beyond 3 x 3: " +
          "  ".join(f"{it}: {100 * warp_huge[it] / warp_n[it]:.0f} %" for it in (1, 2, 3, 5, 10, 20, 29)) +
          f"  all: {100 * warp_huge.sum() / warp_n.sum():.0f} %")
    # dependent L2 round trips of one query: work[i], pts[j_prev], then per scanned row bounds -> points, row after row
    serial = 2 + 2 * total[2] / n
    upfront = 2 + 2  # all row bounds at once, then all point ranges
    print(f"dependent load levels per query: {serial:.2f} today (row after row), {upfront} with the row bounds of the initial "
          f"ball issued up front")


if __name__ == "__main__":
    main()
