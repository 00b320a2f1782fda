"""Generates tests/golden/flann_nn.npz — nearest-neighbour golden vectors from a REAL FLANN.

PCL 1.10's KdTreeFLANN wraps flann::KDTreeSingleIndex<L2_Simple<float>> (leaf 15, exact search).
PCL/FLANN are not installed in this image, but OpenCV's bundled FLANN fork ships the same
KDTreeSingleIndex; cv2.flann_Index(algorithm=4 /*KDTREE_SINGLE*/, leaf_max_size=15) with
knnSearch(checks=-1, eps=0, sorted=True) is therefore the closest runnable stand-in for the
reference's neighbour search.  scipy's cKDTree is stored beside it as a second, unrelated
implementation.  Run once, here (needs cv2 + scipy); the .npz is committed and the tests only
read it.

    python tests/golden/make_golden.py
"""
from __future__ import annotations

import sys
from pathlib import Path

import cv2
import numpy as np
from scipy.spatial import cKDTree

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from pose_estimation_b200.testing import synth  # noqa: E402


def flann_knn(target: np.ndarray, queries: np.ndarray, k: int):
    index = cv2.flann_Index(np.ascontiguousarray(target, np.float32), {"algorithm": 4, "leaf_max_size": 15})
    idx, d2 = index.knnSearch(np.ascontiguousarray(queries, np.float32), k, params={"checks": -1, "eps": 0.0, "sorted": True})
    return idx.astype(np.int32), d2.astype(np.float32)


def main() -> None:
    out = {}
    # case A: the C1 surface clouds, reduced (camera-scale coordinates, surface-like density)
    p = synth.make_c1(n=3000, seed=11)
    tgt = p.target[:, :3].copy()
    qry = p.source[:600, :3].copy()
    out["a_target"], out["a_query"] = tgt, qry
    out["a_idx1"], out["a_d1"] = flann_knn(tgt, qry, 1)
    out["a_idx8"], out["a_d8"] = flann_knn(tgt, qry, 8)
    dd, ii = cKDTree(tgt.astype(np.float64)).query(qry.astype(np.float64), k=8)
    out["a_ckd_idx8"], out["a_ckd_d8"] = ii.astype(np.int32), (dd * dd)
    # case B: uniform random cube, queries partly outside the bounding box
    rng = np.random.default_rng(12)
    tgt = rng.uniform(-1, 1, (4000, 3)).astype(np.float32)
    qry = rng.uniform(-1.5, 1.5, (500, 3)).astype(np.float32)
    out["b_target"], out["b_query"] = tgt, qry
    out["b_idx1"], out["b_d1"] = flann_knn(tgt, qry, 1)
    out["b_idx30"], out["b_d30"] = flann_knn(tgt, qry, 30)
    dd, ii = cKDTree(tgt.astype(np.float64)).query(qry.astype(np.float64), k=30)
    out["b_ckd_idx30"], out["b_ckd_d30"] = ii.astype(np.int32), (dd * dd)
    np.savez_compressed(Path(__file__).with_name("flann_nn.npz"), **out)
    print("wrote flann_nn.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
