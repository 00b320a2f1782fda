"""The step right after the path: which refined hypothesis the node returns and
how it is packed into the 7 floats it publishes.  Host-side mirror of
pose_estimation/src/opencv_surface_match.cpp:100-142 (same functions as in include/pe_b200/pcl_facade.hpp);
no device work and no arithmetic of the hot path."""
from __future__ import annotations

import numpy as np


def select_best_pose(num_votes, residual, n_results_total: int) -> int:
    """opencv_surface_match.cpp:100-124: most votes wins; with more than 5 matcher results the lowest
    residual among poses with more than 400 votes overrides it (last assignment wins, as written)."""
    max_votes = 0
    min_res = np.float32(10000.0)
    best = 0
    for i, (v, r) in enumerate(zip(num_votes, residual)):
        if v > max_votes:
            max_votes = v
            best = i
        if n_results_total > 5 and r < min_res and v > 400:
            min_res = np.float32(r)
            best = i
    return best


def rotation_to_quat(T) -> np.ndarray:
    """4x4 (row-indexed) -> unit quaternion (w, x, y, z), w >= 0."""
    m = np.asarray(T, np.float64)[:3, :3]
    tr = m[0, 0] + m[1, 1] + m[2, 2]
    if tr > 0:
        s = np.sqrt(tr + 1.0) * 2
        q = [0.25 * s, (m[2, 1] - m[1, 2]) / s, (m[0, 2] - m[2, 0]) / s, (m[1, 0] - m[0, 1]) / s]
    elif m[0, 0] > m[1, 1] and m[0, 0] > m[2, 2]:
        s = np.sqrt(1.0 + m[0, 0] - m[1, 1] - m[2, 2]) * 2
        q = [(m[2, 1] - m[1, 2]) / s, 0.25 * s, (m[0, 1] + m[1, 0]) / s, (m[0, 2] + m[2, 0]) / s]
    elif m[1, 1] > m[2, 2]:
        s = np.sqrt(1.0 + m[1, 1] - m[0, 0] - m[2, 2]) * 2
        q = [(m[0, 2] - m[2, 0]) / s, (m[0, 1] + m[1, 0]) / s, 0.25 * s, (m[1, 2] + m[2, 1]) / s]
    else:
        s = np.sqrt(1.0 + m[2, 2] - m[0, 0] - m[1, 1]) * 2
        q = [(m[1, 0] - m[0, 1]) / s, (m[0, 2] + m[2, 0]) / s, (m[1, 2] + m[2, 1]) / s, 0.25 * s]
    q = np.array(q)
    return -q if q[0] < 0 else q


def pack_pose(T, reference_layout: bool = False) -> np.ndarray:
    """{x, y, z, qx, qy, qz, qw} float32.  reference_layout=True reproduces opencv_surface_match.cpp:133-142
    literally (qz is never written: pose[6] gets q[3] and is overwritten with q[0])."""
    T = np.asarray(T, np.float64)
    q = rotation_to_quat(T)
    pose = np.zeros(7, np.float32)
    pose[:3] = T[:3, 3]
    pose[3], pose[4] = q[1], q[2]
    if reference_layout:
        pose[6] = q[3]
        pose[6] = q[0]
    else:
        pose[5], pose[6] = q[3], q[0]
    return pose
