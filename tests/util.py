"""Helpers shared by the parity tests."""
import numpy as np


def tie_ok(target, query, idx_a, idx_b, tol=1e-6) -> bool:
    """True iff wherever the two index arrays differ, both matches are equidistant from the
    query within `tol` metres (the north star's allowance for equidistant ties)."""
    idx_a = np.asarray(idx_a).reshape(-1)
    idx_b = np.asarray(idx_b).reshape(-1)
    bad = np.flatnonzero(idx_a != idx_b)
    if bad.size == 0:
        return True
    if (idx_a[bad] < 0).any() or (idx_b[bad] < 0).any():
        return False
    t = np.asarray(target, np.float64)[:, :3]
    q = np.asarray(query, np.float64)[:, :3]
    q = q[bad] if q.shape[0] == idx_a.shape[0] else np.repeat(q, idx_a.shape[0] // q.shape[0], axis=0)[bad]
    da = np.linalg.norm(t[idx_a[bad]] - q, axis=1)
    db = np.linalg.norm(t[idx_b[bad]] - q, axis=1)
    return bool((np.abs(da - db) <= tol).all())


def pose_delta(A, B):
    """(rotation angle in rad, translation distance in m) between two 4x4 transforms."""
    A = np.asarray(A, np.float64)
    B = np.asarray(B, np.float64)
    R = A[:3, :3] @ B[:3, :3].T
    # robust angle for tiny rotations: |R - R^T| / 2 = sin(angle)
    s = 0.5 * np.linalg.norm([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    c = 0.5 * (np.trace(R) - 1.0)
    return float(np.arctan2(s, c)), float(np.linalg.norm(A[:3, 3] - B[:3, 3]))
