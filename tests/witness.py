"""A SECOND, independent restatement of the PCL 1.10 functions on the path — numpy, written from the semantics listed in
SURVEY.md 8(a) (not from oracle/pcl_oracle.cpp), with the nearest neighbours taken from a real FLANN
(cv2.flann_Index, KDTREE_SINGLE, leaf 15, exact: the kd-tree family PCL's KdTreeFLANN wraps).

TEST INFRASTRUCTURE.  Purpose: the C++ oracle restates PCL from recollection and nothing in the image can run PCL, so
every function of the oracle gets a second witness that was written separately: if the two disagree, one of them has
mis-stated the loop order, a comparison, a float / double choice or the state machine (tests/test_witness.py).

Float arithmetic: numpy's element-wise float32 operations are correctly rounded and never fused, and np.cumsum adds
strictly left to right, so `cumsum(x)[-1]` is the sequential float sum PCL's loops (and Eigen's scalar paths) compute.
The 3 x 3 SVD is numpy's (LAPACK, float64 on the float32 matrix) — a different algorithm from Eigen's JacobiSVD, which
is the point; increments therefore agree with the oracle's float path to float rounding, not bit for bit.
"""
from __future__ import annotations

import numpy as np

F = np.float32
DBL_MAX = float(np.finfo(np.float64).max)

NOT_CONVERGED, ITERATIONS, TRANSFORM, ABS_MSE, REL_MSE, NO_CORRESPONDENCES, FAILURE_AFTER_MAX_ITERATIONS = range(7)


def seq_sum(x: np.ndarray, dtype) -> np.ndarray:
    """Left-to-right sum over axis 0 in `dtype` (what `for (...) acc += x[i]` computes)."""
    x = np.asarray(x, dtype)
    if x.shape[0] == 0:
        return np.zeros(x.shape[1:], dtype)
    return np.cumsum(x, axis=0, dtype=dtype)[-1]


# ---- [PCL] kdtree/impl/kdtree_flann.hpp over [FLANN] KDTreeSingleIndex<L2_Simple<float>> -----------------------------
class FlannTree:
    """Finite points only, results as ORIGINAL indices (PCL's index_mapping_), squared distances, ascending."""

    def __init__(self, pts: np.ndarray):
        import cv2

        xyz = np.ascontiguousarray(np.asarray(pts, F)[:, :3])
        self.map = np.flatnonzero(np.isfinite(xyz).all(1)).astype(np.int32)
        self.xyz = np.ascontiguousarray(xyz[self.map])
        self.index = cv2.flann_Index(self.xyz, {"algorithm": 4, "leaf_max_size": 15}) if len(self.xyz) else None

    def knn(self, q: np.ndarray, k: int):
        q = np.ascontiguousarray(np.asarray(q, F)[:, :3])
        k = min(k, len(self.xyz))
        idx, _ = self.index.knnSearch(q, k, params={"checks": -1, "eps": 0.0, "sorted": True})
        idx = idx.astype(np.int64).reshape(len(q), k)
        # L2_Simple, recomputed here: ((dx*dx) + dy*dy) + dz*dz in float
        d = q[:, None, :] - self.xyz[idx]
        d2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]
        return self.map[idx], d2.astype(F)


# ---- transforms ----------------------------------------------------------------------------------------------------
def transform_icp(T: np.ndarray, p: np.ndarray) -> np.ndarray:
    """[PCL] registration/impl/icp.hpp transformCloud: Eigen 4x4 * (x, y, z, 1), ((c0 x + c1 y) + c2 z) + c3."""
    T = np.asarray(T, F)
    x, y, z = p[:, 0], p[:, 1], p[:, 2]
    return np.stack([((T[r, 0] * x + T[r, 1] * y) + T[r, 2] * z) + T[r, 3] * F(1) for r in range(3)], 1).astype(F)


def transform_tpc(T: np.ndarray, p: np.ndarray) -> np.ndarray:
    """[PCL] common/impl/transforms.hpp Transformer::se3: c0 x + (c1 y + (c2 z + c3))."""
    T = np.asarray(T, F)
    x, y, z = p[:, 0], p[:, 1], p[:, 2]
    return np.stack([x * T[r, 0] + (y * T[r, 1] + (z * T[r, 2] + T[r, 3])) for r in range(3)], 1).astype(F)


def mul4(A: np.ndarray, B: np.ndarray) -> np.ndarray:
    """Matrix4f product, ((a0 b0 + a1 b1) + a2 b2) + a3 b3 per entry, float."""
    A, B = np.asarray(A, F), np.asarray(B, F)
    out = np.empty((4, 4), F)
    for i in range(4):
        for j in range(4):
            out[i, j] = ((A[i, 0] * B[0, j] + A[i, 1] * B[1, j]) + A[i, 2] * B[2, j]) + A[i, 3] * B[3, j]
    return out


# ---- [PCL] common/impl/eigen.hpp umeyama(src, dst, with_scaling = false) ------------------------------------------
def umeyama(src: np.ndarray, dst: np.ndarray, dtype=F) -> np.ndarray:
    """src, dst: (n, 3).  Means, demeaning and sigma in `dtype` with left-to-right sums (float in PCL); the SVD and
    the small products after it in float64 (numpy), rounded to float at the end."""
    n = len(src)
    s, d = np.asarray(src, dtype), np.asarray(dst, dtype)
    one_over_n = dtype(1) / dtype(n)
    sm = seq_sum(s, dtype) * one_over_n
    dm = seq_sum(d, dtype) * one_over_n
    sd, dd = (s - sm).astype(dtype), (d - dm).astype(dtype)
    sigma = np.empty((3, 3), dtype)
    for r in range(3):
        for c in range(3):
            sigma[r, c] = seq_sum(dd[:, r] * sd[:, c], dtype) * one_over_n
    sig = sigma.astype(np.float64)
    U, sv, Vt = np.linalg.svd(sig)
    S = np.ones(3)
    if np.linalg.det(sig) < 0:
        S[2] = -1
    prec = 1e-5 if dtype == F else 1e-12
    rank = int(np.sum(~(np.abs(sv) <= np.abs(sv[0]) * prec)))
    if rank == 2:
        if np.linalg.det(U) * np.linalg.det(Vt) > 0:
            R = U @ Vt
        else:
            R = U @ np.diag([1.0, 1.0, -1.0]) @ Vt
    else:
        R = U @ np.diag(S) @ Vt
    T = np.eye(4, dtype=F)
    T[:3, :3] = R.astype(F)
    T[:3, 3] = (dm.astype(np.float64) - R @ sm.astype(np.float64)).astype(F)
    return T


# ---- [PCL] registration/impl/transformation_estimation_point_to_plane_lls.hpp ---------------------------------------
def point_to_plane_lls(src: np.ndarray, dst: np.ndarray, nrm: np.ndarray) -> np.ndarray:
    """Pairs with a non-finite member are skipped; a b c d in FLOAT, widened, 6 x 6 normal equations in double with
    left-to-right sums, x = inverse(ATA) ATb, R = Rz(gamma) Ry(beta) Rx(alpha)."""
    ok = np.isfinite(src).all(1) & np.isfinite(dst).all(1) & np.isfinite(nrm).all(1)
    s, d, n = np.asarray(src, F)[ok], np.asarray(dst, F)[ok], np.asarray(nrm, F)[ok]
    sx, sy, sz = s.T
    dx, dy, dz = d.T
    nx, ny, nz = n.T
    a = nz * sy - ny * sz
    b = nx * sz - nz * sx
    c = ny * sx - nx * sy
    dd = nx * dx + ny * dy + nz * dz - nx * sx - ny * sy - nz * sz  # left to right, float
    # the n n^T block of ATA is formed from FLOAT products (nx * ny is evaluated in float, then widened)
    J = np.stack([a, b, c, nx, ny, nz], 1).astype(np.float64)
    ATA = np.zeros((6, 6))
    for i in range(6):
        for j in range(i, 6):
            if i >= 3 and j >= 3:
                prod = (n[:, i - 3] * n[:, j - 3]).astype(np.float64)
            else:
                prod = J[:, i] * J[:, j]
            ATA[i, j] = ATA[j, i] = seq_sum(prod, np.float64)
    ATb = np.array([seq_sum(J[:, i] * dd.astype(np.float64), np.float64) for i in range(6)])
    x = np.linalg.inv(ATA) @ ATb
    al, be, ga = x[0], x[1], x[2]
    T = np.eye(4, dtype=F)
    T[0, 0] = F(np.cos(ga) * np.cos(be))
    T[0, 1] = F(-np.sin(ga) * np.cos(al) + np.cos(ga) * np.sin(be) * np.sin(al))
    T[0, 2] = F(np.sin(ga) * np.sin(al) + np.cos(ga) * np.sin(be) * np.cos(al))
    T[1, 0] = F(np.sin(ga) * np.cos(be))
    T[1, 1] = F(np.cos(ga) * np.cos(al) + np.sin(ga) * np.sin(be) * np.sin(al))
    T[1, 2] = F(-np.cos(ga) * np.sin(al) + np.sin(ga) * np.sin(be) * np.cos(al))
    T[2, 0] = F(-np.sin(be))
    T[2, 1] = F(np.cos(be) * np.sin(al))
    T[2, 2] = F(np.cos(be) * np.cos(al))
    T[:3, 3] = x[3:6].astype(F)
    return T


# ---- [PCL] registration/impl/default_convergence_criteria.hpp -----------------------------------------------------
class Criteria:
    """hasConverged() in PCL's order: iterations, transform, absolute MSE, relative MSE; the similar-transforms gate."""

    def __init__(self, max_iterations, rotation_threshold, translation_threshold, mse_rel, mse_abs, max_similar=0,
                 failure_after_max_iter=False):
        self.max_iterations = max_iterations
        self.rot_thr, self.trans_thr, self.mse_rel, self.mse_abs = rotation_threshold, translation_threshold, mse_rel, mse_abs
        self.max_similar, self.failure_after_max_iter = max_similar, failure_after_max_iter
        self.similar = 0
        self.prev_mse = self.cur_mse = DBL_MAX
        self.state = NOT_CONVERGED

    def has_converged(self, iterations: int, inc: np.ndarray, corr_d2: np.ndarray) -> bool:
        if self.state != NOT_CONVERGED:  # a new run of the criteria object
            self.similar = 0
            self.state = NOT_CONVERGED
        similar_now = False
        if iterations >= self.max_iterations:
            if not self.failure_after_max_iter:
                self.state = ITERATIONS
                return True
            self.state = FAILURE_AFTER_MAX_ITERATIONS
        inc = np.asarray(inc, F)
        cos_angle = 0.5 * float((inc[0, 0] + inc[1, 1] + inc[2, 2]) - F(1))                    # float trace, then double
        transl = float((inc[0, 3] * inc[0, 3] + inc[1, 3] * inc[1, 3]) + inc[2, 3] * inc[2, 3])  # float expression
        if cos_angle >= self.rot_thr and transl <= self.trans_thr:
            if self.similar >= self.max_similar:
                self.state = TRANSFORM
                return True
            similar_now = True
        with np.errstate(all="ignore"):
            self.cur_mse = float(seq_sum(np.asarray(corr_d2, np.float64), np.float64) / np.float64(len(corr_d2)))
            diff = abs(self.cur_mse - self.prev_mse)
            if diff < self.mse_abs:
                if self.similar >= self.max_similar:
                    self.state = ABS_MSE
                    return True
                similar_now = True
            if np.float64(diff) / np.float64(self.prev_mse) < self.mse_rel:
                if self.similar >= self.max_similar:
                    self.state = REL_MSE
                    return True
                similar_now = True
        self.similar = self.similar + 1 if similar_now else 0
        self.prev_mse = self.cur_mse
        return False


# ---- [PCL] registration/impl/icp.hpp + registration.hpp + correspondence_estimation.hpp ---------------------------
class ICP:
    def __init__(self, target: np.ndarray, normals: np.ndarray | None = None):
        self.tgt = np.ascontiguousarray(np.asarray(target, F)[:, :3])
        self.nrm = None if normals is None else np.ascontiguousarray(np.asarray(normals, F)[:, :3])
        self.tree = FlannTree(self.tgt)

    def correspondences(self, work: np.ndarray, valid: np.ndarray, max_corr_dist: float, rejector_max_dist: float):
        """determineCorrespondences (d2 > max^2 dropped, compared in double) + CorrespondenceRejectorDistance (keep iff
        d2 < max^2, the limit stored as a float).  -> query indices, match indices, squared distances, query order."""
        q = np.flatnonzero(valid)
        idx, d2 = self.tree.knn(work[q], 1)
        idx, d2 = idx[:, 0], d2[:, 0]
        keep = ~(d2.astype(np.float64) > max_corr_dist * max_corr_dist)
        if rejector_max_dist > 0:
            keep &= d2 < F(rejector_max_dist * rejector_max_dist)
        return q[keep], idx[keep], d2[keep]

    def estimate(self, work, q, m, estimator, wide):
        if estimator == 0:
            return umeyama(work[q], self.tgt[m], np.float64 if wide else F)
        return point_to_plane_lls(work[q], self.tgt[m], self.nrm[m])

    def fitness(self, source: np.ndarray, T: np.ndarray, max_range: float = DBL_MAX) -> float:
        """getFitnessScore: transformPointCloud's association; d2 <= max_range (un-squared: PCL's quirk)."""
        src = np.asarray(source, F)[:, :3]
        ok = np.isfinite(src).all(1)
        _, d2 = self.tree.knn(transform_tpc(T, src[ok]), 1)
        d2 = d2[:, 0].astype(np.float64)
        d2 = d2[d2 <= max_range]
        return float(seq_sum(d2, np.float64) / len(d2)) if len(d2) else DBL_MAX

    def align(self, source, guess=None, *, max_iterations=10, max_corr_dist=np.sqrt(DBL_MAX), transformation_epsilon=0.0,
              rotation_epsilon=0.0, euclidean_fitness_epsilon=-DBL_MAX, abs_mse_threshold=1e-12, min_correspondences=3,
              estimator=0, rejector_max_dist=0.0, max_iterations_similar=0, wide=False, forced_increments=None):
        """forced_increments: a list of 4x4 increments; iteration k computes its own increment (reported in the trace) but
        MOVES by forced_increments[k] — the lock-step mode of tests/test_witness.py."""
        src = np.asarray(source, F)[:, :3]
        valid = np.isfinite(src).all(1)
        final = np.eye(4, dtype=F) if guess is None else np.asarray(guess, F).copy()
        work = src.copy()
        if not np.array_equal(final, np.eye(4, dtype=F)):
            work[valid] = transform_icp(final, src[valid])
        crit = Criteria(max_iterations, rotation_epsilon if rotation_epsilon > 0 else 1.0 - transformation_epsilon,
                        transformation_epsilon, euclidean_fitness_epsilon, abs_mse_threshold, max_iterations_similar)
        trace = {"inc": [], "mse": [], "ncorr": [], "match": []}
        nr, converged = 0, False
        while True:
            q, m, d2 = self.correspondences(work, valid, max_corr_dist, rejector_max_dist)
            trace["ncorr"].append(len(q))
            trace["match"].append((q.copy(), m.copy()))
            if len(q) < min_correspondences:
                crit.state = NO_CORRESPONDENCES
                converged = False
                break
            inc = self.estimate(work, q, m, estimator, wide)
            trace["inc"].append(inc.copy())
            if forced_increments is not None and nr < len(forced_increments):
                inc = np.asarray(forced_increments[nr], F)
            work[valid] = transform_icp(inc, work[valid])
            final = mul4(inc, final)
            nr += 1
            converged = crit.has_converged(nr, inc, d2)
            trace["mse"].append(crit.cur_mse)
            if crit.state != NOT_CONVERGED:
                break
        return {"T": final, "iterations": nr, "converged": converged, "state": crit.state, "last_mse": crit.cur_mse,
                "n_correspondences": trace["ncorr"][-1], "trace": trace}


# ---- [PCL] filters/impl/voxel_grid.hpp ----------------------------------------------------------------------------------
def voxel_grid(points: np.ndarray, leaf: float, min_pts: int = 0):
    """-> (centroids (m, 4) float32, unchanged).  Sort by (voxel id, original index); sequential float centroid."""
    p = np.asarray(points, F)
    xyz = p[:, :3]
    fin = np.isfinite(xyz).all(1)
    if not fin.any():
        return np.empty((0, 4), F), False
    inv = F(1) / F(leaf)
    mn, mx = xyz[fin].min(0), xyz[fin].max(0)
    d = ((mx - mn) * inv).astype(np.int64) + 1
    if int(d[0]) * int(d[1]) * int(d[2]) > np.iinfo(np.int32).max:
        out = p[:, :4].copy()
        out[:, 3] = 1
        return out, True
    min_b = np.floor(mn * inv).astype(np.int32)
    max_b = np.floor(mx * inv).astype(np.int32)
    div = (max_b.astype(np.int64) - min_b + 1)
    mul = np.array([1, div[0], div[0] * div[1]], np.int64)
    ijk = (np.floor(xyz[fin] * inv) - min_b.astype(F)).astype(np.int32).astype(np.int64)
    vid = ijk @ mul
    orig = np.flatnonzero(fin)
    order = np.lexsort((orig, vid))
    vid, orig = vid[order], orig[order]
    starts = np.flatnonzero(np.r_[True, vid[1:] != vid[:-1]])
    ends = np.r_[starts[1:], len(vid)]
    out = []
    for s, e in zip(starts, ends):
        if e - s < min_pts:
            continue
        c = seq_sum(xyz[orig[s:e]], F) / F(e - s)
        out.append([c[0], c[1], c[2], 1.0])
    return np.asarray(out, F).reshape(-1, 4), False


# ---- [PCL] features/impl/normal_3d.hpp (1.10: single-pass float covariance, closed-form eigen33) -----------------------
def _roots2(b, c):
    roots = [F(0), F(0), F(0)]
    d = F(b * b - F(4) * c)
    if d < 0:
        d = F(0)
    sd = np.sqrt(d, dtype=F)
    roots[2] = F(F(0.5) * (b + sd))
    roots[1] = F(F(0.5) * (b - sd))
    return roots


def _roots(m):
    c0 = F(m[0, 0] * m[1, 1] * m[2, 2] + F(2) * m[0, 1] * m[0, 2] * m[1, 2] - m[0, 0] * m[1, 2] * m[1, 2]
           - m[1, 1] * m[0, 2] * m[0, 2] - m[2, 2] * m[0, 1] * m[0, 1])
    c1 = F(m[0, 0] * m[1, 1] - m[0, 1] * m[0, 1] + m[0, 0] * m[2, 2] - m[0, 2] * m[0, 2] + m[1, 1] * m[2, 2] - m[1, 2] * m[1, 2])
    c2 = F(m[0, 0] + m[1, 1] + m[2, 2])
    if abs(c0) < np.finfo(F).eps:
        return _roots2(c2, c1)
    s_inv3, s_sqrt3 = F(1.0 / 3.0), np.sqrt(F(3))
    c2_over_3 = F(c2 * s_inv3)
    a_over_3 = F((c1 - c2 * c2_over_3) * s_inv3)
    if a_over_3 > 0:
        a_over_3 = F(0)
    half_b = F(F(0.5) * (c0 + c2_over_3 * (F(2) * c2_over_3 * c2_over_3 - c1)))
    q = F(half_b * half_b + a_over_3 * a_over_3 * a_over_3)
    if q > 0:
        q = F(0)
    rho = np.sqrt(-a_over_3, dtype=F)
    theta = F(np.arctan2(np.sqrt(-q, dtype=F), half_b, dtype=F) * s_inv3)
    ct, st = np.cos(theta, dtype=F), np.sin(theta, dtype=F)
    r = sorted([F(c2_over_3 + F(2) * rho * ct), F(c2_over_3 - rho * (ct + s_sqrt3 * st)), F(c2_over_3 - rho * (ct - s_sqrt3 * st))])
    if r[0] <= 0:
        return _roots2(c2, c1)
    return r


def normal_of(neigh: np.ndarray, p: np.ndarray, viewpoint=(0.0, 0.0, 0.0)):
    """neigh: the k neighbours in kd-tree order (ascending distance).  -> (nx, ny, nz, curvature)."""
    n = len(neigh)
    if n < 3:
        return np.full(4, np.nan, F)
    x, y, z = np.asarray(neigh, F)[:, :3].T
    acc = np.array([seq_sum(v, F) for v in (x * x, x * y, x * z, y * y, y * z, z * z, x, y, z)], F) / F(n)
    C = np.empty((3, 3), F)
    C[0, 0] = acc[0] - acc[6] * acc[6]
    C[0, 1] = C[1, 0] = acc[1] - acc[6] * acc[7]
    C[0, 2] = C[2, 0] = acc[2] - acc[6] * acc[8]
    C[1, 1] = acc[3] - acc[7] * acc[7]
    C[1, 2] = C[2, 1] = acc[4] - acc[7] * acc[8]
    C[2, 2] = acc[5] - acc[8] * acc[8]
    scale = np.abs(C).max()
    if scale <= np.finfo(F).tiny:
        scale = F(1)
    S = (C / scale).astype(F)
    lam = F(_roots(S)[0])
    eigenvalue = F(lam * scale)
    S = S.copy()
    S[0, 0] -= lam
    S[1, 1] -= lam
    S[2, 2] -= lam
    cands = [np.cross(S[0], S[1]).astype(F), np.cross(S[0], S[2]).astype(F), np.cross(S[1], S[2]).astype(F)]
    lens = [float(np.dot(v.astype(np.float64), v.astype(np.float64))) for v in cands]
    v = cands[int(np.argmax(lens))]
    v = (v / np.sqrt(F(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]), dtype=F)).astype(F)
    tr = F(C[0, 0] + C[1, 1] + C[2, 2])
    curv = F(abs(eigenvalue / tr)) if tr != 0 else F(0)
    vp = np.asarray(viewpoint, F) - np.asarray(p, F)[:3]
    if F(vp[0] * v[0] + vp[1] * v[1] + vp[2] * v[2]) < 0:
        v = -v
    return np.array([v[0], v[1], v[2], curv], F)


def normals(points: np.ndarray, k: int, viewpoint=(0.0, 0.0, 0.0)):
    """-> ((n, 4) normals + curvature, (n, k) neighbour lists); non-finite points get NaN."""
    p = np.asarray(points, F)[:, :3]
    tree = FlannTree(p)
    out = np.full((len(p), 4), np.nan, F)
    ok = np.flatnonzero(np.isfinite(p).all(1))
    idx, _ = tree.knn(p[ok], k)
    for row, i in zip(idx, ok):
        out[i] = normal_of(p[row], p[i], viewpoint)
    nn = np.full((len(p), idx.shape[1]), -1, np.int32)
    nn[ok] = idx
    return out, nn
