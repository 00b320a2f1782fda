"""CPU oracle loader — TEST INFRASTRUCTURE ONLY (see oracle/pcl_oracle.cpp header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product package (pose_estimation_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent


class IcpParams(C.Structure):
    """peb_icp_params (include/pe_b200.h)."""

    _fields_ = [
        ("max_iterations", C.c_int32),
        ("min_correspondences", C.c_int32),
        ("estimator", C.c_int32),
        ("max_iterations_similar", C.c_int32),
        ("max_corr_dist", C.c_double),
        ("transformation_epsilon", C.c_double),
        ("rotation_epsilon", C.c_double),
        ("euclidean_fitness_epsilon", C.c_double),
        ("abs_mse_threshold", C.c_double),
        ("rejector_max_dist", C.c_double),
        ("fitness_max_range", C.c_double),
    ]


class IcpResult(C.Structure):
    """peb_icp_result (include/pe_b200.h)."""

    _fields_ = [
        ("T", C.c_float * 16),
        ("fitness", C.c_double),
        ("last_mse", C.c_double),
        ("iterations", C.c_int32),
        ("converged", C.c_int32),
        ("state", C.c_int32),
        ("n_correspondences", C.c_int32),
    ]

    def matrix(self) -> np.ndarray:
        """4x4 row-indexed numpy matrix (T is column-major in the struct)."""
        return np.array(self.T, dtype=np.float32).reshape(4, 4).T.copy()


class PrefilterParams(C.Structure):
    """peb_prefilter_params (include/pe_b200.h)."""

    _fields_ = [
        ("use_sphere", C.c_int32),
        ("remove_inliers", C.c_int32),
        ("sphere_center", C.c_float * 3),
        ("sphere_radius", C.c_float),
        ("n_planes", C.c_int32),
        ("plane_band", C.c_float),
        ("planes", C.c_float * 32),
    ]


def prefilter_params(sphere=None, remove_inliers=False, planes=(), band=0.005) -> PrefilterParams:
    """sphere = (cx, cy, cz, radius) or None; planes = iterable of (a, b, c, d)."""
    f = PrefilterParams()
    if sphere is not None:
        f.use_sphere = 1
        f.remove_inliers = int(bool(remove_inliers))
        f.sphere_center[:] = [float(v) for v in sphere[:3]]
        f.sphere_radius = float(sphere[3])
    planes = list(planes)
    f.n_planes = len(planes)
    f.plane_band = float(band)
    for k, pl in enumerate(planes):
        for j in range(4):
            f.planes[4 * k + j] = float(pl[j])
    return f


class SacParams(C.Structure):
    """peb_sac_params (include/pe_b200.h)."""

    _fields_ = [
        ("distance_threshold", C.c_double),
        ("probability", C.c_double),
        ("max_iterations", C.c_int32),
        ("optimize_coefficients", C.c_int32),
        ("seed", C.c_uint32),
        ("reserved", C.c_int32),
    ]


def sac_params(distance_threshold=0.0, max_iterations=50, probability=0.99, optimize=True, seed=12345) -> SacParams:
    """PCL 1.10 SACSegmentation defaults."""
    return SacParams(float(distance_threshold), float(probability), int(max_iterations), int(bool(optimize)), int(seed), 0)


class CvIcpParams(C.Structure):
    """peb_cvicp_params (include/pe_b200.h)."""

    _fields_ = [("iterations", C.c_int32), ("num_levels", C.c_int32), ("tolerance", C.c_float), ("rejection_scale", C.c_float)]


def cvicp_params(iterations=250, tolerance=0.005, rejection_scale=2.5, num_levels=8) -> CvIcpParams:
    """cv::ppf_match_3d::ICP(iterations, tolerance, rejectionScale, numLevels) as the reference constructs it."""
    return CvIcpParams(int(iterations), int(num_levels), float(tolerance), float(rejection_scale))


class PpfParams(C.Structure):
    """peb_ppf_params (include/pe_b200.h)."""

    _fields_ = [("relative_sampling_step", C.c_double), ("relative_distance_step", C.c_double), ("num_angles", C.c_double),
                ("position_threshold", C.c_double), ("rotation_threshold", C.c_double), ("use_weighted_avg", C.c_int32),
                ("reserved", C.c_int32)]


class PpfPose(C.Structure):
    """peb_ppf_pose = cv::ppf_match_3d::Pose3D (include/pe_b200.h)."""

    _fields_ = [("pose", C.c_double * 16), ("q", C.c_double * 4), ("t", C.c_double * 3), ("angle", C.c_double),
                ("alpha", C.c_double), ("residual", C.c_double), ("num_votes", C.c_uint64), ("model_index", C.c_uint64)]

    @property
    def matrix(self) -> np.ndarray:
        return np.array(self.pose, np.float64).reshape(4, 4)


def ppf_params(relative_sampling_step=0.03, relative_distance_step=0.03, num_angles=40.0, position_threshold=-1.0,
               rotation_threshold=-1.0, use_weighted_avg=False) -> PpfParams:
    """cv::ppf_match_3d::PPF3DDetector(0.03, 0.03, 40) as the reference constructs it (opencv_surface_match.cpp:45)."""
    return PpfParams(float(relative_sampling_step), float(relative_distance_step), float(num_angles), float(position_threshold),
                     float(rotation_threshold), int(bool(use_weighted_avg)), 0)


DBL_MAX = float(np.finfo(np.float64).max)


def default_params(**kw) -> IcpParams:
    """PCL 1.10 defaults."""
    p = IcpParams(
        max_iterations=10,
        min_correspondences=3,
        estimator=0,
        max_iterations_similar=0,
        max_corr_dist=float(np.sqrt(np.float64(DBL_MAX))),
        transformation_epsilon=0.0,
        rotation_epsilon=0.0,
        euclidean_fitness_epsilon=-DBL_MAX,
        abs_mse_threshold=1e-12,
        rejector_max_dist=0.0,
        fitness_max_range=DBL_MAX,
    )
    for k, v in kw.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


def build(force: bool = False) -> None:
    """Compile the oracle's two shared objects with the committed Makefile."""
    if force:
        subprocess.run(["make", "-C", str(_HERE), "clean"], check=True, capture_output=True)
    subprocess.run(["make", "-C", str(_HERE)], check=True, capture_output=True)


def _cpu_has(*flags: str) -> bool:
    try:
        txt = Path("/proc/cpuinfo").read_text()
    except OSError:
        return False
    line = next((ln for ln in txt.splitlines() if ln.startswith("flags")), "")
    have = set(line.split())
    return all(f in have for f in flags)


_vp = C.c_void_p
_sz = C.c_size_t


def _as_f32(a) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a


class Oracle:
    """ctypes face of libpcl_oracle(.so|_fast.so)."""

    def __init__(self, fast: bool = False):
        name = "libpcl_oracle_fast.so" if fast else "libpcl_oracle.so"
        path = _HERE / name
        if not path.exists():
            build()
        if fast and not _cpu_has("avx2", "fma"):
            path = _HERE / "libpcl_oracle.so"  # the timing build needs x86-64-v3
        self.path = str(path)
        self.fast = fast and path.name.endswith("_fast.so")
        L = C.CDLL(self.path)
        self.L = L
        L.orc_version.restype = C.c_char_p
        L.orc_kdtree_create.restype = _vp
        L.orc_kdtree_create.argtypes = [_vp, _sz, _sz]
        L.orc_kdtree_destroy.argtypes = [_vp]
        L.orc_kdtree_knn.argtypes = [_vp, _vp, _sz, _sz, C.c_int, _vp, _vp]
        L.orc_nn_bruteforce.argtypes = [_vp, _sz, _sz, _vp, _sz, _sz, _vp, _vp]
        L.orc_voxel_grid.restype = _sz
        L.orc_voxel_grid.argtypes = [_vp, _sz, _sz, C.c_float, C.c_float, C.c_float, C.c_uint, _vp, _vp]
        L.orc_normals_knn.argtypes = [_vp, _sz, _sz, C.c_int, _vp, _vp, _vp, C.c_int]
        L.orc_icp_create.restype = _vp
        L.orc_icp_destroy.argtypes = [_vp]
        L.orc_icp_set_wide_accum.argtypes = [_vp, C.c_int]
        L.orc_icp_set_target.argtypes = [_vp, _vp, _sz, _sz, _vp, _sz]
        L.orc_icp_align.argtypes = [_vp, _vp, _sz, _sz, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]
        L.orc_icp_align_batch.argtypes = [_vp, _vp, _sz, _sz, _vp, _sz, _vp, _vp, C.c_int]
        L.orc_icp_fitness.restype = C.c_double
        L.orc_icp_fitness.argtypes = [_vp, _vp, _sz, _sz, _vp, C.c_double, _vp]
        L.orc_umeyama.argtypes = [_vp, _vp, _sz, C.c_int, _vp]
        L.orc_svd3f.argtypes = [_vp] * 4
        L.orc_svd3d.argtypes = [_vp] * 4
        L.orc_eigen33.argtypes = [_vp] * 3
        L.orc_inverse6.argtypes = [_vp] * 2
        L.orc_transform_icp.argtypes = [_vp] * 3
        L.orc_transform_tpc.argtypes = [_vp] * 3
        L.orc_mul4.argtypes = [_vp] * 3
        L.orc_point_to_plane_lls.argtypes = [_vp, _vp, _vp, _sz, _vp]
        L.orc_criteria_script.argtypes = [_vp, _vp, _vp, _sz, _vp, _vp]
        L.orc_max_threads.restype = C.c_int
        L.orc_scene_prefilter.restype = _sz
        L.orc_scene_prefilter.argtypes = [_vp, _sz, _sz, _vp, _vp]
        L.orc_sac_plane.restype = C.c_int
        L.orc_sac_plane.argtypes = [_vp, _sz, _sz, _vp, C.c_int, _vp, _vp, _vp, _vp]
        L.orc_cvicp_register.argtypes = [_vp, _sz, _vp, _sz, _vp, _vp, _sz, _vp]
        L.orc_ppf_train.restype = _vp
        L.orc_ppf_train.argtypes = [_vp, _sz, _vp]
        L.orc_ppf_destroy.argtypes = [_vp]
        L.orc_ppf_model_size.restype = _sz
        L.orc_ppf_model_size.argtypes = [_vp]
        L.orc_ppf_model_sampled.argtypes = [_vp, _vp]
        L.orc_ppf_distance_step.restype = C.c_double
        L.orc_ppf_distance_step.argtypes = [_vp]
        L.orc_ppf_table_size.restype = _sz
        L.orc_ppf_table_size.argtypes = [_vp]
        L.orc_ppf_match.restype = _sz
        L.orc_ppf_match.argtypes = [_vp, _vp, _sz, C.c_double, C.c_double, _vp, _sz, _vp, _vp, _sz, _vp, _sz, _vp]
        L.orc_ppf_feature.argtypes = [_vp] * 6
        L.orc_ppf_transform_rt.argtypes = [_vp] * 4
        L.orc_ppf_sample.restype = _sz
        L.orc_ppf_sample.argtypes = [_vp, _sz, C.c_float, _vp, _sz]
        L.orc_ppf_cluster.restype = _sz
        L.orc_ppf_cluster.argtypes = [_vp, _vp, _sz, _vp, _sz]
        L.orc_mt19937_nth.restype = C.c_uint32
        L.orc_mt19937_nth.argtypes = [C.c_uint32, C.c_uint32]

    # ---- nearest neighbours ------------------------------------------------------------
    def knn(self, target: np.ndarray, queries: np.ndarray, k: int = 1):
        t = _as_f32(target)
        q = _as_f32(queries)
        tree = self.L.orc_kdtree_create(t.ctypes.data, t.shape[0], t.strides[0])
        try:
            idx = np.empty((q.shape[0], k), np.int32)
            d2 = np.empty((q.shape[0], k), np.float32)
            self.L.orc_kdtree_knn(tree, q.ctypes.data, q.shape[0], q.strides[0], k, idx.ctypes.data, d2.ctypes.data)
        finally:
            self.L.orc_kdtree_destroy(tree)
        return idx, d2

    def nn_bruteforce(self, target: np.ndarray, queries: np.ndarray):
        t = _as_f32(target)
        q = _as_f32(queries)
        idx = np.empty(q.shape[0], np.int32)
        d2 = np.empty(q.shape[0], np.float32)
        self.L.orc_nn_bruteforce(t.ctypes.data, t.shape[0], t.strides[0], q.ctypes.data, q.shape[0], q.strides[0],
                                 idx.ctypes.data, d2.ctypes.data)
        return idx, d2

    # ---- VoxelGrid -----------------------------------------------------------------------
    def voxel_grid(self, pts: np.ndarray, leaf, min_pts: int = 0):
        p = _as_f32(pts)
        leaf = (leaf, leaf, leaf) if np.isscalar(leaf) else tuple(leaf)
        out = np.empty((p.shape[0], 4), np.float32)
        unchanged = C.c_int(0)
        m = self.L.orc_voxel_grid(p.ctypes.data, p.shape[0], p.strides[0], leaf[0], leaf[1], leaf[2], min_pts,
                                  out.ctypes.data, C.byref(unchanged))
        return out[:m].copy(), bool(unchanged.value)

    # ---- scene pre-filter (restates the reference's own code) ----------------------------------
    def scene_prefilter(self, pts: np.ndarray, params: "PrefilterParams") -> np.ndarray:
        p = _as_f32(pts)
        out = np.empty((max(p.shape[0], 1), 4), np.float32)
        m = self.L.orc_scene_prefilter(p.ctypes.data, p.shape[0], p.strides[0], C.byref(params), out.ctypes.data)
        return out[:m].copy()

    # ---- SACSegmentation, plane + RANSAC -------------------------------------------------------
    def sac_plane(self, pts: np.ndarray, params: "SacParams", wide_accum: bool = True):
        """-> (found, coefficients float32[4], inlier indices int32, iterations)"""
        p = _as_f32(pts)
        coeff = np.zeros(4, np.float32)
        inl = np.empty(max(p.shape[0], 1), np.int32)
        m = C.c_size_t(0)
        it = C.c_int32(0)
        found = self.L.orc_sac_plane(p.ctypes.data, p.shape[0], p.strides[0], C.byref(params), int(wide_accum),
                                     coeff.ctypes.data, inl.ctypes.data, C.byref(m), C.byref(it))
        return bool(found), coeff, inl[: m.value].copy(), it.value

    # ---- cv::ppf_match_3d::ICP::registerModelToScene -------------------------------------------
    def cvicp_register(self, model6: np.ndarray, scene6: np.ndarray, poses: np.ndarray, params: "CvIcpParams"):
        """model6 / scene6: (n, 6) float32 x y z nx ny nz; poses: (H, 4, 4) float64 -> (refined poses, residuals)."""
        m = np.ascontiguousarray(model6, np.float32)
        sc = np.ascontiguousarray(scene6, np.float32)
        P = np.ascontiguousarray(np.asarray(poses, np.float64).reshape(-1, 16)).copy()
        res = np.zeros(P.shape[0], np.float64)
        self.L.orc_cvicp_register(m.ctypes.data, m.shape[0], sc.ctypes.data, sc.shape[0], C.byref(params), P.ctypes.data,
                                  P.shape[0], res.ctypes.data)
        return P.reshape(-1, 4, 4), res

    # ---- cv::ppf_match_3d::PPF3DDetector ------------------------------------------------------
    def ppf_train(self, model6: np.ndarray, params: "PpfParams | None" = None) -> "OraclePpf":
        return OraclePpf(self, model6, params or ppf_params())

    def ppf_sample(self, pc6: np.ndarray, step: float) -> np.ndarray:
        """samplePCByQuantization over the cloud's own bounding box."""
        p = np.ascontiguousarray(pc6, np.float32)
        out = np.empty((p.shape[0], 6), np.float32)
        n = self.L.orc_ppf_sample(p.ctypes.data, p.shape[0], float(step), out.ctypes.data, out.shape[0])
        return out[:n].copy()

    def ppf_feature(self, p1, n1, p2, n2):
        """(f[4], alpha) of the ordered pair, double."""
        a = [np.ascontiguousarray(v, np.float64) for v in (p1, n1, p2, n2)]
        f = np.zeros(4, np.float64)
        al = C.c_double(0)
        self.L.orc_ppf_feature(a[0].ctypes.data, a[1].ctypes.data, a[2].ctypes.data, a[3].ctypes.data, f.ctypes.data, C.byref(al))
        return f, al.value

    def ppf_transform_rt(self, p1, n1):
        a = [np.ascontiguousarray(v, np.float64) for v in (p1, n1)]
        R = np.zeros(9, np.float64)
        t = np.zeros(3, np.float64)
        self.L.orc_ppf_transform_rt(a[0].ctypes.data, a[1].ctypes.data, R.ctypes.data, t.ctypes.data)
        return R.reshape(3, 3), t

    def mt19937_nth(self, seed: int, nth: int) -> int:
        return int(self.L.orc_mt19937_nth(seed, nth))

    # ---- NormalEstimation ----------------------------------------------------------------
    def normals(self, pts: np.ndarray, k: int, viewpoint=(0.0, 0.0, 0.0), threads: int = 1, want_nn: bool = False):
        p = _as_f32(pts)
        vp = np.asarray(viewpoint, np.float32)
        out = np.empty((p.shape[0], 8), np.float32)
        nn = np.empty((p.shape[0], k), np.int32) if want_nn else None
        self.L.orc_normals_knn(p.ctypes.data, p.shape[0], p.strides[0], k, vp.ctypes.data, out.ctypes.data,
                               nn.ctypes.data if want_nn else None, threads)
        return (out, nn) if want_nn else out

    # ---- ICP -----------------------------------------------------------------------------
    def icp(self, target: np.ndarray, normals: np.ndarray | None = None, wide_accum: bool = False) -> "OracleIcp":
        return OracleIcp(self, target, normals, wide_accum)

    def max_threads(self) -> int:
        return int(self.L.orc_max_threads())


class OraclePpf:
    """A trained detector (PPF3DDetector::trainModel)."""

    def __init__(self, orc: Oracle, model6: np.ndarray, params: "PpfParams"):
        self.o = orc
        m = np.ascontiguousarray(model6, np.float32)
        assert m.ndim == 2 and m.shape[1] == 6
        self.params = params
        self.h = orc.L.orc_ppf_train(m.ctypes.data, m.shape[0], C.byref(params))

    def __del__(self):
        if getattr(self, "h", None):
            self.o.L.orc_ppf_destroy(self.h)
            self.h = None

    @property
    def n(self) -> int:
        return int(self.o.L.orc_ppf_model_size(self.h))

    @property
    def distance_step(self) -> float:
        return float(self.o.L.orc_ppf_distance_step(self.h))

    @property
    def table_size(self) -> int:
        return int(self.o.L.orc_ppf_table_size(self.h))

    def sampled(self) -> np.ndarray:
        out = np.empty((self.n, 6), np.float32)
        self.o.L.orc_ppf_model_sampled(self.h, out.ctypes.data)
        return out

    def match(self, scene6: np.ndarray, relative_scene_sample_step: float = 1.0, relative_scene_distance: float = 0.03):
        """-> (clustered poses, raw per-reference poses, sampled scene)."""
        sc = np.ascontiguousarray(scene6, np.float32)
        cap = sc.shape[0]
        raw = (PpfPose * cap)()
        res = (PpfPose * cap)()
        samp = np.empty((cap, 6), np.float32)
        n_raw = C.c_size_t(0)
        n_s = C.c_size_t(0)
        n = self.o.L.orc_ppf_match(self.h, sc.ctypes.data, sc.shape[0], float(relative_scene_sample_step),
                                   float(relative_scene_distance), raw, cap, C.byref(n_raw), res, cap, samp.ctypes.data, cap,
                                   C.byref(n_s))
        return list(res[:n]), list(raw[: n_raw.value]), samp[: n_s.value].copy()

    def cluster(self, poses):
        arr = (PpfPose * len(poses))(*poses)
        out = (PpfPose * max(len(poses), 1))()
        n = self.o.L.orc_ppf_cluster(self.h, arr, len(poses), out, len(poses))
        return list(out[:n])


class OracleIcp:
    def __init__(self, orc: Oracle, target: np.ndarray, normals, wide_accum: bool):
        self.o = orc
        self.t = _as_f32(target)
        self.n = _as_f32(normals) if normals is not None else None
        self.h = orc.L.orc_icp_create()
        orc.L.orc_icp_set_wide_accum(self.h, int(wide_accum))
        orc.L.orc_icp_set_target(self.h, self.t.ctypes.data, self.t.shape[0], self.t.strides[0],
                                 self.n.ctypes.data if self.n is not None else None,
                                 self.n.strides[0] if self.n is not None else 0)

    def __del__(self):
        try:
            self.o.L.orc_icp_destroy(self.h)
        except Exception:
            pass

    def align(self, source: np.ndarray, guess=None, params: IcpParams | None = None, trace_cap: int = 0):
        s = _as_f32(source)
        prm = params or default_params()
        g = np.ascontiguousarray(np.asarray(guess, np.float32).T) if guess is not None else None  # -> column-major
        res = IcpResult()
        aligned = np.empty((s.shape[0], 4), np.float32)
        idx = np.empty(s.shape[0], np.int32)
        d2 = np.empty(s.shape[0], np.float32)
        tT = np.zeros((max(trace_cap, 1), 16), np.float32)
        tm = np.zeros(max(trace_cap, 1), np.float64)
        tn = C.c_size_t(0)
        self.o.L.orc_icp_align(self.h, s.ctypes.data, s.shape[0], s.strides[0], g.ctypes.data if g is not None else None,
                               C.byref(prm), C.byref(res), aligned.ctypes.data, idx.ctypes.data, d2.ctypes.data,
                               tT.ctypes.data if trace_cap else None, tm.ctypes.data if trace_cap else None,
                               trace_cap, C.byref(tn))
        out = {"result": res, "aligned": aligned, "corr_idx": idx, "corr_d2": d2}
        if trace_cap:
            k = tn.value
            out["trace_T"] = tT[:k].reshape(k, 4, 4).transpose(0, 2, 1).copy()
            out["trace_mse"] = tm[:k].copy()
        return out

    def align_batch(self, source: np.ndarray, guesses: np.ndarray, params: IcpParams | None = None, threads: int = 0):
        s = _as_f32(source)
        prm = params or default_params()
        g = np.ascontiguousarray(np.asarray(guesses, np.float32).transpose(0, 2, 1)).reshape(-1, 16)
        H = g.shape[0]
        res = (IcpResult * H)()
        self.o.L.orc_icp_align_batch(self.h, s.ctypes.data, s.shape[0], s.strides[0], g.ctypes.data, H, C.byref(prm),
                                     res, threads)
        return list(res)

    def fitness(self, source: np.ndarray, T: np.ndarray, max_range: float = DBL_MAX):
        s = _as_f32(source)
        t = np.ascontiguousarray(np.asarray(T, np.float32).T)
        ni = C.c_int32(0)
        f = self.o.L.orc_icp_fitness(self.h, s.ctypes.data, s.shape[0], s.strides[0], t.ctypes.data, max_range,
                                     C.byref(ni))
        return float(f), int(ni.value)
