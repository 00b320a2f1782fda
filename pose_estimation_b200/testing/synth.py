"""Deterministic synthetic clouds for the parity tests and bench.py.

Units are metres, the camera sits at the origin looking along +z with a working distance of
about 0.7 m (the rig of the reference: hand-eye translation (0.190, 0.064, 0.688) m,
pose_estimation_manager/src/pose_transformer.cpp:10-12).  Everything is seeded
(np.random.default_rng(seed), PCG64) and float32 on output, so the oracle and the CUDA path
read identical bytes.  No file of /root/reference is read.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

FOOT_X = 0.12  # object footprint [m]
FOOT_Y = 0.08
RELIEF = 0.03


# ------------------------------------------------------------------------------------------
# rigid-motion helpers (float64 on purpose: these only *generate* inputs)
# ------------------------------------------------------------------------------------------
def rotation_about(axis, angle_rad: float) -> np.ndarray:
    a = np.asarray(axis, np.float64)
    a = a / np.linalg.norm(a)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    return np.eye(3) + np.sin(angle_rad) * K + (1 - np.cos(angle_rad)) * (K @ K)


def make_pose(R: np.ndarray, t) -> np.ndarray:
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = np.asarray(t, np.float64)
    return T


def apply_pose(T: np.ndarray, pts: np.ndarray) -> np.ndarray:
    return pts @ T[:3, :3].T + T[:3, 3]


def random_unit(rng) -> np.ndarray:
    v = rng.normal(size=3)
    return v / np.linalg.norm(v)


def perturb_pose(T: np.ndarray, rng, max_angle_deg: float, max_trans: float, centre=None,
                 exact: bool = False) -> np.ndarray:
    """T' = P o T with P a rotation (about `centre`, default T's translation) of angle U[0,max]
    (or exactly max if `exact`) about a random axis and a translation uniform in a ball (or on its
    surface if `exact`)."""
    ang = np.deg2rad(max_angle_deg) * (1.0 if exact else rng.uniform())
    R = rotation_about(random_unit(rng), ang)
    d = random_unit(rng) * max_trans * (1.0 if exact else rng.uniform() ** (1 / 3))
    c = T[:3, 3] if centre is None else np.asarray(centre, np.float64)
    P = np.eye(4)
    P[:3, :3] = R
    P[:3, 3] = c - R @ c + d
    return P @ T


def pose_error(A: np.ndarray, B: np.ndarray):
    """(rotation angle [rad], translation distance [m]) between two 4x4 poses."""
    A = np.asarray(A, np.float64)
    B = np.asarray(B, np.float64)
    R = A[:3, :3] @ B[:3, :3].T
    # atan2 form stays accurate near 0 (acos of the trace does not)
    w = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    ang = np.arctan2(0.5 * np.linalg.norm(w), 0.5 * (np.trace(R) - 1.0))
    return float(abs(ang)), float(np.linalg.norm(A[:3, 3] - B[:3, 3]))


# ------------------------------------------------------------------------------------------
# the object surface S: a non-symmetric bumpy height field z = f(x, y)
# ------------------------------------------------------------------------------------------
@dataclass
class Surface:
    seed: int = 7
    centres: np.ndarray = field(init=False)
    sigmas: np.ndarray = field(init=False)
    amps: np.ndarray = field(init=False)

    def __post_init__(self):
        rng = np.random.default_rng(1000 + self.seed)
        k = 7
        self.centres = np.stack([rng.uniform(-0.5 * FOOT_X, 0.5 * FOOT_X, k), rng.uniform(-0.5 * FOOT_Y, 0.5 * FOOT_Y, k)], 1)
        self.sigmas = rng.uniform(0.010, 0.028, k)
        amps = rng.uniform(0.4, 1.0, k) * rng.choice([-1.0, 1.0], k, p=[0.3, 0.7])
        self.amps = amps
        # normalise the relief to RELIEF peak-to-peak on a probe lattice
        gx, gy = np.meshgrid(np.linspace(-0.5 * FOOT_X, 0.5 * FOOT_X, 121), np.linspace(-0.5 * FOOT_Y, 0.5 * FOOT_Y, 81))
        z = self._raw(gx.ravel(), gy.ravel())
        self.amps = amps * (RELIEF / (z.max() - z.min()))

    def _raw(self, x, y):
        z = np.zeros_like(x, dtype=np.float64)
        for (cx, cy), s, a in zip(self.centres, self.sigmas, self.amps):
            z += a * np.exp(-((x - cx) ** 2 + (y - cy) ** 2) / (2 * s * s))
        return z

    def height(self, x, y):
        return self._raw(np.asarray(x, np.float64), np.asarray(y, np.float64))

    def gradient(self, x, y):
        x = np.asarray(x, np.float64)
        y = np.asarray(y, np.float64)
        gx = np.zeros_like(x)
        gy = np.zeros_like(x)
        for (cx, cy), s, a in zip(self.centres, self.sigmas, self.amps):
            e = a * np.exp(-((x - cx) ** 2 + (y - cy) ** 2) / (2 * s * s))
            gx += e * (-(x - cx) / (s * s))
            gy += e * (-(y - cy) / (s * s))
        return gx, gy

    def sample(self, n: int, rng):
        """n uniform-area samples of S in the object frame -> (points (n,3), unit normals (n,3)), float64."""
        pts = np.empty((0, 3))
        gmax = 2.5  # bound of sqrt(1+|grad|^2) for these parameters (checked below)
        while pts.shape[0] < n:
            m = int((n - pts.shape[0]) * 1.6) + 64
            x = rng.uniform(-0.5 * FOOT_X, 0.5 * FOOT_X, m)
            y = rng.uniform(-0.5 * FOOT_Y, 0.5 * FOOT_Y, m)
            gx, gy = self.gradient(x, y)
            w = np.sqrt(1 + gx * gx + gy * gy)
            assert w.max() < gmax * 4
            keep = rng.uniform(0, max(gmax, w.max()), m) < w
            p = np.stack([x[keep], y[keep], self.height(x[keep], y[keep])], 1)
            pts = np.concatenate([pts, p], 0)
        pts = pts[:n]
        gx, gy = self.gradient(pts[:, 0], pts[:, 1])
        nrm = np.stack([-gx, -gy, np.ones_like(gx)], 1)
        nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
        return pts, nrm


def default_gt_pose(rng) -> np.ndarray:
    """Object frame -> camera frame: roughly fronto-parallel (the surface's +z faces the camera), ~0.70 m away."""
    R = rotation_about([1, 0, 0], np.pi) @ rotation_about(random_unit(rng), np.deg2rad(rng.uniform(5, 12)))
    t = np.array([rng.uniform(-0.03, 0.03), rng.uniform(-0.02, 0.02), 0.70])
    return make_pose(R, t)


# ------------------------------------------------------------------------------------------
# organized scene rendering (pinhole, Zivid-like 1944 x 1200)
# ------------------------------------------------------------------------------------------
def render_scene(surface: Surface, gt_pose: np.ndarray, rng, width: int = 1944, height: int = 1200,
                 noise_sigma: float = 1e-4, nan_fraction: float = 0.04) -> np.ndarray:
    """Organized cloud, (height*width, 4) float32 rows (x, y, z, 1), NaN xyz for missing pixels."""
    scale = width / 1944.0
    fx = fy = 1944.0 * scale
    cx, cy = 972.0 * scale, 600.0 * scale
    u, v = np.meshgrid(np.arange(width, dtype=np.float64), np.arange(height, dtype=np.float64))
    dx = ((u - cx) / fx).ravel()
    dy = ((v - cy) / fy).ravel()
    npx = dx.shape[0]
    depth = np.full(npx, np.inf)

    # background plane through (0,0,0.75), tilted 5 degrees
    n = rotation_about([1.0, 0.3, 0.0], np.deg2rad(5.0)) @ np.array([0.0, 0.0, 1.0])
    c = n @ np.array([0.0, 0.0, 0.75])
    denom = n[0] * dx + n[1] * dy + n[2]
    t_plane = c / denom
    depth = np.where((t_plane > 0), np.minimum(depth, t_plane), depth)

    # three clutter ellipsoids resting in front of the plane
    for _ in range(3):
        while True:  # keep the clutter off the object
            ctr = np.array([rng.uniform(-0.25, 0.25), rng.uniform(-0.15, 0.15), rng.uniform(0.66, 0.72)])
            if abs(ctr[0] - gt_pose[0, 3]) > 0.13 or abs(ctr[1] - gt_pose[1, 3]) > 0.10:
                break
        rad = rng.uniform(0.02, 0.05, 3)
        # |(t d - ctr)/rad|^2 = 1
        ax = dx / rad[0]
        ay = dy / rad[1]
        az = 1.0 / rad[2]
        bx, by, bz = ctr / rad
        A = ax * ax + ay * ay + az * az
        B = -2 * (ax * bx + ay * by + az * bz)
        Cc = bx * bx + by * by + bz * bz - 1
        disc = B * B - 4 * A * Cc
        hit = disc > 0
        t_e = np.where(hit, (-B - np.sqrt(np.where(hit, disc, 0))) / (2 * A), np.inf)
        depth = np.minimum(depth, np.where(t_e > 0, t_e, np.inf))

    # the object: ray / height-field intersection by fixed-point iteration in the object frame
    R = gt_pose[:3, :3]
    t = gt_pose[:3, 3]
    o = -R.T @ t
    d = np.stack([dx, dy, np.ones_like(dx)], 1) @ R  # = (R^T d) per row
    # candidate rays: those crossing the footprint slab near z_obj = 0
    s0 = -o[2] / d[:, 2]
    x0 = o[0] + s0 * d[:, 0]
    y0 = o[1] + s0 * d[:, 1]
    cand = np.flatnonzero((np.abs(x0) < 0.5 * FOOT_X + 0.02) & (np.abs(y0) < 0.5 * FOOT_Y + 0.02) & (s0 > 0))
    s = s0[cand]
    dc = d[cand]
    for _ in range(30):
        x = o[0] + s * dc[:, 0]
        y = o[1] + s * dc[:, 1]
        s = (surface.height(x, y) - o[2]) / dc[:, 2]
    x = o[0] + s * dc[:, 0]
    y = o[1] + s * dc[:, 1]
    resid = np.abs(o[2] + s * dc[:, 2] - surface.height(x, y))
    inside = (np.abs(x) <= 0.5 * FOOT_X) & (np.abs(y) <= 0.5 * FOOT_Y) & (resid < 1e-7) & (s > 0)
    obj_depth = np.full(npx, np.inf)
    obj_depth[cand[inside]] = s[inside]  # ray parameter == camera z because d_cam.z == 1
    depth = np.minimum(depth, obj_depth)

    # sensor noise along the ray (also breaks the lattice ties of a perfectly regular image)
    depth = depth + rng.normal(0.0, noise_sigma, npx) / np.sqrt(dx * dx + dy * dy + 1.0)
    missing = ~np.isfinite(depth) | (rng.uniform(size=npx) < nan_fraction)
    pts = np.stack([dx * depth, dy * depth, depth, np.ones(npx)], 1).astype(np.float32)
    pts[missing, :3] = np.nan
    return pts


# ------------------------------------------------------------------------------------------
# numpy voxel count (only used to pick a leaf size; the product path is CUDA, the parity
# reference is the oracle)
# ------------------------------------------------------------------------------------------
def voxel_count(points: np.ndarray, leaf: float) -> int:
    p = points[:, :3]
    p = p[np.isfinite(p).all(1)]
    inv = np.float32(1.0) / np.float32(leaf)
    ijk = np.floor(p * inv).astype(np.int64)
    ijk -= ijk.min(0)
    dims = ijk.max(0) + 1
    key = ijk[:, 0] + dims[0] * (ijk[:, 1] + dims[1] * ijk[:, 2])
    return int(np.unique(key).shape[0])


def choose_leaf(points: np.ndarray, lo_count: int, hi_count: int, leaf_lo: float = 1e-4, leaf_hi: float = 2e-2) -> float:
    """Bisection on the leaf size until the voxel count lands in [lo_count, hi_count]."""
    for _ in range(60):
        mid = float(np.sqrt(leaf_lo * leaf_hi))
        c = voxel_count(points, mid)
        if c > hi_count:
            leaf_lo = mid
        elif c < lo_count:
            leaf_hi = mid
        else:
            return float(np.float32(mid))
    raise RuntimeError("choose_leaf did not converge")


def xyz4(p: np.ndarray) -> np.ndarray:
    out = np.ones((p.shape[0], 4), np.float32)
    out[:, :3] = p[:, :3]
    return out


# ------------------------------------------------------------------------------------------
# the configurations of BASELINE.json
# ------------------------------------------------------------------------------------------
@dataclass
class Problem:
    name: str
    source: np.ndarray            # (Ns, 4) float32 model, object frame (or C1: perturbed copy)
    target: np.ndarray            # (Nt, 4) float32 scene (already down-sampled unless `organized`)
    guess: np.ndarray             # (4, 4) float64 initial pose(s); (H, 4, 4) for the batch config
    gt_pose: np.ndarray           # (4, 4) float64 pose that maps source onto target
    organized: np.ndarray | None = None   # raw organized scene when the config has one
    leaf: float | None = None
    meta: dict = field(default_factory=dict)


# ------------------------------------------------------------------------------------------
# on-disk form of a problem: raw little-endian float32 N x 4 records (x, y, z, 1.0 — the memory image of
# pcl::PointXYZ) + one JSON sidecar, so that the oracle, the library and the C++ checks read the same bytes
# ------------------------------------------------------------------------------------------
def save_problem(problem: Problem, directory) -> None:
    """source.f32 / target.f32 / organized.f32 (N x 4 float32, C order), guess.f64 (H x 16 or 16 float64, row-major 4 x 4)
    and problem.json (name, counts, organized width x height, ground-truth pose, leaf, generator meta)."""
    import json
    from pathlib import Path

    d = Path(directory)
    d.mkdir(parents=True, exist_ok=True)
    for name in ("source", "target", "organized"):
        a = getattr(problem, name)
        if a is not None:
            np.ascontiguousarray(a, "<f4").tofile(d / f"{name}.f32")
    np.ascontiguousarray(problem.guess, "<f8").tofile(d / "guess.f64")
    side = {
        "name": problem.name, "format": "float32 little-endian, N x 4 (x y z 1.0) per cloud; guess float64 row-major 4 x 4",
        "n_source": int(len(problem.source)), "n_target": int(len(problem.target)),
        "n_organized": None if problem.organized is None else int(len(problem.organized)),
        "n_guesses": int(np.asarray(problem.guess).reshape(-1, 16).shape[0]), "guess_is_batch": bool(np.asarray(problem.guess).ndim == 3),
        "gt_pose": np.asarray(problem.gt_pose, np.float64).tolist(), "leaf": problem.leaf, "meta": problem.meta,
    }
    (d / "problem.json").write_text(json.dumps(side, indent=1))


def load_problem(directory) -> Problem:
    import json
    from pathlib import Path

    d = Path(directory)
    side = json.loads((d / "problem.json").read_text())

    def cloud(name, n):
        if n is None:
            return None
        a = np.fromfile(d / f"{name}.f32", "<f4")
        if a.size != 4 * n:
            raise ValueError(f"{name}.f32 holds {a.size} floats, the sidecar says {n} x 4")
        return a.reshape(n, 4)

    g = np.fromfile(d / "guess.f64", "<f8").reshape(-1, 4, 4)
    return Problem(side["name"], cloud("source", side["n_source"]), cloud("target", side["n_target"]),
                   g if side["guess_is_batch"] else g[0], np.array(side["gt_pose"], np.float64),
                   organized=cloud("organized", side["n_organized"]), leaf=side["leaf"], meta=side["meta"])


def make_c1(n: int = 20000, seed: int = 1) -> Problem:
    """C1: 20k-pt model vs a rigidly moved, 1 mm-noise copy; the CPU-runnable case."""
    rng = np.random.default_rng(seed)
    surf = Surface(seed)
    pts, _ = surf.sample(n, rng)
    pose = default_gt_pose(rng)
    tgt = apply_pose(pose, pts)
    cen = tgt.mean(0)
    R = rotation_about([1, 2, 3], np.deg2rad(3.0))
    P = np.eye(4)
    P[:3, :3] = R
    P[:3, 3] = cen - R @ cen + np.array([0.004, -0.003, 0.005])
    src = apply_pose(P, tgt) + rng.normal(0, 1e-3, tgt.shape)
    return Problem("C1", xyz4(src.astype(np.float32)), xyz4(tgt.astype(np.float32)), np.eye(4), np.linalg.inv(P),
                   meta={"seed": seed, "n": n})


def make_scene_problem(name: str, seed: int, n_model: int, target_lo: int, target_hi: int, width: int = 1944,
                       height: int = 1200, guess_angle_deg: float = 2.0, guess_trans: float = 0.003,
                       n_guesses: int = 0, batch_angle_deg: float = 6.0, batch_trans: float = 0.008,
                       downsample=None) -> Problem:
    """C2/C3/C4/C5: organized scene -> VoxelGrid -> target; model = uniform-area samples of S.

    `downsample(points, leaf) -> (M,4)` performs the voxel filter (tests pass the oracle or the CUDA
    path); when None the raw organized cloud is returned un-filtered in `organized` and `target`
    is left empty for the caller to fill.
    """
    rng = np.random.default_rng(seed)
    surf = Surface(seed)
    gt = default_gt_pose(rng)
    scene = render_scene(surf, gt, rng, width, height)
    leaf = choose_leaf(scene, target_lo, target_hi)
    model, _ = surf.sample(n_model, rng)
    guess = perturb_pose(gt, rng, guess_angle_deg, guess_trans, exact=True)
    if n_guesses:
        guess = np.stack([perturb_pose(gt, rng, batch_angle_deg, batch_trans) for _ in range(n_guesses)], 0)
    tgt = downsample(scene, leaf) if downsample is not None else np.empty((0, 4), np.float32)
    return Problem(name, xyz4(model.astype(np.float32)), tgt, guess, gt, organized=scene, leaf=leaf,
                   meta={"seed": seed, "width": width, "height": height})


def make_c2(scale: float = 1.0, seed: int = 2, downsample=None) -> Problem:
    w, h = int(round(1944 * scale)), int(round(1200 * scale))
    s2 = scale * scale
    return make_scene_problem("C2", seed, max(int(50000 * s2), 500), int(195000 * s2), int(205000 * s2), w, h,
                              downsample=downsample)


def make_c4(scale: float = 1.0, n_guesses: int = 1024, seed: int = 4, downsample=None) -> Problem:
    w, h = int(round(1944 * scale)), int(round(1200 * scale))
    s2 = scale * scale
    return make_scene_problem("C4", seed, max(int(50000 * s2), 500), int(490000 * s2), int(510000 * s2), w, h,
                              n_guesses=n_guesses, downsample=downsample)


# ------------------------------------------------------------------------------------------
# clouds with normals + start poses for the cv::ppf_match_3d::ICP mode (N x 6 float rows)
# ------------------------------------------------------------------------------------------
def make_cvicp_case(seed=0, n_model=4000, n_scene=9000, n_poses=6, noise=1e-4, clutter=0, angle=4.0, trans=0.006):
    rng = np.random.default_rng(seed)
    surf = Surface(3 + seed)
    pts, nrm = surf.sample(n_model, rng)
    gt = default_gt_pose(rng)
    model = np.concatenate([pts, nrm], 1).astype(np.float32)
    spts, snrm = surf.sample(n_scene, rng)
    xyz = apply_pose(gt, spts) + rng.normal(0, noise, spts.shape)
    sn = (gt[:3, :3] @ snrm.T).T
    if clutter:
        cx = rng.uniform([-0.2, -0.2, 0.6], [0.2, 0.2, 0.8], (clutter, 3))
        cn = rng.normal(size=(clutter, 3))
        cn /= np.linalg.norm(cn, axis=1, keepdims=True)
        xyz, sn = np.concatenate([xyz, cx]), np.concatenate([sn, cn])
    scene = np.concatenate([xyz, sn], 1).astype(np.float32)
    poses = np.stack([perturb_pose(gt, rng, angle, trans) for _ in range(n_poses)])
    return model, scene, poses, gt
