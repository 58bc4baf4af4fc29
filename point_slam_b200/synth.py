"""Deterministic synthetic scenes / frames / neural point clouds (numpy only).

No datasets are reachable from the build or GPU boxes, so every test and
`bench.py` draws its inputs from here: an analytic
6 x 4 x 2.7 m room with a few boxes, exact ray-cast depth, a smooth procedural
colour field, Point-SLAM-like surface-hugging point clouds (3 points per
surface location at {0.98, 1.00, 1.02} x depth along a viewing ray, the adding
rule of the reference `neural_point.py:126-145`) and float64 dynamic-radius
maps built like the reference callers build them (`Tracker.py:235-248`).

Seed convention: 1219 is the reference's `setup_seed` (configs/point_slam.yaml:6).
"""
from __future__ import annotations

import numpy as np

ROOM = np.array([[0.0, 0.0, 0.0], [6.0, 4.0, 2.7]])
BOXES = np.array([                       # (lo, hi) axis-aligned furniture
    [[1.0, 0.8, 0.0], [2.2, 1.6, 0.75]],     # table
    [[3.6, 2.4, 0.0], [4.4, 3.6, 1.8]],      # cabinet
    [[4.9, 0.3, 0.0], [5.7, 1.1, 0.45]],     # stool
    [[0.2, 2.9, 0.0], [1.4, 3.8, 0.9]],      # sofa
])

TUM_INTRINSICS = dict(H=480, W=640, fx=517.3, fy=516.5, cx=318.6, cy=255.3)   # configs/TUM_RGBD/tum.yaml:22-27


def colour_field(x: np.ndarray) -> np.ndarray:
    """Smooth RGB in [0,1] as a function of world position (..., 3)."""
    f = np.stack([
        0.5 + 0.35 * np.sin(2.1 * x[..., 0] + 0.7 * x[..., 1]) * np.cos(1.3 * x[..., 2]),
        0.5 + 0.35 * np.sin(1.7 * x[..., 1] - 0.9 * x[..., 2] + 0.5),
        0.5 + 0.35 * np.cos(1.1 * x[..., 0] - 1.9 * x[..., 1] + 0.8 * x[..., 2]),
    ], -1)
    stripes = 0.12 * (np.sin(18.0 * (x[..., 0] + x[..., 1] + x[..., 2])) > 0.6)
    return np.clip(f + stripes[..., None], 0.0, 1.0)


def look_at(eye, target, up=(0.0, 0.0, 1.0)) -> np.ndarray:
    """4x4 camera-to-world; camera looks along -z, +y up (reference convention, datasets.py:142-148)."""
    eye, target, up = (np.asarray(v, dtype=np.float64) for v in (eye, target, up))
    back = eye - target
    back /= np.linalg.norm(back)
    right = np.cross(up, back)
    right /= np.linalg.norm(right)
    upv = np.cross(back, right)
    c2w = np.eye(4)
    c2w[:3, 0], c2w[:3, 1], c2w[:3, 2], c2w[:3, 3] = right, upv, back, eye
    return c2w


def trajectory(n_frames: int, seed: int = 1219) -> np.ndarray:
    """Smooth closed camera path inside the room, (n,4,4) float64."""
    rng = np.random.default_rng(seed)
    ph = rng.uniform(0, 2 * np.pi)
    poses = []
    for k in range(n_frames):
        a = ph + 2 * np.pi * k / max(n_frames, 1) * 0.35
        eye = np.array([3.0 + 1.1 * np.cos(a), 2.0 + 0.7 * np.sin(a), 1.35 + 0.15 * np.sin(2 * a)])
        tgt = np.array([3.0 + 2.6 * np.cos(a + 2.4), 2.0 + 1.7 * np.sin(a + 2.4), 1.0 + 0.3 * np.cos(a)])
        poses.append(look_at(eye, tgt))
    return np.stack(poses)


def _slab(o, d, lo, hi):
    with np.errstate(divide='ignore', invalid='ignore'):
        inv = 1.0 / d
        t0 = (lo - o) * inv
        t1 = (hi - o) * inv
    tn = np.nanmax(np.minimum(t0, t1), -1)
    tf = np.nanmin(np.maximum(t0, t1), -1)
    return tn, tf


def ray_cast(o: np.ndarray, d: np.ndarray) -> np.ndarray:
    """First-hit ray parameter t for rays o + t d (d un-normalised) against room interior + boxes."""
    o = np.broadcast_to(np.asarray(o, np.float64), d.shape)
    d = np.asarray(d, np.float64)
    _, t_room = _slab(o, d, ROOM[0], ROOM[1])
    t = t_room.copy()
    for lo, hi in BOXES:
        tn, tf = _slab(o, d, lo, hi)
        hit = (tn <= tf) & (tn > 1e-6)
        t = np.where(hit & (tn < t), tn, t)
    return t


def pixel_rays(c2w: np.ndarray, H, W, fx, fy, cx, cy):
    """Un-normalised world rays of every pixel (camera z component -1 => ray parameter == z-depth)."""
    j, i = np.meshgrid(np.arange(H, dtype=np.float64), np.arange(W, dtype=np.float64), indexing='ij')
    cam = np.stack([(i - cx) / fx, -(j - cy) / fy, -np.ones_like(i)], -1)
    d = cam @ c2w[:3, :3].T
    return c2w[:3, 3], d


def make_frame(c2w: np.ndarray, intr=TUM_INTRINSICS, noise: bool = False, holes: float = 0.0, seed: int = 0):
    """-> gt_depth (H,W) f32 (z-depth, metres), gt_color (H,W,3) f32 in [0,1].
    noise=True adds Kinect-style sigma = 0.0012 + 0.0019 (z-0.4)^2; holes = fraction of zero-depth pixels."""
    H, W = intr['H'], intr['W']
    o, d = pixel_rays(c2w, H, W, intr['fx'], intr['fy'], intr['cx'], intr['cy'])
    t = ray_cast(o, d)
    pts = o + d * t[..., None]
    col = colour_field(pts)
    rng = np.random.default_rng(1219 + seed)
    if noise:
        t = t + rng.standard_normal(t.shape) * (0.0012 + 0.0019 * (t - 0.4) ** 2)
    if holes > 0:
        t = np.where(rng.uniform(size=t.shape) < holes, 0.0, t)
    return t.astype(np.float32), col.astype(np.float32)


def _surface_faces():
    faces = []          # (origin, edge_u, edge_v, area)
    def add_box(lo, hi, skip_bottom):
        ext = hi - lo
        for ax in range(3):
            u, v = [(1, 2), (0, 2), (0, 1)][ax]
            eu, ev = np.zeros(3), np.zeros(3)
            eu[u], ev[v] = ext[u], ext[v]
            for side in (0, 1):
                if skip_bottom and ax == 2 and side == 0:
                    continue
                org = lo.copy()
                org[ax] = hi[ax] if side else lo[ax]
                faces.append((org, eu, ev, ext[u] * ext[v]))
    add_box(ROOM[0], ROOM[1], False)
    for lo, hi in BOXES:
        add_box(lo, hi, True)
    return faces


def make_cloud(n_points: int, seed: int = 1219, n_add: int = 3, jitter: float = 0.0) -> np.ndarray:
    """Surface-hugging neural point cloud, (n_points,3) float32, stored location-major
    (the n_add points of one location are consecutive, neural_point.py:142-145)."""
    rng = np.random.default_rng(seed)
    faces = _surface_faces()
    area = np.array([f[3] for f in faces])
    n_loc = (n_points + n_add - 1) // n_add
    fid = rng.choice(len(faces), size=n_loc, p=area / area.sum())
    uv = rng.uniform(size=(n_loc, 2))
    org = np.stack([faces[k][0] for k in range(len(faces))])[fid]
    eu = np.stack([faces[k][1] for k in range(len(faces))])[fid]
    ev = np.stack([faces[k][2] for k in range(len(faces))])[fid]
    loc = org + eu * uv[:, :1] + ev * uv[:, 1:]
    if jitter > 0:
        loc = loc + rng.standard_normal(loc.shape) * jitter
    eye = np.array([3.0, 2.0, 1.35]) + rng.uniform(-1.0, 1.0, size=(n_loc, 3)) * np.array([1.2, 0.8, 0.4])
    ray = loc - eye
    scale = np.linspace(0.98, 1.02, n_add)                         # near/far_end_surface
    pts = eye[:, None, :] + ray[:, None, :] * scale[None, :, None]
    return pts.reshape(-1, 3)[:n_points].astype(np.float32)


def sobel_radius_map(gt_color: np.ndarray, r_add_max=0.08, r_add_min=0.02, ratio=2.0, thresh=0.15):
    """Per-pixel float64 (r_add, r_query) maps the way the reference callers derive them
    (Tracker.py:235-248): grey = 0.2125 R + 0.7154 G + 0.0721 B, 3x3 Sobel (/4 smoothing,
    zero 1-px border), piece-wise linear map [0,0.01,thresh] -> [r_max,r_max,r_min]."""
    g = gt_color.astype(np.float64) @ np.array([0.2125, 0.7154, 0.0721])
    gy = np.zeros_like(g)
    gx = np.zeros_like(g)
    gy[1:-1, 1:-1] = ((g[2:, :-2] + 2 * g[2:, 1:-1] + g[2:, 2:]) - (g[:-2, :-2] + 2 * g[:-2, 1:-1] + g[:-2, 2:])) / 4.0
    gx[1:-1, 1:-1] = ((g[:-2, 2:] + 2 * g[1:-1, 2:] + g[2:, 2:]) - (g[:-2, :-2] + 2 * g[1:-1, :-2] + g[2:, :-2])) / 4.0
    mag = np.clip(np.sqrt(gx ** 2 + gy ** 2), 0.0, thresh)
    r_add = np.interp(mag, [0.0, 0.01, thresh], [r_add_max, r_add_max, r_add_min])
    return r_add, ratio * r_add


def make_features(n_points: int, c_dim: int = 32, seed: int = 1219):
    """geo/col features ~ N(0, 0.1^2) float32 (neural_point.py:151-159), numpy RNG so the
    same values appear on every box."""
    rng = np.random.default_rng(seed + 7)
    return ((rng.standard_normal((n_points, c_dim)) * 0.1).astype(np.float32),
            (rng.standard_normal((n_points, c_dim)) * 0.1).astype(np.float32))
