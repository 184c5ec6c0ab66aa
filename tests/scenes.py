"""Synthetic scenes + camera helpers shared by the tests, smoke() and bench.py (numpy only).

Scenes follow SURVEY.md 8(d): analytic depth of a sphere resting on a ground plane ("S-sphere") and a
table with cubes ("S-table") rendered per pixel by exact ray casting, pinhole camera with hfov 90 deg
(fx = fy = W/2, cx = W/2, cy = H/2 -- nvblox_torch/tests/helpers/camera_utils.py:15-35).
"""
import numpy as np


def intrinsics(width: int, height: int, hfov_deg: float = 90.0) -> np.ndarray:
    fx = (width / 2.0) / np.tan(np.deg2rad(hfov_deg) / 2.0)
    K = np.eye(3, dtype=np.float32)
    K[0, 0] = K[1, 1] = fx
    K[0, 2] = width / 2.0
    K[1, 2] = height / 2.0
    return K


def look_at(eye, target, up=(0.0, 0.0, 1.0)) -> np.ndarray:
    """T_W_C (4x4 float32) of a camera at `eye` looking at `target` (camera z forward, x right, y down)."""
    eye = np.asarray(eye, np.float64)
    z = np.asarray(target, np.float64) - eye
    z /= np.linalg.norm(z)
    up = np.asarray(up, np.float64)
    if abs(np.dot(z, up)) > 0.999:
        up = np.array([1.0, 0.0, 0.0])
    x = np.cross(z, up)
    x /= np.linalg.norm(x)
    y = np.cross(z, x)
    T = np.eye(4)
    T[:3, 0], T[:3, 1], T[:3, 2], T[:3, 3] = x, y, z, eye
    return T.astype(np.float32)


def orbit_pose(i: int, n: int = 64, radius: float = 0.45, height: float = 0.5, center=(0.35, 0.0, 0.1)):
    a = 2.0 * np.pi * i / n
    eye = (center[0] + radius * np.cos(a), center[1] + radius * np.sin(a), height)
    return look_at(eye, center)


def _rays(K, H, W, T):
    u = (np.arange(W, dtype=np.float64) + 0.5 - K[0, 2]) / K[0, 0]
    v = (np.arange(H, dtype=np.float64) + 0.5 - K[1, 2]) / K[1, 1]
    uu, vv = np.meshgrid(u, v)
    d_c = np.stack([uu, vv, np.ones_like(uu)], -1)  # z = 1: depth == ray parameter
    R = T[:3, :3].astype(np.float64)
    return d_c @ R.T, T[:3, 3].astype(np.float64)


def render_depth(K, H, W, T, spheres=(), boxes=(), plane_z=None, max_depth=20.0) -> np.ndarray:
    """Exact depth (z in camera frame) image of spheres [(cx,cy,cz,r)], boxes [(min3, max3)] and z=plane_z."""
    d, o = _rays(K, H, W, T)
    best = np.full((H, W), np.inf)
    if plane_z is not None:
        with np.errstate(divide='ignore', invalid='ignore'):
            t = (plane_z - o[2]) / d[..., 2]
        t[~(t > 1e-6)] = np.inf
        best = np.minimum(best, t)
    for (cx, cy, cz, r) in spheres:
        oc = o - np.array([cx, cy, cz])
        a = (d * d).sum(-1)
        b = 2.0 * (d * oc).sum(-1)
        c = (oc * oc).sum() - r * r
        disc = b * b - 4 * a * c
        with np.errstate(invalid='ignore'):
            t = (-b - np.sqrt(disc)) / (2 * a)
        t[~(disc >= 0)] = np.inf
        t[~(t > 1e-6)] = np.inf
        best = np.minimum(best, t)
    for (mn, mx) in boxes:
        mn = np.asarray(mn, np.float64)
        mx = np.asarray(mx, np.float64)
        with np.errstate(divide='ignore', invalid='ignore'):
            t1 = (mn - o) / d
            t2 = (mx - o) / d
        tmin = np.nanmax(np.minimum(t1, t2), -1)
        tmax = np.nanmin(np.maximum(t1, t2), -1)
        t = np.where((tmax >= tmin) & (tmin > 1e-6), tmin, np.inf)
        best = np.minimum(best, t)
    best[~np.isfinite(best)] = 0.0
    best[best > max_depth] = 0.0
    return best.astype(np.float32)


# S-sphere: sphere r=0.4 at (0.35, 0, 0.25) + ground plane z=0 (scaled-down test_feature_integrator fixture)
S_SPHERE = dict(spheres=[(0.35, 0.0, 0.25, 0.4)], plane_z=0.0)

# smaller sphere that leaves room for an orbiting wrist camera inside the cube-stacking workspace
S_SPHERE_SMALL = dict(spheres=[(0.35, 0.0, 0.12, 0.12)], plane_z=0.0)

# S-table: plane z=0 + six 6 cm cubes inside the cube-stacking workspace box
S_TABLE = dict(plane_z=0.0,
               boxes=[((0.30 + 0.09 * i, -0.2 + 0.08 * i, 0.0), (0.36 + 0.09 * i, -0.14 + 0.08 * i, 0.06))
                      for i in range(6)])

# Workspace boxes, mindmap/mapping/nvblox_mapper_constants.py:54-71
WS_CUBE_STACKING = ((-0.25, -0.65, -0.07), (1.0, 0.62, 0.56))
WS_DRILL_IN_BOX = ((-0.37, -0.75, -0.13), (0.95, 0.75, 0.65))


def feature_frame(H: int, W: int, C: int, seed: int) -> np.ndarray:
    """N(0,1) fp16 feature image, regenerated per frame from its seed."""
    rng = np.random.default_rng(seed)
    return rng.standard_normal((H, W, C), dtype=np.float32).astype(np.float16)


def color_frame(H: int, W: int, seed: int) -> np.ndarray:
    """Uniform random uint8 RGB image (every pixel differs from its neighbours, so the bilinear rounding and the
    blend rounding of the colour path are exercised on every voxel)."""
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8)


def border_lower_half_mask(H: int, W: int, border_percent: int = 5) -> np.ndarray:
    """Mask variant of SURVEY 8(d): 5 % border and the lower half zero."""
    m = np.ones((H, W), np.uint8)
    b = int(round(H * border_percent / 100.0))
    m[:b] = 0
    m[-b:] = 0
    m[:, :b] = 0
    m[:, -b:] = 0
    m[H // 2:] = 0
    return m
