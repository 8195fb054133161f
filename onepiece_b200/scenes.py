"""Deterministic synthetic RGB-D inputs for the parity tests and bench.py.

No dataset ships with the reference (its TestData is an external download, README.md:13) and there is no
network, so every input of the hot path is generated here from closed-form scenes.  Nothing in this file
computes any part of the hot path itself.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np


@dataclass(frozen=True)
class Camera:
    """Pinhole intrinsics; defaults are the reference's OPEN3D_DATASET camera (src/Camera/Camera.h:94-105)."""

    fx: float = 514.817
    fy: float = 515.375
    cx: float = 318.771
    cy: float = 238.447
    width: int = 640
    height: int = 480
    depth_scale: float = 1000.0

    def scaled(self, s: int) -> "Camera":
        return Camera(self.fx * s, self.fy * s, self.cx * s, self.cy * s, self.width * s, self.height * s,
                      self.depth_scale)

    def next_pyramid(self) -> "Camera":
        # PinholeCamera::GenerateNextPyramid (Camera.h:38-42) in float32 arithmetic
        f = np.float32
        return Camera(float(f(self.fx) / f(2)), float(f(self.fy) / f(2)), float(f(self.cx) / f(2)),
                      float(f(self.cy) / f(2)), self.width // 2, self.height // 2, self.depth_scale)


def wavy_wall(cam: Camera = Camera(), k: int = 0, noise: bool = True):
    """Scene S1: depth(v,u) = 2.0 + 0.3 sin(0.02 u) cos(0.015 v) metres (float32), bgr = (u, v, u+v) & 255.

    Frame k >= 1 adds N(0, 1 mm) noise and zeroes 2 % of the pixels (exercises the d <= 0 path)."""
    u = np.arange(cam.width, dtype=np.float32)[None, :]
    v = np.arange(cam.height, dtype=np.float32)[:, None]
    s = 640.0 / cam.width  # keep the same surface when the image is up-scaled
    depth = (np.float32(2.0) + np.float32(0.3) * np.sin(np.float32(0.02 * s) * u) *
             np.cos(np.float32(0.015 * s) * v)).astype(np.float32)
    if k >= 1 and noise:
        rng = np.random.default_rng(k)
        depth = (depth + rng.normal(0.0, 1e-3, depth.shape).astype(np.float32)).astype(np.float32)
        drop = np.random.default_rng(1000 + k).random(depth.shape) < 0.02
        depth[drop] = 0.0
    ui = np.arange(cam.width, dtype=np.int64)[None, :]
    vi = np.arange(cam.height, dtype=np.int64)[:, None]
    bgr = np.stack([(ui + 0 * vi) & 255, (vi + 0 * ui) & 255, (ui + vi) & 255], axis=-1).astype(np.uint8)
    return np.ascontiguousarray(depth), np.ascontiguousarray(bgr)


def se3_exp(xi) -> np.ndarray:
    """Closed-form SE(3) exponential, tangent order (translation, rotation) as Sophus uses (float64)."""
    xi = np.asarray(xi, dtype=np.float64)
    ups, om = xi[:3], xi[3:]
    th = float(np.linalg.norm(om))
    W = np.array([[0, -om[2], om[1]], [om[2], 0, -om[0]], [-om[1], om[0], 0]], dtype=np.float64)
    if th < 1e-10:
        R = np.eye(3) + W
        V = np.eye(3) + 0.5 * W
    else:
        R = np.eye(3) + math.sin(th) / th * W + (1 - math.cos(th)) / th**2 * (W @ W)
        V = np.eye(3) + (1 - math.cos(th)) / th**2 * W + (th - math.sin(th)) / th**3 * (W @ W)
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = V @ ups
    return T


S2_XI = np.array([2e-3, 1e-3, 0.5e-3, 1.0e-3, 1.5e-3, 0.5e-3])


def room_pose(k: int) -> np.ndarray:
    """Scene S2 camera-to-world pose of frame k: exp(k * xi)."""
    return se3_exp(k * S2_XI)


def _texture(p):
    return 128.0 + 100.0 * np.sin(7.0 * p[..., 0]) * np.sin(9.0 * p[..., 1] + 3.0 * p[..., 2])


def room(cam: Camera = Camera(), k: int = 0, depth_u16: bool = True, with_normals: bool = False, pose=None):
    """Scene S2: analytic ray-cast of the inside of an asymmetric box room, x in [-1.4, 1.8], y in [-1.1, 1.3],
    z in [-1, 3] m, plus a sphere r=0.45 m at (0.4,-0.3,2.2), seen from room_pose(k) (or `pose`).  (SURVEY.md's
    symmetric 4 m box with an on-axis sphere shows the camera only the far wall and the sphere, which leaves the
    rotation about the optical axis unconstrained; side walls, floor and ceiling are visible here, so the 6x6
    systems of ICP and dense odometry are well conditioned.)  Returns (depth, bgr, pose_c2w float32[, normals]).  Depth is
    quantised to uint16 millimetres like a sensor unless depth_u16=False (then float32 metres).  normals
    (H x W x 3 float32, camera frame) are the analytic surface normals at every pixel."""
    T = room_pose(k) if pose is None else np.asarray(pose, np.float64)
    R, t = T[:3, :3], T[:3, 3]
    u = np.arange(cam.width, dtype=np.float64)[None, :]
    v = np.arange(cam.height, dtype=np.float64)[:, None]
    d_cam = np.stack([(u - cam.cx) / cam.fx + 0 * v, (v - cam.cy) / cam.fy + 0 * u, np.ones((cam.height, cam.width))], -1)
    d = d_cam @ R.T  # world ray directions (not normalised: parameter == camera z)
    o = t
    lo = np.array([-1.4, -1.1, -1.0])
    hi = np.array([1.8, 1.3, 3.0])
    with np.errstate(divide="ignore", invalid="ignore"):
        t1 = (lo - o) / d
        t2 = (hi - o) / d
    t_exit = np.min(np.maximum(t1, t2), axis=-1)  # inside the box: first wall hit
    # sphere
    c = np.array([0.4, -0.3, 2.2])
    rad = 0.45
    oc = o - c
    a = np.sum(d * d, -1)
    b = 2.0 * np.sum(d * oc, -1)
    cc = float(oc @ oc) - rad * rad
    disc = b * b - 4 * a * cc
    with np.errstate(invalid="ignore"):
        ts = np.where(disc > 0, (-b - np.sqrt(np.maximum(disc, 0))) / (2 * a), np.inf)
    ts = np.where(ts > 0, ts, np.inf)
    z = np.minimum(t_exit, ts)
    p = o + d * z[..., None]
    inten = np.clip(_texture(p), 0, 255)
    bgr = np.stack([inten, inten, inten], -1).astype(np.uint8)
    if depth_u16:
        depth = np.clip(np.rint(z * cam.depth_scale), 0, 65535).astype(np.uint16)
    else:
        depth = z.astype(np.float32)
    if not with_normals:
        return np.ascontiguousarray(depth), np.ascontiguousarray(bgr), T.astype(np.float32)
    # world-space normals: the wall that was hit (axis of the binding slab), or the sphere's radial direction
    tm = np.maximum(t1, t2)
    axis = np.argmin(tm, axis=-1)
    n_w = np.zeros_like(p)
    sign = -np.sign(np.take_along_axis(d, axis[..., None], -1))[..., 0]
    np.put_along_axis(n_w, axis[..., None], sign[..., None], -1)
    on_sphere = ts < t_exit
    n_w[on_sphere] = (p[on_sphere] - c) / rad
    n_c = n_w @ R  # world -> camera: R^T n
    return (np.ascontiguousarray(depth), np.ascontiguousarray(bgr), T.astype(np.float32),
            np.ascontiguousarray(n_c.astype(np.float32)))


def backproject(depth, cam: Camera):
    """Organised point cloud from a depth map in float32, raster order, z > 0 only -- the same arithmetic as
    PointCloud::LoadFromDepth (src/Geometry/PointCloud.cpp:72-100): x = (j - cx) * z / fx."""
    f = np.float32
    if depth.dtype == np.uint16:
        z = depth.astype(np.float32) / f(cam.depth_scale)
    else:
        z = depth.astype(np.float32)
    j = np.arange(cam.width, dtype=np.float32)[None, :]
    i = np.arange(cam.height, dtype=np.float32)[:, None]
    x = ((j - f(cam.cx)) * z) / f(cam.fx)
    y = ((i - f(cam.cy)) * z) / f(cam.fy)
    m = z > 0
    return np.ascontiguousarray(np.stack([x[m], y[m], z[m]], -1).astype(np.float32))
