"""Synthetic RGB-D depth workloads: input synthesis for tests and bench.py.

* ``boxroom``  -- 640x480, TUM1 intrinsics (/root/reference/Examples/RGB-D/TUM1.yaml:8-11): a 6x3x5 m room
  (floor, ceiling, 4 walls) with a 0.8 m cube and a 1.2x0.4x0.9 m cabinet; camera at the room centre, 1.4 m above
  the floor; pose sequence = yaw 0..360 deg with a +-10 deg pitch sinusoid and a small seeded jitter.
* ``realsense`` -- 1280x720, fx=fy=640, cx=639.5, cy=359.5: the same room plus 40 small tilted patches.

Depth is the exact ray/rectangle z-depth in metres (float32), 0 where a ray hits nothing.  ``add_noise`` applies the
Kinect-like / RealSense-like noise models of the survey with numpy PCG64, seed = 1000 + frame index.
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


@dataclass(frozen=True)
class Intrinsics:
    fx: float
    fy: float
    cx: float
    cy: float
    width: int
    height: int


TUM1 = Intrinsics(517.306408, 516.469215, 318.643040, 255.313989, 640, 480)
REALSENSE = Intrinsics(640.0, 640.0, 639.5, 359.5, 1280, 720)


def _lib():
    path = os.path.join(_HERE, "libspx_scenes.so")
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
    lib = ctypes.CDLL(path)
    lib.spx_scene_render_batch.argtypes = [
        ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
        ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double,
        ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
    lib.spx_scene_render_batch.restype = ctypes.c_int
    return lib


def _box_faces(lo, hi, inward=False):
    """The six faces of an axis-aligned box as (origin, eu, ev) rectangles."""
    lo = np.asarray(lo, float)
    hi = np.asarray(hi, float)
    d = hi - lo
    ex, ey, ez = np.diag(d)
    faces = [
        (lo, ex, ey), (lo + ez, ex, ey),      # z = lo, z = hi
        (lo, ex, ez), (lo + ey, ex, ez),      # y = lo, y = hi
        (lo, ey, ez), (lo + ex, ey, ez),      # x = lo, x = hi
    ]
    return [np.concatenate([o, u, v]) for o, u, v in faces]


def boxroom_rects() -> np.ndarray:
    """World frame: x right, y DOWN, z forward; camera at the origin; floor at y=+1.4, ceiling at y=-1.6."""
    rects = []
    rects += _box_faces([-3.0, -1.6, -2.5], [3.0, 1.4, 2.5])
    # 0.8 m cube standing on the floor, and a 1.2 x 0.9 (high) x 0.4 cabinet against the +z wall
    rects += _box_faces([1.2, 0.6, 0.9], [2.0, 1.4, 1.7])
    rects += _box_faces([-2.2, 0.5, 2.1], [-1.0, 1.4, 2.5])
    return np.ascontiguousarray(np.stack(rects), dtype=np.float64)


def clutter_rects(n: int = 40, seed: int = 4321) -> np.ndarray:
    """``n`` small planar patches (0.15-0.5 m), random tilt, scattered through the room."""
    rng = np.random.Generator(np.random.PCG64(seed))
    rects = []
    for _ in range(n):
        c = np.array([rng.uniform(-2.6, 2.6), rng.uniform(-1.2, 1.2), rng.uniform(-2.1, 2.1)])
        # keep a 0.7 m clear bubble around the camera
        if np.linalg.norm(c) < 0.9:
            c *= 0.9 / max(np.linalg.norm(c), 1e-3)
        su, sv = rng.uniform(0.15, 0.5, size=2)
        a = rng.normal(size=3)
        a /= np.linalg.norm(a)
        b = rng.normal(size=3)
        b -= a * (a @ b)
        b /= np.linalg.norm(b)
        eu, ev = a * su, b * sv
        rects.append(np.concatenate([c - 0.5 * eu - 0.5 * ev, eu, ev]))
    return np.ascontiguousarray(np.stack(rects), dtype=np.float64)


def poses(n_frames: int, total: int = 1000, start: int = 0, seed: int = 1234) -> np.ndarray:
    """(n_frames, 12) camera->world rotation (row major) + camera centre, frames ``start .. start+n_frames``
    of a ``total``-frame orbit: yaw 0..360 deg, pitch +-10 deg sinusoid (3 periods), +-2 cm seeded jitter."""
    rng = np.random.Generator(np.random.PCG64(seed))
    jitter = rng.uniform(-0.02, 0.02, size=(total, 3))
    out = np.zeros((n_frames, 12))
    for k in range(n_frames):
        i = (start + k) % total
        yaw = 2.0 * np.pi * i / total
        pitch = np.deg2rad(10.0) * np.sin(2.0 * np.pi * 3.0 * i / total)
        cy_, sy_ = np.cos(yaw), np.sin(yaw)
        cp, sp = np.cos(pitch), np.sin(pitch)
        r_yaw = np.array([[cy_, 0.0, sy_], [0.0, 1.0, 0.0], [-sy_, 0.0, cy_]])
        r_pitch = np.array([[1.0, 0.0, 0.0], [0.0, cp, -sp], [0.0, sp, cp]])
        out[k, :9] = (r_yaw @ r_pitch).reshape(-1)
        out[k, 9:] = jitter[i]
    return out


def render(rects: np.ndarray, pose12: np.ndarray, intr: Intrinsics, n_threads: int | None = None) -> np.ndarray:
    pose12 = np.ascontiguousarray(np.atleast_2d(pose12), dtype=np.float64)
    rects = np.ascontiguousarray(rects, dtype=np.float64)
    n = pose12.shape[0]
    out = np.empty((n, intr.height, intr.width), dtype=np.float32)
    if n_threads is None:
        n_threads = min(os.cpu_count() or 1, 32, n)
    _lib().spx_scene_render_batch(
        rects.ctypes.data, rects.shape[0], pose12.ctypes.data, n,
        intr.fx, intr.fy, intr.cx, intr.cy, intr.width, intr.height, out.ctypes.data, int(n_threads))
    return out


def boxroom_sequence(n_frames: int, start: int = 0, total: int = 1000) -> np.ndarray:
    """Clean 640x480 box-room frames ``start .. start+n_frames`` (frame 0 is BASELINE.json configs[0])."""
    return render(boxroom_rects(), poses(n_frames, total, start), TUM1)


def realsense_sequence(n_frames: int, start: int = 0, total: int = 1000) -> np.ndarray:
    rects = np.concatenate([boxroom_rects(), clutter_rects()])
    d = render(rects, poses(n_frames, total, start), REALSENSE)
    return np.where((d >= 0.3) & (d <= 6.0), d, 0.0).astype(np.float32)


def _dropout_mask(rng, shape, fraction: float) -> np.ndarray:
    """Zero-depth holes covering ~``fraction`` of the image as random discs (radius 2-14 px).  Real sensor dropouts
    are spatially clustered; i.i.d. per-pixel dropouts would put a depth edge inside every smoothing window and the
    reference's integral-image normals would be NaN almost everywhere."""
    h, w = shape
    mask = np.zeros(shape, dtype=bool)
    yy, xx = np.mgrid[0:h, 0:w]
    target = fraction * h * w
    covered = 0.0
    while covered < target:
        r = rng.uniform(2.0, 14.0)
        cx, cy = rng.uniform(0, w), rng.uniform(0, h)
        x0, x1 = int(max(cx - r, 0)), int(min(cx + r + 1, w))
        y0, y1 = int(max(cy - r, 0)), int(min(cy + r + 1, h))
        sub = (xx[y0:y1, x0:x1] - cx) ** 2 + (yy[y0:y1, x0:x1] - cy) ** 2 <= r * r
        mask[y0:y1, x0:x1] |= sub
        covered += np.pi * r * r
    return mask


def add_noise(depth: np.ndarray, frame_idx: int, model: str = "kinect") -> np.ndarray:
    """Noisy copy of one depth frame.  kinect: sigma_z = 0.0012 + 0.0019 (z-0.4)^2, 2 % dropouts (clustered discs), quantised through the
    uint16 / 5000 PNG encoding (DepthMapFactor, TUM1.yaml:35).  realsense: sigma_z = 0.001 z^2, 1 mm quantisation, 5 %
    dropouts."""
    rng = np.random.Generator(np.random.PCG64(1000 + int(frame_idx)))
    z = depth.astype(np.float64)
    valid = z > 0
    if model == "kinect":
        sigma = 0.0012 + 0.0019 * (z - 0.4) ** 2
        drop = 0.02
    elif model == "realsense":
        sigma = 0.001 * z ** 2
        drop = 0.05
    else:
        raise ValueError(model)
    zn = z + rng.normal(size=z.shape) * sigma
    zn[~valid] = 0.0
    zn[_dropout_mask(rng, z.shape, drop)] = 0.0
    zn = np.clip(zn, 0.0, 13.0)
    if model == "kinect":
        u16 = np.round(zn * 5000.0).astype(np.uint16)
        return (u16.astype(np.float32) * np.float32(1.0 / 5000.0)).astype(np.float32)
    return (np.round(zn * 1000.0) / 1000.0).astype(np.float32)
