"""Depth map -> point cloud on B200: drop-in for point_cloud() of the reference's
depth2pointcloud.py:12-31 with the camera->world transform of :103-108 fused.

The reference takes a Blender camera object (cam.data.angle_x / clip_start / clip_end,
cam.matrix_world); any object with those attributes works here, as does the small `Camera` struct.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from types import SimpleNamespace

import numpy as np
import torch

from . import _lib

__all__ = ["Camera", "point_cloud", "point_cloud_world", "write_ply"]


@dataclass
class Camera:
    angle_x: float
    clip_start: float
    clip_end: float
    matrix_world: object = None  # 4x4, row-major

    @property
    def data(self):  # mimic bpy: cam.data.angle_x
        return SimpleNamespace(angle_x=self.angle_x, clip_start=self.clip_start, clip_end=self.clip_end)


def _launch(depth, cam, matrix, dtype):
    lib = _lib.load()
    was_numpy = isinstance(depth, np.ndarray)
    if was_numpy:
        if not torch.cuda.is_available():
            raise RuntimeError("point_cloud needs a CUDA device; there is no CPU fallback")
        depth = torch.from_numpy(np.ascontiguousarray(depth, dtype=np.float32)).cuda()
    dev = _lib.require_cuda(depth)
    d = depth.detach().to(torch.float32).contiguous()
    h, w = int(d.shape[-2]), int(d.shape[-1])
    n_img = d.numel() // (h * w)
    data = cam.data
    mat = None
    if matrix is not None:
        m = np.asarray(matrix, dtype=np.float32).reshape(16)
        mat = (C.c_float * 16)(*[float(v) for v in m])
    with torch.cuda.device(dev):
        out = torch.empty(tuple(d.shape) + (3,), dtype=dtype, device=dev)
        _lib.check(lib.mde_point_cloud(_lib.ptr(d), n_img, h, w, float(data.angle_x), float(data.clip_start),
                                       float(data.clip_end), mat, 1 if dtype == torch.float64 else 0, _lib.ptr(out),
                                       _lib.stream_ptr(dev)))
    return out.cpu().numpy() if was_numpy else out


def point_cloud(depth, cam, dtype=torch.float64):
    """depth [H,W] (or [...,H,W]) -> [...,H,W,3] camera-space points; float64 as the reference's numpy
    promotion gives (pass dtype=torch.float32 for 12 B/px output). numpy in -> numpy out."""
    return _launch(depth, cam, None, dtype)


def point_cloud_world(depth, cam, dtype=torch.float64):
    """point_cloud followed by `cam.matrix_world @ p` for every point (depth2pointcloud.py:103-108), fused."""
    return _launch(depth, cam, cam.matrix_world, dtype)


def write_ply(path, front_points, color_bgr, back_points=None):
    """The ASCII PLY file of reference depth2pointcloud.py:131-154: for every pixel the front point, then the back
    point, each skipped when its x is NaN; colours arrive as cv2 gives them (BGR) and are written RGB with alpha 0.
    front_points / back_points: [..., 3] world coordinates (numpy or tensor), color_bgr: uint8 [..., 3] with the same
    number of pixels. Host-side writer (C ABI mde_write_ply). Returns the number of vertices written."""
    lib = _lib.load()

    def host(x, dtype):
        if isinstance(x, torch.Tensor):
            x = x.detach().cpu().numpy()
        return np.ascontiguousarray(np.asarray(x, dtype=dtype).reshape(-1, 3))

    f = host(front_points, np.float64)
    c = host(color_bgr, np.uint8)
    if c.shape[0] != f.shape[0]:
        raise ValueError("color must hold one BGR triple per point")
    b = None
    if back_points is not None:
        b = host(back_points, np.float64)
        if b.shape[0] != f.shape[0]:
            raise ValueError("front and back point sets must have the same number of pixels")
    n = lib.mde_write_ply(str(path).encode(), C.c_void_p(f.ctypes.data), None if b is None else C.c_void_p(b.ctypes.data),
                          C.c_void_p(c.ctypes.data), f.shape[0])
    if n < 0:
        _lib.check(int(n))
    return int(n)
