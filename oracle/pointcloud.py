"""ORACLE (test infrastructure, not product code): numpy restatement of depth -> point cloud.

Follows reference depth2pointcloud.py:12-31 (point_cloud) and :103-108 (camera -> world with
cam.matrix_world). The reference file is a Blender script (imports bpy/mathutils and carries
U+200B characters on blank lines) and cannot be imported, but its `point_cloud` function only needs
numpy and `tan`: oracle/gen_golden.py compiles that function from the reference source and stores its
outputs in tests/golden/pointcloud.npz, and this restatement reproduces them BIT FOR BIT
(tests/test_oracle_vs_golden.py, tests/test_oracle_vs_reference.py) - PINNED. Only `to_world`
(:103-108, a per-point mathutils product inside Blender) remains a restatement without a reference run.
"""
from __future__ import annotations

from math import tan

import numpy as np


def point_cloud(depth, angle_x, clip_start, clip_end):
    """depth [H,W] float32 -> [H,W,3] float64 (np.dstack promotes; z is -depth or NaN)."""
    factor = 2.0 * tan(angle_x / 2.0)
    rows, cols = depth.shape
    c, r = np.meshgrid(np.arange(cols), np.arange(rows), sparse=True)
    valid = (depth > clip_start) & (depth < clip_end)
    z = -np.where(valid, depth, np.nan)
    ratio = max(rows, cols)
    x = -np.where(valid, factor * z * (c - (cols / 2)) / ratio, 0)
    y = np.where(valid, factor * z * (r - (rows / 2)) / ratio, 0)
    return np.dstack((x, y, z))


def to_world(points, matrix_world):
    """depth2pointcloud.py:103-108: each point p -> matrix_world @ (p, 1). mathutils promotes a
    3-vector to (x,y,z,1) for a 4x4 product and returns the xyz part; mathutils stores fp32."""
    M = np.asarray(matrix_world, dtype=np.float32).astype(np.float64)
    P = points.reshape(-1, 3).astype(np.float32).astype(np.float64)
    out = P @ M[:3, :3].T + M[:3, 3]
    return out.reshape(points.shape)
