"""Host-side camera set-up: the O(1) per-view work the reference does on the CPU before any
per-Gaussian math (/root/reference/rasterize.py:336-364, spherical_harmonics.py:35).

The result is a ``gsr_camera`` C struct (include/gsr_b200.h) passed by value to every kernel
launch.  Values are produced with the same torch-CPU float32 operations, in the same order, as the
reference, because downstream decisions (depth order, ceil/floor of radius and tile rect) are
discontinuous in them.
"""
from __future__ import annotations

import ctypes
import math
from typing import Sequence

import numpy as np
import torch

Z_FAR = 100.0
Z_NEAR = 0.01


class CameraStruct(ctypes.Structure):
    """Mirror of ``gsr_camera`` in include/gsr_b200.h (and of the oracle's struct)."""

    _fields_ = [
        ("w2c_t", ctypes.c_float * 16),
        ("full_t", ctypes.c_float * 16),
        ("cam_center", ctypes.c_float * 3),
        ("focal_x", ctypes.c_float),
        ("focal_y", ctypes.c_float),
        ("limx", ctypes.c_float),
        ("limy", ctypes.c_float),
        ("width", ctypes.c_int32),
        ("height", ctypes.c_int32),
        ("sh_degree", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
    ]


def rotation_from_quaternion(q: torch.Tensor) -> torch.Tensor:
    """``[4,N]`` (w,x,y,z) -> ``[3,3,N]`` float32; no normalisation (reference rasterize.py:41-56)."""
    w, x, y, z = q[0], q[1], q[2], q[3]
    rows = (
        (1 - 2 * y ** 2 - 2 * z ** 2, 2 * x * y - 2 * z * w, 2 * x * z + 2 * y * w),
        (2 * x * y + 2 * z * w, 1 - 2 * x ** 2 - 2 * z ** 2, 2 * y * z - 2 * x * w),
        (2 * x * z - 2 * y * w, 2 * y * z + 2 * x * w, 1 - 2 * x ** 2 - 2 * y ** 2),
    )
    return torch.stack([torch.stack(r) for r in rows]).float()


def world_to_camera(qvec: torch.Tensor, tvec: torch.Tensor) -> torch.Tensor:
    """4x4 ``[R|t; 0 0 0 1]`` in float32 from float64 COLMAP pose (reference rasterize.py:59-77)."""
    m = torch.zeros((4, 4))
    m[:3, :3] = rotation_from_quaternion(qvec.reshape(4, 1))[:, :, 0]
    m[:3, 3] = tvec
    m[3, 3] = 1
    return m


def perspective(fov_x: float, fov_y: float) -> torch.Tensor:
    """3DGS-style perspective matrix (reference rasterize.py:123-151); only columns feeding x, y, w matter."""
    tx, ty = math.tan(fov_x / 2), math.tan(fov_y / 2)
    top, right = ty * Z_NEAR, tx * Z_NEAR
    bottom, left = -top, -right
    p = torch.zeros(4, 4)
    p[0, 0] = 2.0 * Z_NEAR / (right - left)
    p[1, 1] = 2.0 * Z_NEAR / (top - bottom)
    p[0, 2] = (right + left) / (right - left)
    p[1, 2] = (top + bottom) / (top - bottom)
    p[3, 2] = 1.0
    p[2, 2] = Z_FAR / (Z_FAR - Z_NEAR)
    p[2, 3] = -(Z_FAR * Z_NEAR) / (Z_FAR - Z_NEAR)
    return p


def make_camera(qvec: Sequence[float], tvec: Sequence[float], fx: float, fy: float, cam_width: int,
                cam_height: int, width: int, height: int, degree: int = 3) -> CameraStruct:
    """Camera struct for one view.

    ``fx, fy, cam_width, cam_height`` are the COLMAP (full-resolution) intrinsics, ``width, height``
    the rendered image size; the EWA focal is ``fx/2`` whatever the scale factor (reference
    rasterize.py:216,336-345 - SURVEY.md Q3).
    """
    if not 0 <= int(degree) <= 3:
        raise ValueError("SH degree must be 0..3")
    fov_x = 2 * np.arctan(cam_width / (2 * fx))
    fov_y = 2 * np.arctan(cam_height / (2 * fy))
    tan_x, tan_y = math.tan(fov_x * 0.5), math.tan(fov_y * 0.5)
    q = torch.as_tensor(np.asarray(qvec, dtype=np.float64))
    t = torch.as_tensor(np.asarray(tvec, dtype=np.float64))
    w2c_t = world_to_camera(q, t).transpose(0, 1)
    proj_t = perspective(fov_x, fov_y).transpose(0, 1)
    full_t = w2c_t.unsqueeze(0).bmm(proj_t.unsqueeze(0)).squeeze(0)
    centre = w2c_t.inverse()[3, :3]
    cam = CameraStruct()
    cam.w2c_t[:] = w2c_t.contiguous().reshape(-1).tolist()
    cam.full_t[:] = full_t.contiguous().reshape(-1).tolist()
    cam.cam_center[:] = centre.tolist()
    cam.focal_x = float(np.float32(fx / 2))
    cam.focal_y = float(np.float32(fy / 2))
    cam.limx = float(np.float32(1.3 * tan_x))
    cam.limy = float(np.float32(1.3 * tan_y))
    cam.width, cam.height, cam.sh_degree, cam.reserved = int(width), int(height), int(degree), 0
    return cam


def camera_for(intr, view, degree: int = 3) -> CameraStruct:
    """Convenience: ``scenes.Intrinsics`` + ``scenes.View`` -> struct."""
    return make_camera(view.qvec, view.tvec, intr.fx, intr.fy, intr.cam_width, intr.cam_height, intr.width,
                       intr.height, degree)
