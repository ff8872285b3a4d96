"""ctypes binding of libgsr_b200.so (include/gsr_b200.h).  No torch types cross this boundary: only
raw device pointers, sizes and a stream handle.

There is NO fallback: if the library is missing or a call fails, a ``GsrError`` is raised.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, byref, c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_uint32, c_uint64, c_void_p

from .camera import CameraStruct

PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.environ.get("GSR_B200_LIB") or os.path.join(PKG, "csrc", "libgsr_b200.so")  # override: A/B builds only

GSR_OK, GSR_ERR_INVALID_ARGUMENT, GSR_ERR_CUDA, GSR_ERR_WORKSPACE_TOO_SMALL, GSR_ERR_DUP_OVERFLOW = range(5)
REC_FLOATS = 12
TILE = 16


class GsrError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"libgsr_b200 status {status}: {message}")
        self.status = status


class Gaussians(ctypes.Structure):
    """``gsr_gaussians``: strided f32 views (element strides)."""

    _fields_ = [
        ("n", c_int64),
        ("xyz", c_void_p), ("xyz_sn", c_int64), ("xyz_sc", c_int64),
        ("log_scales", c_void_p), ("scale_sn", c_int64), ("scale_sc", c_int64),
        ("rots", c_void_p), ("rot_sn", c_int64), ("rot_sc", c_int64),
        ("opacity_logit", c_void_p), ("opacity_sn", c_int64),
        ("sh", c_void_p), ("sh_sn", c_int64), ("sh_sk", c_int64), ("sh_sc", c_int64),
    ]


class FrameInfo(ctypes.Structure):
    _fields_ = [("status", c_uint32), ("n_visible", c_uint32), ("n_dup", c_uint32), ("n_dup_peak", c_uint32),
                ("warp_evals", c_uint64), ("pixel_evals", c_uint64)]


# name -> argtypes, exactly the prototypes of include/gsr_b200.h
PROTOTYPES = {
    "gsr_abi_version": [],
    "gsr_device_info": [POINTER(c_int), POINTER(c_int), POINTER(c_int), POINTER(c_size_t)],
    "gsr_fp32_fma_peak": [POINTER(ctypes.c_double), POINTER(ctypes.c_double)],
    "gsr_debug_blend_variant": [c_int, c_int],
    "gsr_render_workspace_bytes": [c_int64, c_int32, c_int32, c_int64, POINTER(c_size_t)],
    "gsr_render_fwd": [POINTER(Gaussians), POINTER(CameraStruct), c_int64, c_void_p, c_size_t, c_void_p, c_void_p,
                       c_void_p, c_uint32, c_uint32, c_int32, c_int32, c_void_p, c_void_p],
    "gsr_prepared_bytes": [c_int64, POINTER(c_size_t)],
    "gsr_prepare": [POINTER(Gaussians), c_void_p, c_size_t, c_void_p],
    "gsr_preprocess_prepared_fwd": [c_void_p, c_int64, POINTER(CameraStruct), c_int32, c_int32, c_void_p, c_void_p,
                                    c_void_p, c_void_p, c_void_p, c_void_p],
    "gsr_render_prepared_fwd": [c_void_p, c_int64, POINTER(CameraStruct), c_int64, c_void_p, c_size_t, c_void_p,
                                c_void_p, c_void_p, c_uint32, c_uint32, c_int32, c_int32, c_void_p, c_void_p],
    "gsr_render_views_fwd": [c_void_p, c_int64, POINTER(CameraStruct), c_int32, c_int64, POINTER(c_void_p), c_size_t,
                             c_int32, POINTER(c_void_p), c_void_p, c_int64, c_void_p],
    "gsr_render_host": [POINTER(Gaussians), POINTER(CameraStruct), c_int64, c_void_p, POINTER(FrameInfo)],
    "gsr_host_renderer_create": [c_int64, c_int32, c_int32, c_int64, c_int32, POINTER(c_void_p)],
    "gsr_host_renderer_destroy": [c_void_p],
    "gsr_host_renderer_submit": [c_void_p, POINTER(Gaussians), POINTER(CameraStruct), c_void_p, POINTER(c_int32)],
    "gsr_host_renderer_wait": [c_void_p, c_int32, POINTER(FrameInfo)],
    "gsr_host_renderer_render": [c_void_p, POINTER(Gaussians), POINTER(CameraStruct), c_void_p, POINTER(FrameInfo)],
    "gsr_preprocess_fwd": [POINTER(Gaussians), POINTER(CameraStruct), c_int32, c_int32, c_void_p, c_void_p, c_void_p,
                           c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    "gsr_sh_fwd": [c_int64, c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int64, c_int64, POINTER(c_float), c_int32,
                   c_void_p, c_void_p],
    "gsr_sort_workspace_bytes": [c_int64, POINTER(c_size_t)],
    "gsr_sort_pairs": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int32, c_int32,
                       c_void_p, c_size_t, c_void_p],
    "gsr_bin_workspace_bytes": [c_int64, c_int32, c_int32, c_int64, POINTER(c_size_t)],
    "gsr_bin": [c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int64, c_void_p, c_void_p, c_void_p,
                c_void_p, c_void_p, c_void_p, c_size_t, c_void_p],
    "gsr_blend_fwd": [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_uint32, c_uint32,
                      c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    "gsr_blend_bwd": [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p,
                      c_void_p, c_void_p],
    "gsr_preprocess_bwd": [POINTER(Gaussians), POINTER(CameraStruct), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                           c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p],
    "gsr_render_bwd": [POINTER(Gaussians), POINTER(CameraStruct), c_int64, c_void_p, c_size_t, c_void_p, c_void_p,
                       c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64,
                       c_int32, c_int32, c_void_p],
    "gsr_cov3d_fwd": [c_int64, c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int64, c_void_p, c_void_p],
    "gsr_quat_to_rotmat": [c_int64, c_void_p, c_int64, c_int64, c_void_p, c_void_p],
    "gsr_affine_fwd": [c_int64, c_void_p, c_int64, c_int64, POINTER(c_float), c_int32, c_void_p, c_void_p],
    "gsr_cov2d_fwd": [c_int64, c_void_p, c_void_p, POINTER(CameraStruct), c_void_p, c_void_p],
    "gsr_bbox_fwd": [c_int64, c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p],
    "gsr_ply_unpack": [c_void_p, c_int64, c_int32, POINTER(c_int32), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                       c_void_p],
    "gsr_blend_one": [POINTER(c_int64), POINTER(c_float), POINTER(c_float), POINTER(c_float), c_float, c_int32,
                      c_int32, c_void_p, c_void_p, c_void_p],
}

_lib = None


def load() -> ctypes.CDLL:
    """Load the shared library and attach prototypes.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GsrError(-1, f"{LIB_PATH} not built; run `python -m gsr_b200.build` (nvcc, sm_100a). There is no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    lib.gsr_last_error.restype = c_char_p
    lib.gsr_last_error.argtypes = []
    for name, argtypes in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError here means header and library disagree
        fn.argtypes = argtypes
        fn.restype = c_int
    _lib = lib
    return lib


def check(status: int) -> None:
    if status != GSR_OK:
        raise GsrError(status, load().gsr_last_error().decode("utf-8", "replace"))


def call(name: str, *args) -> None:
    check(getattr(load(), name)(*args))


def require_cuda():
    """The product path has no CPU implementation."""
    import torch

    if not torch.cuda.is_available():
        raise GsrError(-2, "no CUDA device: the B200 render path has no CPU fallback")
    load()


def ptr(t) -> c_void_p:
    return c_void_p(t.data_ptr()) if t is not None else c_void_p(0)


def stream_handle() -> c_void_p:
    import torch

    return c_void_p(torch.cuda.current_stream().cuda_stream)


def device_info():
    sm, maj, mnr, l2 = c_int(), c_int(), c_int(), c_size_t()
    call("gsr_device_info", byref(sm), byref(maj), byref(mnr), byref(l2))
    return dict(sm_count=sm.value, cc=(maj.value, mnr.value), l2_bytes=l2.value)
