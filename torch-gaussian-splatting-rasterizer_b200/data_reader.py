"""COLMAP binary model readers used by the render path (drop-in for the reference's ``data_reader``).

Only what ``read_scene`` needs is implemented: ``sparse/0/images.bin`` and ``sparse/0/cameras.bin``
(/root/reference/data_reader.py:140-198).  The reference's text / points3D / dense-depth readers are never
called by its render path and are out of scope.  File formats follow COLMAP's
``src/base/reconstruction.cc`` (Read/Write{Images,Cameras}Binary).
"""
from __future__ import annotations

import collections
import struct

import numpy as np

CameraModel = collections.namedtuple("CameraModel", ["model_id", "model_name", "num_params"])
Camera = collections.namedtuple("Camera", ["id", "model", "width", "height", "params"])
BaseImage = collections.namedtuple("Image", ["id", "qvec", "tvec", "camera_id", "name", "xys", "point3D_ids"])
Point3D = collections.namedtuple("Point3D", ["id", "xyz", "rgb", "error", "image_ids", "point2D_idxs"])

_MODEL_TABLE = (
    (0, "SIMPLE_PINHOLE", 3), (1, "PINHOLE", 4), (2, "SIMPLE_RADIAL", 4), (3, "RADIAL", 5), (4, "OPENCV", 8),
    (5, "OPENCV_FISHEYE", 8), (6, "FULL_OPENCV", 12), (7, "FOV", 5), (8, "SIMPLE_RADIAL_FISHEYE", 4),
    (9, "RADIAL_FISHEYE", 5), (10, "THIN_PRISM_FISHEYE", 12),
)
CAMERA_MODELS = {CameraModel(*row) for row in _MODEL_TABLE}
CAMERA_MODEL_IDS = {m.model_id: m for m in CAMERA_MODELS}
CAMERA_MODEL_NAMES = {m.model_name: m for m in CAMERA_MODELS}

_IMAGE_HEAD = struct.Struct("<idddddddi")   # image_id, qvec[4], tvec[3], camera_id
_CAMERA_HEAD = struct.Struct("<iiQQ")       # camera_id, model_id, width, height
_U64 = struct.Struct("<Q")


def read_next_bytes(fid, num_bytes, format_char_sequence, endian_character="<"):
    """Read ``num_bytes`` from ``fid`` and unpack them (kept for API compatibility)."""
    return struct.unpack(endian_character + format_char_sequence, fid.read(num_bytes))


def read_extrinsics_binary(path_to_model_file):
    """``images.bin`` -> {image_id: BaseImage}; 2-D observations are returned as arrays like the reference."""
    with open(path_to_model_file, "rb") as fh:
        blob = fh.read()
    images, off = {}, 0
    (count,) = _U64.unpack_from(blob, off)
    off += _U64.size
    for _ in range(count):
        head = _IMAGE_HEAD.unpack_from(blob, off)
        off += _IMAGE_HEAD.size
        end = blob.index(b"\x00", off)
        name = blob[off:end].decode("utf-8")
        off = end + 1
        (n2d,) = _U64.unpack_from(blob, off)
        off += _U64.size
        obs = np.frombuffer(blob, dtype=np.dtype([("x", "<f8"), ("y", "<f8"), ("id", "<i8")]), count=n2d, offset=off)
        off += 24 * n2d
        images[head[0]] = BaseImage(id=head[0], qvec=np.array(head[1:5]), tvec=np.array(head[5:8]), camera_id=head[8],
                                    name=name, xys=np.column_stack([obs["x"], obs["y"]]).astype(np.float64),
                                    point3D_ids=obs["id"].astype(np.int64))
    return images


def read_intrinsics_binary(path_to_model_file):
    """``cameras.bin`` -> {camera_id: Camera(model name, width, height, params)}."""
    with open(path_to_model_file, "rb") as fh:
        blob = fh.read()
    cameras, off = {}, 0
    (count,) = _U64.unpack_from(blob, off)
    off += _U64.size
    for _ in range(count):
        cam_id, model_id, width, height = _CAMERA_HEAD.unpack_from(blob, off)
        off += _CAMERA_HEAD.size
        model = CAMERA_MODEL_IDS[model_id]
        params = np.frombuffer(blob, dtype="<f8", count=model.num_params, offset=off).copy()
        off += 8 * model.num_params
        cameras[cam_id] = Camera(id=cam_id, model=model.model_name, width=width, height=height, params=params)
    assert len(cameras) == count
    return cameras
