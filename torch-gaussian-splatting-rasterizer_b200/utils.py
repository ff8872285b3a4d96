"""Scene / PLY attribute access (drop-in for the reference's ``utils``)."""
from __future__ import annotations

import os
from typing import Dict, Tuple

import numpy as np
import torch

from data_reader import BaseImage, Camera, read_extrinsics_binary, read_intrinsics_binary
from gsr_b200.ply import PlyData, PlyElement  # noqa: F401  (stand-in for the un-vendored `plyfile`)


def read_color_components(plydata) -> torch.Tensor:
    """SH coefficients as ``[N,16,3]``: ``sh[n,0,c] = f_dc_c`` and ``sh[n,k,c] = f_rest_{15c+k-1}``.

    Like the reference (/root/reference/utils.py:21-31) the storage is the contiguous planar block ``[3,16,N]``
    and the result is its transposed view - which is also the layout the preprocess kernel reads fastest.
    """
    el = plydata.elements[0]
    n = len(el["f_dc_0"])
    planar = np.empty((3, 16, n), dtype=np.float32)
    for c in range(3):
        planar[c, 0] = el[f"f_dc_{c}"]
        for k in range(1, 16):
            planar[c, k] = el[f"f_rest_{15 * c + k - 1}"]
    return torch.from_numpy(planar).transpose(2, 0)


def read_scene(path_to_scene: str) -> Tuple[Dict[int, BaseImage], Dict[int, Camera]]:
    """``(images keyed by COLMAP image_id, cameras keyed by camera_id)`` from ``<scene>/sparse/0``."""
    model_dir = os.path.join(path_to_scene, "sparse/0")
    return (read_extrinsics_binary(os.path.join(model_dir, "images.bin")),
            read_intrinsics_binary(os.path.join(model_dir, "cameras.bin")))
