"""Seeded synthetic scenes and on-disk scene directories.

A *scene directory* is what the reference's ``run_rasterization`` consumes
(/root/reference/rasterize.py:328-353):

    root/sparse/0/cameras.bin                       COLMAP intrinsics  (data_reader.py:176-198 format)
    root/sparse/0/images.bin                        COLMAP extrinsics  (data_reader.py:140-173 format)
    root/images_<scale>/<name>                      any W x H image; only its size is used (rasterize.py:333-338)
    root/point_cloud/iteration_30000/point_cloud.ply   3DGS-layout PLY  (rasterize.py:353)

Distributions are frozen: ``numpy.random.default_rng(seed)`` (PCG64), draw order
xyz, scales, rots, opacity, f_dc, f_rest, everything cast to float32.
"""
from __future__ import annotations

import math
import os
import struct
from dataclasses import dataclass
from typing import List, Optional, Tuple

import numpy as np

from .ply import GS_PROPERTIES, write_gaussian_ply

SH_0 = 0.28209479177387814


@dataclass
class GaussianCloud:
    """Host-side 3DGS parameters, raw (pre-activation) as stored in the PLY."""

    xyz: np.ndarray       # [N,3] f32
    f_dc: np.ndarray      # [N,3] f32
    f_rest: np.ndarray    # [N,45] f32, channel-major: f_rest[:, 15*c + (k-1)] (utils.py:23-28)
    opacity: np.ndarray   # [N] f32 logits
    scales: np.ndarray    # [N,3] f32 log-scales
    rots: np.ndarray      # [N,4] f32 (w,x,y,z), un-normalised

    @property
    def n(self) -> int:
        return int(self.xyz.shape[0])

    def ply_table(self) -> np.ndarray:
        t = np.zeros((self.n, len(GS_PROPERTIES)), dtype=np.float32)
        t[:, 0:3] = self.xyz
        t[:, 6:9] = self.f_dc
        t[:, 9:54] = self.f_rest
        t[:, 54] = self.opacity
        t[:, 55:58] = self.scales
        t[:, 58:62] = self.rots
        return t

    def sh_planar(self) -> np.ndarray:
        """SH coefficients as the contiguous ``[3,16,N]`` block the reference builds (utils.py:21-31)."""
        out = np.empty((3, 16, self.n), dtype=np.float32)
        out[:, 0, :] = self.f_dc.T
        out[:, 1:, :] = self.f_rest.T.reshape(3, 15, self.n)
        return out

    def subset(self, idx) -> "GaussianCloud":
        return GaussianCloud(self.xyz[idx], self.f_dc[idx], self.f_rest[idx], self.opacity[idx],
                             self.scales[idx], self.rots[idx])


@dataclass
class View:
    """One COLMAP image: world->camera quaternion (w,x,y,z) and translation, both float64."""

    image_id: int
    qvec: np.ndarray
    tvec: np.ndarray
    name: str = "view.png"


@dataclass
class Intrinsics:
    """COLMAP PINHOLE camera (full-resolution) plus the size of the image actually rendered.

    The reference takes fx, fy, width, height from ``cam_info[1]`` and W, H from the
    ``images_<scale>`` file (rasterize.py:333-345); cx, cy are read and ignored.
    """

    cam_width: int
    cam_height: int
    fx: float
    fy: float
    width: int     # rendered W
    height: int    # rendered H

    @staticmethod
    def synthetic(width: int, height: int) -> "Intrinsics":
        return Intrinsics(2 * width, 2 * height, 1.2 * (2 * width), 1.2 * (2 * width), width, height)


def synth_v1(n: int, seed: int, scale_factor: float = 0.2) -> GaussianCloud:
    rng = np.random.default_rng(seed)
    xyz = rng.uniform(-1.5, 1.5, (n, 3))
    s0 = scale_factor * (27.0 / n) ** (1.0 / 3.0)
    scales = rng.normal(math.log(s0), 0.5, (n, 3))
    rots = rng.normal(0.0, 1.0, (n, 4))
    opacity = rng.normal(0.0, 2.0, n)
    f_dc = rng.normal(0.0, 1.0, (n, 3))
    f_rest = rng.normal(0.0, 0.15, (n, 45))
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    return GaussianCloud(f32(xyz), f32(f_dc), f32(f_rest), f32(opacity), f32(scales), f32(rots))


def random_init(n: int, seed: int) -> GaussianCloud:
    """'random-init' cloud of BASELINE.json config 1: SH degree 0, identity rotations, opacity 0.1."""
    rng = np.random.default_rng(seed)
    xyz = rng.uniform(-1.5, 1.5, (n, 3))
    scales = np.log(rng.uniform(0.5, 1.5, (n, 3)) * 0.554 * (27.0 / n) ** (1.0 / 3.0))
    rots = np.zeros((n, 4)); rots[:, 0] = 1.0
    opacity = np.full(n, math.log(0.1 / 0.9))
    f_dc = (rng.uniform(0.0, 1.0, (n, 3)) - 0.5) / 0.28209479
    f_rest = np.zeros((n, 45))
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    return GaussianCloud(f32(xyz), f32(f_dc), f32(f_rest), f32(opacity), f32(scales), f32(rots))


def front_view(image_id: int = 0) -> View:
    return View(image_id, np.array([1.0, 0.0, 0.0, 0.0]), np.array([0.0, 0.0, 4.0]))


def orbit_views(count: int, total: Optional[int] = None, first: int = 0) -> List[View]:
    """Camera centre on a radius-4 circle in the xz-plane looking at the origin (R_k = R_y(2 pi k / total))."""
    total = count if total is None else total
    out = []
    for k in range(first, first + count):
        half = math.pi * k / total
        out.append(View(k, np.array([math.cos(half), 0.0, math.sin(half), 0.0]), np.array([0.0, 0.0, 4.0])))
    return out


# BASELINE.json:configs, in order.  (name, generator, N, W, H, degree, seed)
CONFIGS = {
    "C1": dict(gen="random_init", n=50_000, width=640, height=420, degree=0, seed=1, views=1),
    "C2": dict(gen="synth_v1", n=500_000, width=1297, height=840, degree=3, seed=2, views=1),
    "C2_full": dict(gen="synth_v1", n=500_000, width=2594, height=1681, degree=3, seed=2, views=1),
    "C3": dict(gen="synth_v1", n=3_000_000, width=1920, height=1080, degree=3, seed=3, views=1),
    "C4": dict(gen="synth_v1", n=3_000_000, width=1920, height=1080, degree=3, seed=3, views=256),
    "C5": dict(gen="synth_v1", n=6_000_000, width=3840, height=2160, degree=3, seed=5, views=1),
}


def make_cloud(gen: str, n: int, seed: int) -> GaussianCloud:
    if gen == "synth_v1":
        return synth_v1(n, seed)
    if gen == "random_init":
        return random_init(n, seed)
    raise ValueError(gen)


def make_config(name: str, n_override: Optional[int] = None) -> Tuple[GaussianCloud, Intrinsics, List[View], int]:
    c = CONFIGS[name]
    n = c["n"] if n_override is None else n_override
    cloud = make_cloud(c["gen"], n, c["seed"])
    intr = Intrinsics.synthetic(c["width"], c["height"])
    views = [front_view(0)] if c["views"] == 1 else orbit_views(c["views"])
    return cloud, intr, views, c["degree"]


# ----------------------------------------------------------------------------------------------
# scene directory writer (COLMAP binary model + PLY + dummy image)
# ----------------------------------------------------------------------------------------------

def write_cameras_bin(path: str, intr: Intrinsics, camera_id: int = 1) -> None:
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "wb") as fh:
        fh.write(struct.pack("<Q", 1))
        fh.write(struct.pack("<iiQQ", camera_id, 1, intr.cam_width, intr.cam_height))  # model 1 = PINHOLE
        fh.write(struct.pack("<dddd", intr.fx, intr.fy, intr.cam_width / 2.0, intr.cam_height / 2.0))


def write_images_bin(path: str, views: List[View], camera_id: int = 1) -> None:
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "wb") as fh:
        fh.write(struct.pack("<Q", len(views)))
        for v in views:
            fh.write(struct.pack("<idddddddi", v.image_id, *[float(q) for q in v.qvec],
                                 *[float(t) for t in v.tvec], camera_id))
            fh.write(v.name.encode("utf-8") + b"\x00")
            fh.write(struct.pack("<Q", 0))


def write_scene_dir(root: str, cloud: GaussianCloud, intr: Intrinsics, views: List[View], scale_factor: int = 2) -> str:
    from PIL import Image

    write_cameras_bin(os.path.join(root, "sparse", "0", "cameras.bin"), intr)
    write_images_bin(os.path.join(root, "sparse", "0", "images.bin"), views)
    img_dir = os.path.join(root, f"images_{scale_factor}")
    os.makedirs(img_dir, exist_ok=True)
    for name in {v.name for v in views}:
        Image.new("RGB", (intr.width, intr.height)).save(os.path.join(img_dir, name))
    write_gaussian_ply(os.path.join(root, "point_cloud", "iteration_30000", "point_cloud.ply"), cloud.ply_table())
    return root
