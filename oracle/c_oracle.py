"""ctypes front-end of oracle/gsr_oracle.c.  TEST INFRASTRUCTURE ONLY (never imported by the product)."""
from __future__ import annotations

import ctypes
from typing import Dict, Optional

import numpy as np

from . import build as _build

_lib = None
_f = ctypes.POINTER(ctypes.c_float)


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(_build.build())
        _lib.gsr_oracle_blend.restype = ctypes.c_int64
        _lib.gsr_oracle_count_evals.restype = ctypes.c_int64
        _lib.gsr_oracle_threads.restype = ctypes.c_int
    return _lib


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _sh_strides(sh: np.ndarray):
    """sh is indexed [n,k,c]; return element strides."""
    assert sh.dtype == np.float32 and sh.ndim == 3 and sh.shape[1:] == (16, 3)
    return tuple(s // 4 for s in sh.strides)


def sh_view(cloud) -> np.ndarray:
    """[N,16,3] view of the planar [3,16,N] block, like the reference's tensor (utils.py:31)."""
    return cloud.sh_planar().transpose(2, 1, 0)


def preprocess(cloud, cam, sh: Optional[np.ndarray] = None) -> Dict[str, np.ndarray]:
    n = cloud.n
    sh = sh_view(cloud) if sh is None else sh
    sn, sk, sc = _sh_strides(sh)
    out = dict(
        cov3d=np.empty((n, 3, 3), np.float32), cam_means=np.empty((n, 3), np.float32),
        mean2d=np.empty((n, 2), np.float32), cov2d=np.empty((n, 2, 2), np.float32),
        conic=np.empty((n, 3), np.float32), tile_rect=np.empty((n, 4), np.int32),
        pixel_bbox=np.empty((n, 4), np.int32), rgb=np.empty((n, 3), np.float32),
        opacity=np.empty(n, np.float32), depth=np.empty(n, np.float32), radius=np.empty(n, np.float32),
        skip=np.empty(n, np.uint8),
    )
    c = np.ascontiguousarray
    lib().gsr_oracle_preprocess(
        ctypes.c_int64(n), _p(c(cloud.xyz)), _p(c(cloud.scales)), _p(c(cloud.rots)), _p(c(cloud.opacity)),
        _p(sh), ctypes.c_int64(sn), ctypes.c_int64(sk), ctypes.c_int64(sc), ctypes.byref(cam),
        _p(out["cov3d"]), _p(out["cam_means"]), _p(out["mean2d"]), _p(out["cov2d"]), _p(out["conic"]),
        _p(out["tile_rect"]), _p(out["pixel_bbox"]), _p(out["rgb"]), _p(out["opacity"]), _p(out["depth"]),
        _p(out["radius"]), _p(out["skip"]))
    out["_sh_keepalive"] = sh
    return out


def argsort_depth(depth: np.ndarray) -> np.ndarray:
    depth = np.ascontiguousarray(depth, np.float32)
    order = np.empty(depth.shape[0], np.int64)
    lib().gsr_oracle_argsort(ctypes.c_int64(depth.shape[0]), _p(depth), _p(order))
    return order


def blend(pre: Dict[str, np.ndarray], order: np.ndarray, width: int, height: int, first: int = 0,
          last: Optional[int] = None):
    """Returns (image [H,W,3], final transmittance [H,W], number of Gaussians blended)."""
    screen = np.zeros((width, height, 3), np.float32)
    T = np.ones((width, height), np.float32)
    last = order.shape[0] if last is None else last
    c = np.ascontiguousarray  # fixtures loaded from .npz may be Fortran-ordered
    order = c(order, dtype=np.int64)
    pre = dict(skip=c(pre["skip"], dtype=np.uint8), pixel_bbox=c(pre["pixel_bbox"], dtype=np.int32),
               mean2d=c(pre["mean2d"], dtype=np.float32), conic=c(pre["conic"], dtype=np.float32),
               rgb=c(pre["rgb"], dtype=np.float32), opacity=c(pre["opacity"], dtype=np.float32))
    nb = lib().gsr_oracle_blend(
        ctypes.c_int64(order.shape[0]), _p(order), _p(pre["skip"]), _p(pre["pixel_bbox"]), _p(pre["mean2d"]),
        _p(pre["conic"]), _p(pre["rgb"]), _p(pre["opacity"]), ctypes.c_int(width), ctypes.c_int(height),
        _p(screen), _p(T), ctypes.c_int64(first), ctypes.c_int64(last))
    return np.ascontiguousarray(screen.transpose(1, 0, 2)), np.ascontiguousarray(T.T), int(nb)


def render(cloud, cam, first: int = 0, last: Optional[int] = None) -> Dict[str, np.ndarray]:
    """Full oracle render of one view; returns image plus every intermediate."""
    pre = preprocess(cloud, cam)
    order = argsort_depth(pre["depth"])
    img, T, nb = blend(pre, order, cam.width, cam.height, first, last)
    pre.update(order=order, image=img, final_T=T, n_blended=nb)
    return pre


def sh_to_rgb(xyz: np.ndarray, sh: np.ndarray, cam_center, degree: int) -> np.ndarray:
    xyz = np.ascontiguousarray(xyz, np.float32)
    sn, sk, sc = _sh_strides(sh)
    cc = np.asarray(cam_center, np.float32)
    out = np.empty((xyz.shape[0], 3), np.float32)
    lib().gsr_oracle_sh(ctypes.c_int64(xyz.shape[0]), _p(xyz), _p(sh), ctypes.c_int64(sn), ctypes.c_int64(sk),
                        ctypes.c_int64(sc), _p(cc), ctypes.c_int(degree), _p(out))
    return out


def count_evals(pre) -> int:
    return int(lib().gsr_oracle_count_evals(ctypes.c_int64(pre["skip"].shape[0]), _p(pre["skip"]),
                                            _p(pre["pixel_bbox"])))


def tile_lists(pre: Dict[str, np.ndarray], width: int, height: int):
    """Per-tile Gaussian lists the binning stage must reproduce: sorted by (tile, depth bits, index).

    A Gaussian touches tile (tx,ty) when tx0 <= tx < tx1 and ty0 <= ty < ty1 of its tile rect
    (clipped to the tile grid) and the reference does not skip it (rasterize.py:415-420,441).
    Returns (tile_ids[n_dup], gaussian_ids[n_dup], ranges[tiles,2]).
    """
    tiles_x, tiles_y = (width + 15) // 16, (height + 15) // 16
    r = pre["tile_rect"].astype(np.int64)
    x0, y0 = np.minimum(r[:, 0], tiles_x), np.minimum(r[:, 1], tiles_y)
    x1, y1 = np.minimum(r[:, 2], tiles_x), np.minimum(r[:, 3], tiles_y)
    keep = (pre["skip"] == 0) & (x1 > x0) & (y1 > y0)
    ids = np.nonzero(keep)[0]
    tile_l, gid_l = [], []
    for g in ids:
        tx, ty = np.meshgrid(np.arange(x0[g], x1[g]), np.arange(y0[g], y1[g]), indexing="xy")
        t = (ty * tiles_x + tx).reshape(-1)
        tile_l.append(t)
        gid_l.append(np.full(t.shape[0], g, np.int64))
    if not tile_l:
        return np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros((tiles_x * tiles_y, 2), np.int64)
    tile = np.concatenate(tile_l); gid = np.concatenate(gid_l)
    dbits = pre["depth"].view(np.uint32).astype(np.int64)[gid]  # visible => z >= 0.2 > 0: bits are monotone
    perm = np.lexsort((gid, dbits, tile))
    tile, gid = tile[perm], gid[perm]
    starts = np.searchsorted(tile, np.arange(tiles_x * tiles_y), side="left")
    ends = np.searchsorted(tile, np.arange(tiles_x * tiles_y), side="right")
    return tile, gid, np.stack([starts, ends], axis=1)


def pair_hits(pre: Dict[str, np.ndarray], gids: np.ndarray, tiles: np.ndarray, width: int, height: int) -> np.ndarray:
    """For each (Gaussian, tile) pair: does any pixel of the tile pass the reference's keep test?"""
    c = np.ascontiguousarray
    gids, tiles = c(gids, dtype=np.int64), c(tiles, dtype=np.int64)
    hits = np.zeros(gids.shape[0], np.uint8)
    mean2d, conic, op = c(pre["mean2d"], dtype=np.float32), c(pre["conic"], dtype=np.float32), c(pre["opacity"], dtype=np.float32)
    lib().gsr_oracle_pair_hits(ctypes.c_int64(gids.shape[0]), _p(gids), _p(tiles), ctypes.c_int((width + 15) // 16),
                               ctypes.c_int(width), ctypes.c_int(height), _p(mean2d), _p(conic), _p(op), _p(hits))
    return hits
