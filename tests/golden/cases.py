"""Scene recipes behind the golden fixtures (shared by make_golden.py and the tests)."""
from __future__ import annotations

import hashlib

import numpy as np

from gsr_b200 import scenes

# name -> scene recipe.  Sizes that are not multiples of 16 exercise partial tiles.
CASES = {
    "synth_small": dict(gen="synth_v1", n=2000, seed=7, scale=0.6, width=160, height=112, degree=3,
                        qvec=[1, 0, 0, 0], tvec=[0, 0, 4]),
    "random_init": dict(gen="random_init", n=1500, seed=11, scale=None, width=150, height=100, degree=3,
                        qvec=[1, 0, 0, 0], tvec=[0, 0, 4]),
    "inside": dict(gen="synth_v1", n=1200, seed=13, scale=0.5, width=131, height=77, degree=3,
                   qvec=[1, 0, 0, 0], tvec=[0.1, -0.2, 0.6]),
    "orbit_a": dict(gen="synth_v1", n=1500, seed=17, scale=0.6, width=120, height=90, degree=3,
                    qvec="orbit:5/256", tvec=[0, 0, 4]),
    "orbit_b": dict(gen="synth_v1", n=1500, seed=17, scale=0.6, width=120, height=90, degree=3,
                    qvec="orbit:77/256", tvec=[0, 0, 4]),
    "ties": dict(gen="ties", n=800, seed=19, scale=1.0, width=96, height=64, degree=3,
                 qvec=[1, 0, 0, 0], tvec=[0, 0, 4]),
    "tilted": dict(gen="synth_v1", n=1000, seed=23, scale=0.8, width=100, height=75, degree=3,
                   qvec=[0.9, 0.1, -0.3, 0.2], tvec=[0.3, -0.1, 3.5]),  # un-normalised qvec: used as given (Q2)
    # needle-like splats in every orientation: det = c00 c11 - c01^2 cancels by up to 5 digits in f32, so the conic the
    # blend evaluates is far from the covariance's inverse - the reference's own result in that regime is the golden
    "needles": dict(gen="needles", n=240, seed=123, scale=None, width=333, height=217, degree=3,
                    qvec=[1, 0, 0, 0], tvec=[0, 0, 4]),
}
GRAD_CASE = dict(gen="synth_v1", n=48, seed=29, scale=1.2, width=48, height=32, degree=3,
                 qvec=[1, 0, 0, 0], tvec=[0, 0, 4])


def needles(n: int, seed: int):
    """Long thin splats (1.5 or 0.4 world units by 0.002 by 0.002), random orientations, opacity 0.88, strong colours."""
    rng = np.random.default_rng(seed)
    xyz = rng.uniform(-1.2, 1.2, (n, 3)).astype(np.float32)
    scales = np.tile(np.log(np.array([1.5, 0.002, 0.002], np.float32)), (n, 1))
    scales[n // 2:, 0] = np.log(0.4)                      # a shorter kind: rects under the 64-tile keep-mask limit too
    rots = rng.normal(size=(n, 4)).astype(np.float32)
    f_dc = rng.uniform(0.5, 1.7, (n, 3)).astype(np.float32)
    return scenes.GaussianCloud(xyz, f_dc, np.zeros((n, 45), np.float32), np.full(n, 2.0, np.float32),
                                scales.astype(np.float32), rots)


def build_case(c):
    if c["gen"] == "synth_v1":
        cloud = scenes.synth_v1(c["n"], c["seed"], scale_factor=c["scale"])
    elif c["gen"] == "random_init":
        cloud = scenes.random_init(c["n"], c["seed"])
    elif c["gen"] == "ties":
        cloud = scenes.synth_v1(c["n"], c["seed"], scale_factor=c["scale"])
        cloud.xyz[:, 2] = np.round(cloud.xyz[:, 2] * 4) / 4  # front view: depth = z + 4 -> many exact ties
        cloud.opacity[:] = np.abs(cloud.opacity) + 1.0       # opaque enough that blend order is visible
    elif c["gen"] == "needles":
        cloud = needles(c["n"], c["seed"])
    else:
        raise ValueError(c["gen"])
    q = c["qvec"]
    if isinstance(q, str):
        k, tot = q.split(":")[1].split("/")
        view = scenes.orbit_views(1, total=int(tot), first=int(k))[0]
        view.image_id = 0
    else:
        view = scenes.View(0, np.array(q, dtype=np.float64), np.array(c["tvec"], dtype=np.float64))
    intr = scenes.Intrinsics.synthetic(c["width"], c["height"])
    return cloud, intr, view


def cloud_digest(cloud) -> str:
    h = hashlib.sha256()
    for a in (cloud.xyz, cloud.f_dc, cloud.f_rest, cloud.opacity, cloud.scales, cloud.rots):
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()
