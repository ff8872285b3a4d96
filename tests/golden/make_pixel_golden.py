"""Attribute the worst full-size pixels against the UNMODIFIED reference (run in the build container).

Input : gpurun_out/attrib_<config>.npz written on the GPU box by profiles/attribute_pixels.py (the Gaussians whose tile
        rect holds the pixel's tile, raw parameters in original index order; the CUDA and oracle values of the pixel).
Output: tests/golden/pixel_<config>.npz - the cut-down scene plus the reference's own 16x16 tile around the pixel, and
        a line per pixel saying which side the reference agrees with.

    python tests/golden/make_pixel_golden.py C2_full [C5]
"""
import json
import logging
import os
import sys
import tempfile

import numpy as np

os.environ.setdefault("TQDM_DISABLE", "1")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(ROOT, "torch-gaussian-splatting-rasterizer_b200"), ROOT, HERE]

from gsr_b200 import scenes  # noqa: E402
from oracle import ref_harness  # noqa: E402


def main(config):
    logging.disable(logging.CRITICAL)
    a = np.load(os.path.join(ROOT, "gpurun_out", f"attrib_{config}.npz"), allow_pickle=False)
    W, H, degree = int(a["width"]), int(a["height"]), int(a["degree"])
    intr = scenes.Intrinsics.synthetic(W, H)
    keep = dict(width=a["width"], height=a["height"], degree=a["degree"], config=a["config"])
    report = []
    k = 0
    while f"p{k}_xy" in a.files:
        cloud = scenes.GaussianCloud(a[f"p{k}_xyz"], a[f"p{k}_f_dc"], a[f"p{k}_f_rest"], a[f"p{k}_opacity"],
                                     a[f"p{k}_scales"], a[f"p{k}_rots"])
        x, y = (int(v) for v in a[f"p{k}_xy"])
        with tempfile.TemporaryDirectory() as d:
            scenes.write_scene_dir(d, cloud, intr, [scenes.front_view(0)])
            out = ref_harness.run_reference(d, 0, stable=True, capture=False)
        tx, ty = x // 16, y // 16
        ref_tile = out["image"][ty * 16:ty * 16 + 16, tx * 16:tx * 16 + 16].copy()
        ref_px = out["image"][y, x]
        d_cuda = float(np.abs(ref_px - a[f"p{k}_cuda"]).max())
        d_oracle = float(np.abs(ref_px - a[f"p{k}_oracle"]).max())
        report.append(dict(config=config, x=x, y=y, n_gaussians=cloud.n, reference=ref_px.tolist(),
                           cuda=a[f"p{k}_cuda"].tolist(), oracle=a[f"p{k}_oracle"].tolist(),
                           abs_reference_minus_cuda=d_cuda, abs_reference_minus_oracle=d_oracle,
                           tile_maxabs_reference_vs_cuda=float(np.abs(ref_tile - a[f"p{k}_cuda_tile"]).max()),
                           tile_maxabs_reference_vs_oracle=float(np.abs(ref_tile - a[f"p{k}_oracle_tile"]).max()),
                           verdict="reference agrees with CUDA" if d_cuda < d_oracle else "reference agrees with the oracle"))
        for f in ("xy", "xyz", "f_dc", "f_rest", "opacity", "scales", "rots", "cuda", "oracle"):
            keep[f"p{k}_{f}"] = a[f"p{k}_{f}"]
        keep[f"p{k}_reference_tile"] = ref_tile
        keep[f"p{k}_reference_px"] = ref_px
        print(json.dumps(report[-1]), flush=True)
        k += 1
    np.savez_compressed(os.path.join(HERE, f"pixel_{config}.npz"), **keep)
    with open(os.path.join(ROOT, "profiles", f"r02_pixel_attribution_{config}.json"), "w") as fh:
        json.dump(report, fh, indent=1)


if __name__ == "__main__":
    for c in (sys.argv[1:] or ["C2_full"]):
        main(c)
