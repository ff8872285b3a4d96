"""BASELINE.json's own configurations at FULL size, CUDA path vs the C oracle on identical inputs.

north_star's bar: max-abs <= 2e-3 in [0,1] RGB and PSNR > 50 dB.  The oracle (oracle/gsr_oracle.c, OpenMP) needs a few
seconds per frame at these sizes; it is pinned to outputs of the unmodified reference by tests/test_oracle_cpu.py.
Every run appends its numbers to gpurun_out/parity_fullsize.json (copied to profiles/ by the builder).
"""
import json
import os

import numpy as np
import pytest
import torch

from gsr_b200 import camera, scenes
from oracle import c_oracle

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# (config, view index): C2 at both readings of "~2600x1700 / 2", the headline C3, one orbit view of C4 (= C3's scene from
# another side: what every rank renders in the view-sharded run), and C5 (6 M Gaussians at 3840x2160).
CASES = [("C2", 0), ("C2_full", 0), ("C3", 0), ("C4", 37), ("C5", 0)]


def _record(row):
    path = os.path.join(ROOT, "gpurun_out", "parity_fullsize.json")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    rows = []
    if os.path.exists(path):
        with open(path) as fh:
            rows = json.load(fh)
    rows = [r for r in rows if (r["config"], r["view"]) != (row["config"], row["view"])] + [row]
    with open(path, "w") as fh:
        json.dump(rows, fh, indent=1)


@pytest.mark.parametrize("config,view", CASES)
def test_full_size_render_meets_the_bar(config, view):
    from gsr_b200 import render as R

    cloud, intr, views, degree = scenes.make_config(config)
    cam = camera.camera_for(intr, views[view], degree)
    o = c_oracle.render(cloud, cam)
    r = R.Renderer()
    model = R.GaussianModel.from_cloud(cloud)
    img_t = r.render(model, cam)                      # resident-scene path: prepared model
    info = r.frame_info()
    raw_t = r.render(model, cam, prepared=False)      # fused raw-parameter kernel
    assert torch.equal(img_t, raw_t), "prepared and raw preprocess paths differ"
    img = img_t.cpu().numpy()
    diff = np.abs(img - o["image"])
    per_px = diff.max(axis=2)
    mse = float(np.mean((img.astype(np.float64) - o["image"].astype(np.float64)) ** 2))
    psnr = 200.0 if mse == 0 else 10 * np.log10(1.0 / mse)
    over = np.argwhere(per_px > 2e-3)
    row = dict(config=config, view=view, n=cloud.n, width=cam.width, height=cam.height, max_abs=float(diff.max()),
               psnr_db=float(psnr), pixels_over_2e3=int(over.shape[0]), pixels_over_1e3=int((per_px > 1e-3).sum()),
               pixels_over_2e4=int((per_px > 2e-4).sum()), median=float(np.median(diff)),
               over_xy=[[int(x), int(y)] for y, x in over[:8]], n_visible=info["n_visible"], n_dup=info["n_dup"],
               oracle_visible=int((o["skip"] == 0).sum()))
    _record(row)
    assert info["n_visible"] == row["oracle_visible"]
    assert not img[-1].any() and not img[:, -1].any()  # Q8
    assert psnr > 50.0, row
    # The keep test alpha > 1/255 is discontinuous and torch's exp is not bit-reproducible: a pixel whose alpha lies
    # within an ulp of the threshold can flip between two float implementations (each flip is worth <= 3.9e-3 T c).
    # Zero pixels over the bar is the expectation; a pixel that does exceed it must be one of those attributed against
    # the unmodified reference in tests/golden/pixel_*.npz (test_attributed_pixels_agree_with_the_reference).
    known = _attributed(config)
    unexplained = [xy for xy in row["over_xy"] if tuple(xy) not in known]
    assert row["pixels_over_2e3"] == len([1 for xy in row["over_xy"] if tuple(xy) in known]) and not unexplained, row


def _attributed(config):
    path = os.path.join(ROOT, "tests", "golden", f"pixel_{config}.npz")
    if not os.path.exists(path):
        return set()
    g = np.load(path, allow_pickle=False)
    return {tuple(int(v) for v in g[k]) for k in g.files if k.endswith("_xy")}


@pytest.mark.parametrize("config", ["C2_full", "C5"])
def test_attributed_pixels_agree_with_the_reference(config):
    """Pixels where the CUDA path and the C oracle disagreed most at full size were re-run through the UNMODIFIED
    reference on the cut-down scene (the Gaussians whose rect holds the pixel's tile, original order): the reference's own
    16x16 tile is the golden.  The CUDA render of the same cut-down scene must match it within the bar."""
    from gsr_b200 import render as R

    path = os.path.join(ROOT, "tests", "golden", f"pixel_{config}.npz")
    if not os.path.exists(path):
        pytest.skip("no attributed pixel recorded for this config")
    g = np.load(path, allow_pickle=False)
    W, H, degree = int(g["width"]), int(g["height"]), int(g["degree"])
    intr = scenes.Intrinsics.synthetic(W, H)
    cam = camera.camera_for(intr, scenes.front_view(0), degree)
    k = 0
    while f"p{k}_xy" in g.files:
        cloud = scenes.GaussianCloud(g[f"p{k}_xyz"], g[f"p{k}_f_dc"], g[f"p{k}_f_rest"], g[f"p{k}_opacity"],
                                     g[f"p{k}_scales"], g[f"p{k}_rots"])
        x, y = (int(v) for v in g[f"p{k}_xy"])
        img = R.Renderer().render(R.GaussianModel.from_cloud(cloud), cam).cpu().numpy()
        tx, ty = x // 16, y // 16
        tile = img[ty * 16:ty * 16 + 16, tx * 16:tx * 16 + 16]
        ref_tile = g[f"p{k}_reference_tile"]
        err = float(np.abs(tile - ref_tile).max())
        assert err <= 2e-3, (config, x, y, err)
        k += 1
    assert k > 0
