"""Attribution of the worst pixels of a full-size render (run on the GPU box).

For each config: render with the CUDA path and with the C oracle, take the pixels where they differ most, and for each
dump what is needed to re-run that pixel through the UNMODIFIED reference in the build container
(tests/golden/make_pixel_golden.py): the Gaussians of the pixel's tile list (raw parameters, original index order), the
camera, both pixel values, and a per-Gaussian trace (power, alpha of both sides) naming the keep decision that differs.

    python profiles/attribute_pixels.py C2_full C5   ->  gpurun_out/attrib_<config>.npz, gpurun_out/attrib.json
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "torch-gaussian-splatting-rasterizer_b200"), ROOT]

import numpy as np
import torch

from gsr_b200 import camera, render as R, scenes
from oracle import c_oracle


def trace_pixel(o, gids, x, y):
    """Reference arithmetic (numpy f32, the oracle's order) for one pixel over its tile list: per Gaussian
    power, o*exp(power) and the keep decision."""
    f = np.float32
    rows = []
    for g in gids:
        mx, my = o["mean2d"][g]
        a, c, b = o["conic"][g]
        dx, dy = f(mx - f(x)), f(my - f(y))
        power = f(f(f(-0.5) * f(f(a * f(dx * dx)) + f(c * f(dy * dy)))) - f(f(b * dx) * dy))
        e64 = np.exp(np.float64(power))
        alpha = min(f(o["opacity"][g] * f(e64)), f(0.99))
        rows.append((int(g), float(power), float(alpha), float(o["opacity"][g] * e64), bool(alpha > f(1 / 255) and power <= 0)))
    return rows


def run(name, top=2, view=0):
    cloud, intr, views, degree = scenes.make_config(name)
    cam = camera.camera_for(intr, views[view], degree)
    o = c_oracle.render(cloud, cam)
    model = R.GaussianModel.from_cloud(cloud)
    pre = R.preprocess(model, cam)
    n_pairs = int((pre["tiles_touched"] & 0x7FFFFFFF).sum().item())
    binned = R.bin_tiles(pre, cam.width, cam.height, n_pairs + 1024)
    out = R.blend(pre, binned, cam.width, cam.height)
    img = out["image"].cpu().numpy()
    diff = np.abs(img - o["image"]).max(axis=2)
    flat = np.argsort(diff.reshape(-1))[::-1][:top]
    tiles_x = (cam.width + 15) // 16
    rng = binned["tile_ranges"].cpu().numpy().astype(np.int64)
    gids_all = binned["tile_gids"].cpu().numpy().astype(np.int64)
    summary, blobs = [], {}
    for k, p in enumerate(flat):
        y, x = divmod(int(p), cam.width)
        t = (y // 16) * tiles_x + x // 16
        lst = gids_all[rng[t, 0]:rng[t, 1]]  # depth order (after culling: pairs that cannot touch the tile are gone)
        # the reference evaluates every Gaussian whose rect holds the tile; culled pairs contribute nothing by proof,
        # but keep them in the cut-down scene anyway so the reference sees its full list
        r = o["tile_rect"].astype(np.int64)
        tx, ty = x // 16, y // 16
        full = np.nonzero((o["skip"] == 0) & (r[:, 0] <= tx) & (tx < r[:, 2]) & (r[:, 1] <= ty) & (ty < r[:, 3]))[0]
        ids = np.sort(full)
        sub = cloud.subset(ids)
        tr = trace_pixel(o, [g for g in o["order"] if g in set(full.tolist())], x, y)
        near = [row for row in tr if abs(row[2] - 1 / 255) < 2e-5]
        summary.append(dict(config=name, x=x, y=y, tile=int(t), diff=float(diff[y, x]), cuda=img[y, x].tolist(),
                            oracle=o["image"][y, x].tolist(), list_len=int(lst.shape[0]), rect_len=int(ids.shape[0]),
                            near_threshold=near[:8], final_T_cuda=float(out["final_T"][y, x].item()),
                            final_T_oracle=float(o["final_T"][y, x])))
        blobs.update({f"p{k}_ids": ids, f"p{k}_xy": np.array([x, y]), f"p{k}_cuda": img[y, x], f"p{k}_oracle": o["image"][y, x],
                      f"p{k}_xyz": sub.xyz, f"p{k}_f_dc": sub.f_dc, f"p{k}_f_rest": sub.f_rest, f"p{k}_opacity": sub.opacity,
                      f"p{k}_scales": sub.scales, f"p{k}_rots": sub.rots,
                      f"p{k}_cuda_tile": img[ty * 16:ty * 16 + 16, tx * 16:tx * 16 + 16],
                      f"p{k}_oracle_tile": o["image"][ty * 16:ty * 16 + 16, tx * 16:tx * 16 + 16]})
    blobs["width"], blobs["height"], blobs["degree"] = np.int64(cam.width), np.int64(cam.height), np.int64(degree)
    blobs["config"] = np.array(name)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    np.savez_compressed(os.path.join(ROOT, "gpurun_out", f"attrib_{name}.npz"), **blobs)
    for s in summary:
        print(json.dumps(s), flush=True)
    return summary


if __name__ == "__main__":
    res = []
    for n in (sys.argv[1:] or ["C2_full"]):
        res += run(n)
    with open(os.path.join(ROOT, "gpurun_out", "attrib.json"), "w") as fh:
        json.dump(res, fh, indent=1)
