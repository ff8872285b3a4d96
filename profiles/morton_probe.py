"""Probe (run on the GPU box): what would a spatially ordered prepared model buy?  The same C3 cloud in its random order
and permuted on the HOST along a Morton curve; per-view preprocess time of the prepared kernel (whole-warp early out for
blocks that are off screen as a whole), whole-frame time, image difference (tie order may differ: this is a probe)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "torch-gaussian-splatting-rasterizer_b200"), ROOT]

import numpy as np
import torch

from gsr_b200 import camera, render as R, scenes


def morton_order(xyz, bits=10):
    lo, hi = xyz.min(axis=0), xyz.max(axis=0)
    q = np.clip(((xyz - lo) / np.maximum(hi - lo, 1e-30) * ((1 << bits) - 1)).astype(np.uint64), 0, (1 << bits) - 1)
    code = np.zeros(xyz.shape[0], np.uint64)
    for b in range(bits):
        for c in range(3):
            code |= ((q[:, c] >> np.uint64(b)) & np.uint64(1)) << np.uint64(3 * b + c)
    return np.argsort(code, kind="stable")


def run(cloud, cam, tag):
    model = R.GaussianModel.from_cloud(cloud)
    r = R.Renderer(); img = r.render(model, cam).clone()
    st = R.Stages(model, cam.width, cam.height, r.max_dup)
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    acc = np.zeros(3)
    for it in range(25):
        e[0].record(); st.preprocess(cam); e[1].record(); st.bin(); e[2].record(); st.blend(); e[3].record()
        torch.cuda.synchronize()
        if it >= 5:
            acc += [e[k].elapsed_time(e[k + 1]) for k in range(3)]
    out = dict(order=tag, preprocess_ms=acc[0] / 20, binning_ms=acc[1] / 20, blend_ms=acc[2] / 20, n_visible=r.frame_info()["n_visible"])
    print(json.dumps(out), flush=True)
    return img, out


if __name__ == "__main__":
    cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
    cloud, intr, views, degree = scenes.make_config(cfg)
    rows = []
    for vi in (0, 37) if len(views) > 37 else (0,):
        cam = camera.camera_for(intr, views[vi], degree)
        a, ra = run(cloud, cam, "as generated (random)")
        b, rb = run(cloud.subset(morton_order(cloud.xyz)), cam, "Morton")
        rows += [ra, rb, dict(view=vi, image_maxabs_diff=float((a - b).abs().max()))]
        print(json.dumps(rows[-1]), flush=True)
    with open(os.path.join(ROOT, "gpurun_out", f"morton_probe_{cfg}.json"), "w") as fh:
        json.dump(rows, fh, indent=1)
