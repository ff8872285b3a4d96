"""Full-size parity of the CUDA path against the C oracle on BASELINE.json configs (run on the GPU box).

    python profiles/parity_full.py C3 [C2 C2_full C5 ...]  ->  gpurun_out/parity_full.json
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "torch-gaussian-splatting-rasterizer_b200"), ROOT]

import numpy as np
import torch

from gsr_b200 import camera, render as R, scenes
from oracle import c_oracle


def run(name):
    cloud, intr, views, degree = scenes.make_config(name)
    cam = camera.camera_for(intr, views[0], degree)
    t0 = time.time(); o = c_oracle.render(cloud, cam); t_oracle = time.time() - t0
    r = R.Renderer()
    model = R.GaussianModel.from_cloud(cloud)
    img = r.render(model, cam)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        r.render(model, cam, out=img, check=False)
    e1.record(); torch.cuda.synchronize()
    info = r.frame_info()
    got = img.cpu().numpy()
    diff = np.abs(got - o["image"])
    mse = float(np.mean((got.astype(np.float64) - o["image"].astype(np.float64)) ** 2))
    bad = diff.max(axis=2) > 2e-3
    out = dict(config=name, n=cloud.n, width=cam.width, height=cam.height, max_abs=float(diff.max()),
               psnr_db=float(200.0 if mse == 0 else 10 * np.log10(1.0 / mse)), pixels_over_2e3=int(bad.sum()),
               pixels_over_1e3=int((diff.max(axis=2) > 1e-3).sum()), p999=float(np.quantile(diff, 0.999)),
               median=float(np.median(diff)), n_visible=info["n_visible"], n_dup=info["n_dup"],
               oracle_visible=int((o["skip"] == 0).sum()), last_row_col_zero=bool(not got[-1].any() and not got[:, -1].any()),
               cuda_ms_per_frame=e0.elapsed_time(e1) / 10, oracle_c_openmp_seconds=t_oracle,
               oracle_threads=int(c_oracle.lib().gsr_oracle_threads()),
               reference_evals=int(c_oracle.count_evals(o)))
    print(json.dumps(out), flush=True)
    return out


if __name__ == "__main__":
    res = [run(n) for n in (sys.argv[1:] or ["C3"])]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "parity_full.json"), "w") as fh:
        json.dump(res, fh, indent=1)
