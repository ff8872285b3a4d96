"""Per-view preprocess of one rank's band of tile rows (config C5 by default), the SH prefetch up front vs for survivors only:
    GSR_PRE_LATE_SH=0 python profiles/prof_band_pre.py ; GSR_PRE_LATE_SH=1 python profiles/prof_band_pre.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "torch-gaussian-splatting-rasterizer_b200"), ROOT]

import torch

from gsr_b200 import camera, render as R, scenes

cloud, intr, views, degree = scenes.make_config(os.environ.get("CONFIG", "C5"))
cam = camera.camera_for(intr, views[0], degree)
model = R.GaussianModel.from_cloud(cloud)
model.prepare()
st = R.Stages(model, cam.width, cam.height, 8 * cloud.n)
tiles_y = (cam.height + 15) // 16
parts = int(os.environ.get("PARTS", "8"))
for name, rows in (("full frame", (0, 1 << 30)), (f"band {parts // 2} of {parts}", (tiles_y * (parts // 2) // parts, tiles_y * (parts // 2 + 1) // parts))):
    for _ in range(3):
        st.preprocess(cam, rows=rows)
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    e[0].record()
    for _ in range(10):
        st.preprocess(cam, rows=rows)
    e[1].record()
    torch.cuda.synchronize()
    print("%-16s preprocess %.3f ms  (GSR_PRE_LATE_SH=%s)" % (name, e[0].elapsed_time(e[1]) / 10, os.environ.get("GSR_PRE_LATE_SH", "auto")))
