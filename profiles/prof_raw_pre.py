"""Time the raw-parameter per-Gaussian kernel (training forward / one-shot renders) on config C3:
    GSR_B200_LIB=... python profiles/prof_raw_pre.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "torch-gaussian-splatting-rasterizer_b200"), ROOT]

import torch

from gsr_b200 import camera, render as R, scenes

cloud, intr, views, degree = scenes.make_config(os.environ.get("CONFIG", "C3"))
cam = camera.camera_for(intr, views[0], degree)
model = R.GaussianModel.from_cloud(cloud)
r = R.Renderer()
img = r.render(model, cam)
st = R.Stages(model, cam.width, cam.height, r.max_dup)
for _ in range(3):
    st.preprocess(cam, prepared=False)
torch.cuda.synchronize()
e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
e[0].record()
for _ in range(10):
    st.preprocess(cam, prepared=False)
e[1].record()
torch.cuda.synchronize()
print("raw preprocess %.3f ms  image checksum %.6f" % (e[0].elapsed_time(e[1]) / 10, float(img.double().sum())))
