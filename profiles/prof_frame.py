"""One frame of config C3 stage by stage, for ncu (run on the GPU box):

    python profiles/prof_frame.py && \
    ncu --set full --clock-control none --import-source on -k regex:"^(blend|onesweep|prepare|preprocess|scan_emit|tile_ranges)" -s 31 -c 11 -o gpurun_out/r02_frame python profiles/prof_frame.py

Launch order of OUR kernels (-k regex:gsr): prepare (1), one checked frame (10), two warm-up frames (20), then the
captured ones: gsr_prepare again (1) and one frame (10): preprocess, 4 depth-sort passes, scan+emit, 2 tile-sort passes,
tile ranges, blend.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "torch-gaussian-splatting-rasterizer_b200"), ROOT]

import torch

from gsr_b200 import camera, render as R, scenes

config = os.environ.get("CONFIG", "C3")
cloud, intr, views, degree = scenes.make_config(config)
cam = camera.camera_for(intr, views[0], degree)
model = R.GaussianModel.from_cloud(cloud)
model.prepare()
r = R.Renderer()
r.render(model, cam)
st = R.Stages(model, cam.width, cam.height, r.max_dup)
for _ in range(2):
    st.preprocess(cam); st.bin(); st.blend()
torch.cuda.synchronize()
model.prepare(force=True)
e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
e[0].record(); st.preprocess(cam); e[1].record(); st.bin(); e[2].record(); st.blend(); e[3].record()
torch.cuda.synchronize()
print("preprocess %.3f ms, binning %.3f ms, blend %.3f ms" % tuple(e[k].elapsed_time(e[k + 1]) for k in range(3)))
