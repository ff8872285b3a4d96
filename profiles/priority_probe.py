"""Probe (GPU box): frames in flight with the latency-bound binning kernels on HIGH-priority streams.
Three lanes; per lane preprocess and blend run on a low-priority stream, binning on a high-priority one, chained with events.
Prints ms per frame for (a) the plain per-lane single stream, (b) the split-priority arrangement."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "torch-gaussian-splatting-rasterizer_b200"), ROOT]
import torch
from gsr_b200 import camera, render as R, scenes

cloud, intr, views, degree = scenes.make_config("C3")
cam = camera.camera_for(intr, views[0], degree)
model = R.GaussianModel.from_cloud(cloud)
model.prepare()
r0 = R.Renderer(); r0.render(model, cam)
LANES, STEPS = 3, 150
stages = [R.Stages(model, cam.width, cam.height, r0.max_dup) for _ in range(LANES)]
lo_pri, hi_pri = torch.cuda.Stream.priority_range() if hasattr(torch.cuda.Stream, "priority_range") else (0, -1)


def run(split):
    low = [torch.cuda.Stream(priority=0) for _ in range(LANES)]
    high = [torch.cuda.Stream(priority=-1) for _ in range(LANES)] if split else low
    def frame(i):
        st, L, Hs = stages[i % LANES], low[i % LANES], high[i % LANES]
        with torch.cuda.stream(L):
            st.preprocess(cam)
            if split:
                e = torch.cuda.Event(); e.record(L); Hs.wait_event(e)
        with torch.cuda.stream(Hs):
            st.bin()
            if split:
                e = torch.cuda.Event(); e.record(Hs); L.wait_event(e)
        with torch.cuda.stream(L):
            st.blend()
    for i in range(9):
        frame(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ev = torch.cuda.Event(); ev.record()
    for s in low + (high if split else []):
        s.wait_event(ev)
    for i in range(STEPS):
        frame(i)
    for s in low + (high if split else []):
        torch.cuda.current_stream().wait_stream(s)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / STEPS


for rep in range(2):
    print("single stream per lane: %.4f ms/frame" % run(False), flush=True)
    print("binning on high-priority streams: %.4f ms/frame" % run(True), flush=True)
