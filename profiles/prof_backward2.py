"""Forward-for-training + backward timing on C3 (GPU box): python profiles/prof_backward2.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "torch-gaussian-splatting-rasterizer_b200"), ROOT]
import torch
from gsr_b200 import camera, render as R, scenes

cloud, intr, views, degree = scenes.make_config(os.environ.get("CONFIG", "C3"))
cam = camera.camera_for(intr, views[0], degree)
model = R.GaussianModel.from_cloud(cloud)
H, W = cam.height, cam.width
fT = torch.empty((H, W), device="cuda"); nC = torch.zeros((H, W), dtype=torch.int32, device="cuda")
G = torch.rand((H, W, 3), device="cuda"); frame = torch.empty((H, W, 3), device="cuda")
r = R.Renderer()
r.render(model, cam, out=frame, final_T=fT, n_contrib=nC, prepared=False)
for _ in range(2):
    r.render(model, cam, out=frame, final_T=fT, n_contrib=nC, check=False, prepared=False); g = r.backward(model, cam, fT, nC, G)
torch.cuda.synchronize()
e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
e[0].record()
for _ in range(10):
    r.render(model, cam, out=frame, final_T=fT, n_contrib=nC, check=False, prepared=False)
e[1].record()
for _ in range(10):
    g = r.backward(model, cam, fT, nC, G)
e[2].record(); torch.cuda.synchronize()
print("forward(tracked) %.3f ms, backward %.3f ms, |d_xyz|sum %.6e" % (e[0].elapsed_time(e[1]) / 10, e[1].elapsed_time(e[2]) / 10, float(g[0].double().abs().sum())))
