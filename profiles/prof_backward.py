"""Forward + backward timing of the differentiable render on one GPU (config C3 by default)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "torch-gaussian-splatting-rasterizer_b200"), ROOT]
import numpy as np, torch
from gsr_b200 import camera, render as R, scenes

def run(name):
    cloud, intr, views, degree = scenes.make_config(name)
    cam = camera.camera_for(intr, views[0], degree)
    model = R.GaussianModel.from_cloud(cloud)
    r = R.Renderer()
    H, W = cam.height, cam.width
    final_T = torch.empty((H, W), device="cuda"); n_contrib = torch.zeros((H, W), dtype=torch.int32, device="cuda")
    img = r.render(model, cam, final_T=final_T, n_contrib=n_contrib)
    G = torch.rand((H, W, 3), device="cuda")
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    fw = bw = 0.0
    reps = 5
    for i in range(reps + 2):
        ev[0].record()
        r.render(model, cam, out=img, final_T=final_T, n_contrib=n_contrib, check=False)
        ev[1].record()
        grads = r.backward(model, cam, final_T, n_contrib, G)
        ev[2].record(); torch.cuda.synchronize()
        if i >= 2:
            fw += ev[0].elapsed_time(ev[1]); bw += ev[1].elapsed_time(ev[2])
    out = dict(config=name, forward_ms=fw / reps, backward_ms=bw / reps,
               grad_norms=[float(g.float().norm()) for g in grads])
    print(json.dumps(out), flush=True)
    return out

if __name__ == "__main__":
    res = [run(n) for n in (sys.argv[1:] or ["C3"])]
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "backward_timing.json"), "w"), indent=1)
