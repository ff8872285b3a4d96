import sys, os, torch, numpy as np
sys.path[:0]=[os.path.join(os.getcwd(),'torch-gaussian-splatting-rasterizer_b200'), os.getcwd()]
from gsr_b200 import scenes, camera, render as R
n=int(os.environ.get("N","3000000"))
cloud=scenes.synth_v1(n,3); intr=scenes.Intrinsics.synthetic(1920,1080)
cam=camera.camera_for(intr,scenes.front_view(0),3)
model=R.GaussianModel.from_cloud(cloud)
for i in range(3):
    pre=R.preprocess(model,cam); binned=R.bin_tiles(pre,1920,1080,8*n)
torch.cuda.synchronize()
e=[torch.cuda.Event(enable_timing=True) for _ in range(3)]
e[0].record(); pre=R.preprocess(model,cam); e[1].record(); binned=R.bin_tiles(pre,1920,1080,8*n); e[2].record(); torch.cuda.synchronize()
print("preprocess ms", e[0].elapsed_time(e[1]), "bin ms (incl. allocs)", e[1].elapsed_time(e[2]))
