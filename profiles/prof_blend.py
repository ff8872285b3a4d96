import sys, os, torch, numpy as np
sys.path[:0]=[os.path.join(os.getcwd(),'torch-gaussian-splatting-rasterizer_b200'), os.getcwd()]
from gsr_b200 import scenes, camera, render as R
n=int(os.environ.get("N","3000000"))
cloud=scenes.synth_v1(n,3); intr=scenes.Intrinsics.synthetic(1920,1080)
cam=camera.camera_for(intr,scenes.front_view(0),3)
model=R.GaussianModel.from_cloud(cloud)
pre=R.preprocess(model,cam); binned=R.bin_tiles(pre,1920,1080,8*n)
for i in range(3):
    out=R.blend(pre,binned,1920,1080,track=False)
torch.cuda.synchronize()
e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
e0.record(); out=R.blend(pre,binned,1920,1080,track=False); e1.record(); torch.cuda.synchronize()
print("blend ms", e0.elapsed_time(e1))
