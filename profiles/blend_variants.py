"""Blend kernel variants on config C3 in one process (run on the GPU box):
staging path (cp.async per CTA / cp.async.bulk + mbarrier / cp.async per warp) x batched rejection test.

    python profiles/blend_variants.py [C3]  ->  gpurun_out/blend_variants.json
Every variant must reproduce the reference variant's image bit for bit; times are CUDA-event means over 30 launches of
the blend stage alone (tile lists built once), evaluations counted by the instrumented instantiation.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "torch-gaussian-splatting-rasterizer_b200"), ROOT]

import numpy as np
import torch

from gsr_b200 import _cabi, camera, render as R, scenes

NAMES = {0: "cp.async per CTA (128-record batches, 2 barriers per batch)", 1: "cp.async.bulk + mbarrier (TMA unit)",
         2: "cp.async per warp (64-record batches, no CTA barrier)"}


def main(config="C3"):
    cloud, intr, views, degree = scenes.make_config(config)
    cam = camera.camera_for(intr, views[0], degree)
    model = R.GaussianModel.from_cloud(cloud)
    r = R.Renderer(); r.render(model, cam)
    st = R.Stages(model, cam.width, cam.height, r.max_dup)
    st.preprocess(cam); st.bin()
    torch.cuda.synchronize()
    out, ref_img = [], None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for stage, pre in ((0, 0), (2, 0), (1, 0), (0, 0), (2, 0)):
        _cabi.call("gsr_debug_blend_variant", stage, pre)
        for _ in range(5):
            st.blend()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(30):
            st.blend()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 30
        img = st.image.clone()
        if ref_img is None:
            ref_img = img
        count = torch.zeros(8, dtype=torch.int32, device="cuda")
        st.blend(count); torch.cuda.synchronize()
        c = count.cpu().numpy().view(np.uint32)
        row = dict(stage=stage, stage_name=NAMES[stage], pretest=bool(pre), blend_ms=ms,
                   bit_identical_to_first=bool(torch.equal(img, ref_img)),
                   warp_evals=int(c[4]) | (int(c[5]) << 32))
        print(json.dumps(row), flush=True)
        out.append(row)
    _cabi.call("gsr_debug_blend_variant", -1, 0)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    tag = os.environ.get("VARIANT_TAG", "")
    with open(os.path.join(ROOT, "gpurun_out", f"blend_variants_{config}{tag}.json"), "w") as fh:
        json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main(*(sys.argv[1:2] or ["C3"]))
