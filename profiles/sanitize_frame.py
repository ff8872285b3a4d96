"""Every kernel of the library on small scenes, for compute-sanitizer ("race detection / sanitizers:
compute-sanitizer memcheck / racecheck on small scenes, determinism = bit-identical image across runs").

    for tool in memcheck racecheck synccheck; do
      compute-sanitizer --tool $tool --error-exitcode 9 --log-file gpurun_out/sanitize_$tool.log \
          python profiles/sanitize_frame.py
    done

Sizes are chosen so that every multi-CTA path runs (several one-sweep tiles per sort pass and the decoupled look-back of
the emission: 40 k Gaussians, > 100 k (Gaussian, tile) pairs, a ragged last block of 32, partial tiles on both image
edges) while a tool that slows kernels down 10-100x still finishes in about a minute.  Prints one line per section;
exits non-zero if a section's own check fails (the sanitizer's verdict is its exit code / log).
"""
import ctypes
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "torch-gaussian-splatting-rasterizer_b200"), ROOT]

import numpy as np
import torch

from gsr_b200 import _cabi, camera, ply, render as R, scenes

N = int(os.environ.get("SANITIZE_N", "40013"))     # not a multiple of 32: ragged last block of the prepared layout
W, H = 333, 217                                    # partial tiles on the right and bottom edges
cloud = scenes.synth_v1(N, 7, scale_factor=0.3)
intr = scenes.Intrinsics.synthetic(W, H)
views = scenes.orbit_views(3, total=16, first=2)
cams = [camera.camera_for(intr, v, 3) for v in views]
model = R.GaussianModel.from_cloud(cloud)


def same(a, b, what):
    if not torch.equal(a, b):
        raise SystemExit(f"{what}: not bit-identical (max-abs {float((a - b).abs().max()):.3e})")


# 1. resident scene: gsr_prepare + prepared per-view kernel, binning, blend; run to run and prepared vs raw
r = R.Renderer()
img_a = r.render(model, cams[0]).clone()
info = r.frame_info()
img_b = r.render(model, cams[0]).clone()
same(img_a, img_b, "prepared render, run to run")
img_raw = r.render(model, cams[0], prepared=False).clone()
same(img_a, img_raw, "prepared vs raw-parameter render")
print(f"frame ok: n_visible {info['n_visible']}, n_dup {info['n_dup']}")

# 2. every SH degree, raw and prepared kernels
for deg in range(4):
    c = camera.camera_for(intr, views[1], deg)
    same(r.render(model, c).clone(), r.render(model, c, prepared=False).clone(), f"SH degree {deg}")
print("sh degrees ok")

# 3. the three staging variants of the blend kernel (cp.async per CTA, cp.async.bulk + mbarrier, cp.async per warp)
for stage in (0, 1, 2):
    _cabi.call("gsr_debug_blend_variant", stage, 0)
    same(img_a, r.render(model, cams[0]).clone(), f"blend staging variant {stage}")
_cabi.call("gsr_debug_blend_variant", -1, 0)
print("blend variants ok")

# 4. progressive render (rank window) and a band of tile rows
part = r.render(model, cams[0], first_rank=0, last_rank=info["n_visible"] // 2).clone()
full = r.render(model, cams[0], first_rank=0, last_rank=info["n_visible"] + 5).clone()
same(img_a, full, "rank window covering every Gaussian")
rows = (3, 9)
band = torch.zeros_like(img_a)
r.render(model, cams[0], out=band, tile_rows=rows)
same(band[rows[0] * 16:rows[1] * 16], img_a[rows[0] * 16:rows[1] * 16], "band of tile rows")
print(f"rank window / band ok (half-depth image differs from the full one: {not torch.equal(part, img_a)})")

# 5. pair-capacity overflow: detected on the device, redrawn with the reported capacity
small = R.Renderer(dup_factor=0.25)
small.adaptive_capacity = False
same(img_a, small.render(model, cams[0]).clone(), "frame redrawn after a pair-capacity overflow")
print("overflow redraw ok")

# 6. a batch of views through one C call, and the lane pool
batch = R.BatchRenderer(lanes=2).render(model, cams)
lanes = R.render_views(model, intr, views, 3, lanes=2)
same(batch, lanes, "gsr_render_views_fwd vs per-lane renderers")
same(batch[0], img_a, "batched view 0")
print("batched views ok")

# 7. training path: forward with contributor tracking + backward (blend_bwd, preprocess_bwd)
params = [t.clone().requires_grad_(True) for t in (model.xyz, model.log_scales, model.rots, model.opacity, model.sh)]
tr = R.Renderer()
image = R.render_differentiable(*params, cams[2], renderer=tr)
weights = torch.linspace(0.5, 1.5, image.numel(), device=image.device).reshape(image.shape)
(image * weights).sum().backward()
g1 = [p.grad.clone() for p in params]
for p in params:
    p.grad = None
image = R.render_differentiable(*params, cams[2], renderer=tr, tile_rows=(0, 7))
image2 = R.render_differentiable(*params, cams[2], renderer=R.Renderer(), tile_rows=(7, 14))
((image * weights).sum() + (image2 * weights).sum()).backward()
worst = max(float((p.grad - g).abs().max() / (g.abs().max() + 1e-20)) for p, g in zip(params, g1))
if not worst < 1e-4:
    raise SystemExit(f"band partial gradients do not sum to the full gradient: {worst}")
print(f"backward ok (sum of two band partials vs full gradient: {worst:.1e} of the tensor scale)")

# 8. host-buffer pipeline (chunked upload + per-chunk raw preprocess) and the one-shot host entry
host = [torch.from_numpy(a) for a in (cloud.xyz, cloud.scales, cloud.rots, cloud.opacity)]
sh_host = torch.from_numpy(cloud.sh_planar()).permute(2, 1, 0)
hr = R.HostRenderer(N, W, H, depth=2)
out0 = torch.empty((H, W, 3)).pin_memory()
out1 = torch.empty((H, W, 3)).pin_memory()
t0 = hr.submit(*host, sh_host, cams[0], out0)
t1 = hr.submit(*host, sh_host, cams[1], out1)
hr.wait(t0); hr.wait(t1)
hr.close()
same(out0.cuda(), img_a, "host pipeline, frame 0")
one = torch.empty((H, W, 3))
g_host, keep = R._host_struct(*host, sh_host)
fi = _cabi.FrameInfo()
_cabi.call("gsr_render_host", ctypes.byref(g_host), ctypes.byref(cams[0]), 8 * N, _cabi.ptr(one), ctypes.byref(fi))
same(one.cuda(), img_a, "gsr_render_host")
print("host entries ok")

# 9. PLY vertex block -> device SoA (gsr_ply_unpack)
with tempfile.TemporaryDirectory() as d:
    path = os.path.join(d, "cloud.ply")
    ply.write_gaussian_ply(path, cloud.ply_table())
    m2 = R.GaussianModel.from_ply(path)
    same(r.render(m2, cams[0]).clone(), img_a, "PLY device loader")
print("ply unpack ok")

# 10. the radix sort alone: multi-tile, all 32 bits, device-side count
gen = torch.Generator(device="cuda").manual_seed(1)
keys = torch.randint(-2**31, 2**31 - 1, (100_003,), dtype=torch.int32, device="cuda", generator=gen)
vals = torch.arange(keys.numel(), dtype=torch.int32, device="cuda")
ko, vo = R.sort_pairs(keys, vals)
ref = np.argsort(keys.cpu().numpy().view(np.uint32), kind="stable")
if not np.array_equal(vo.cpu().numpy(), ref):
    raise SystemExit("radix sort differs from numpy's stable argsort")
print("sort ok")

# 11. empty and tiny inputs
e_model = R.GaussianModel.from_cloud(cloud.subset(np.arange(0)))
if float(R.Renderer().render(e_model, cams[0]).abs().max()) != 0.0:
    raise SystemExit("empty cloud is not black")
tiny = R.GaussianModel.from_cloud(scenes.synth_v1(5, 1, scale_factor=0.3))
R.Renderer().render(tiny, camera.camera_for(scenes.Intrinsics.synthetic(17, 9), views[0], 3))
torch.cuda.synchronize()
print("edge sizes ok")
print("SANITIZE-SCRIPT-DONE")
