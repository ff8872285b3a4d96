"""Edge cases of the render path on the GPU, each against the C oracle on identical inputs."""
import numpy as np
import pytest
import torch

from gsr_b200 import camera, scenes
from oracle import c_oracle

pytestmark = pytest.mark.gpu


def _render_both(cloud, cam, **kw):
    from gsr_b200 import render as R

    r = R.Renderer(**kw)
    img = r.render(R.GaussianModel.from_cloud(cloud), cam).cpu().numpy()
    ref = c_oracle.render(cloud, cam)
    return img, ref, r


def _one(xyz, log_scale, logit, rgb_dc=(1.0, 0.2, -0.5)):
    n = len(xyz)
    z = lambda *s: np.zeros(s, np.float32)
    rots = z(n, 4); rots[:, 0] = 1
    f_dc = np.tile(np.array(rgb_dc, np.float32), (n, 1))
    return scenes.GaussianCloud(np.array(xyz, np.float32), f_dc, z(n, 45), np.full(n, logit, np.float32),
                                np.full((n, 3), log_scale, np.float32), rots)


def test_empty_cloud_renders_black():
    from gsr_b200 import render as R

    cloud = _one(np.zeros((0, 3)), 0.0, 0.0)
    cam = camera.camera_for(scenes.Intrinsics.synthetic(70, 50), scenes.front_view(0), 3)
    img = R.Renderer().render(R.GaussianModel.from_cloud(cloud), cam)
    assert tuple(img.shape) == (50, 70, 3) and not img.any()


@pytest.mark.parametrize("w,h", [(5, 7), (16, 16), (17, 33), (1, 1), (2, 2)])
def test_images_smaller_than_or_straddling_a_tile(w, h):
    cloud = scenes.synth_v1(300, 3, scale_factor=2.0)
    cam = camera.camera_for(scenes.Intrinsics.synthetic(w, h), scenes.front_view(0), 3)
    img, ref, _ = _render_both(cloud, cam)
    assert np.abs(img - ref["image"]).max() < 2e-4
    assert not img[-1].any() and not img[:, -1].any()


def test_gaussian_covering_the_whole_frame_takes_the_unculled_large_rect_path():
    # one opaque splat with a radius of hundreds of pixels (rect far above the 64-tile keep-mask limit) + small ones
    big = _one([[0.0, 0.0, 0.0]], np.log(1.5), 4.0)
    small = scenes.synth_v1(500, 9, scale_factor=0.8)
    cloud = scenes.GaussianCloud(*[np.concatenate([a, b]) for a, b in zip(
        (big.xyz, big.f_dc, big.f_rest, big.opacity, big.scales, big.rots),
        (small.xyz, small.f_dc, small.f_rest, small.opacity, small.scales, small.rots))])
    cam = camera.camera_for(scenes.Intrinsics.synthetic(400, 300), scenes.front_view(0), 3)
    img, ref, r = _render_both(cloud, cam)
    assert int(ref["tile_rect"][0, 2] - ref["tile_rect"][0, 0]) * int(ref["tile_rect"][0, 3] - ref["tile_rect"][0, 1]) > 64
    assert np.abs(img - ref["image"]).max() < 2e-4
    assert img.max() > 0.5


def test_long_thin_splats_lose_no_pixel_to_the_tile_culling():
    """for needle-like splats lying diagonally det = c00 c11 - c01^2 cancels in f32 (here by 5-6 digits), so the
    ellipse of the ROUNDED conic - the one the reference's blend evaluates - differs measurably from the covariance's.
    The exact tile culling must follow the rounded conic (or keep the whole rect): one lost pixel of a front splat
    is worth > 1e-3 in the image, and the image is compared at 2e-4.  Few splats, so transmittance stays high."""
    rng = np.random.default_rng(123)
    n = 240
    xyz = rng.uniform(-1.2, 1.2, (n, 3)).astype(np.float32)
    scales = np.tile(np.log(np.array([1.5, 0.002, 0.002], np.float32)), (n, 1))
    scales[n // 2:, 0] = np.log(0.4)                      # a shorter kind: rects under the 64-tile mask limit too
    rots = rng.normal(size=(n, 4)).astype(np.float32)     # every orientation, most of them diagonal on screen
    z = lambda *s: np.zeros(s, np.float32)
    f_dc = rng.uniform(0.5, 1.7, (n, 3)).astype(np.float32)
    cloud = scenes.GaussianCloud(xyz, f_dc, z(n, 45), np.full(n, 2.0, np.float32), scales.astype(np.float32), rots)
    for w, h in ((640, 400), (333, 217)):
        cam = camera.camera_for(scenes.Intrinsics.synthetic(w, h), scenes.front_view(0), 3)
        img, ref, r = _render_both(cloud, cam)
        c2 = ref["cov2d"].reshape(-1, 4).astype(np.float64)
        cancel = (c2[:, 0] * c2[:, 3]) / np.maximum(np.abs(c2[:, 0] * c2[:, 3] - c2[:, 1] * c2[:, 2]), 1e-30)
        assert (cancel > 1e3).sum() > 20                  # the scenario is really there: det cancels by > 3 digits
        assert img.max() > 0.3
        assert np.abs(img - ref["image"]).max() < 2e-4
        assert r.frame_info()["n_visible"] == int((ref["skip"] == 0).sum())


@pytest.mark.parametrize("degree", [0, 1, 2, 3])
def test_every_sh_degree(degree):
    cloud = scenes.synth_v1(2000, 77, scale_factor=0.6)
    intr = scenes.Intrinsics.synthetic(120, 80)
    cam = camera.camera_for(intr, scenes.orbit_views(1, total=8, first=1)[0], degree)
    img, ref, _ = _render_both(cloud, cam)
    assert np.abs(img - ref["image"]).max() < 2e-4


def test_invisible_and_degenerate_gaussians_are_harmless():
    cloud = scenes.synth_v1(400, 5, scale_factor=0.8)
    cloud.opacity[:50] = -30.0            # opacity ~ 1e-13: never passes alpha > 1/255
    cloud.xyz[50:100, 2] = -10.0          # behind the camera: culled (z < 0.2)
    cloud.xyz[100:150, 0] = 500.0         # far off screen
    cloud.scales[150:160] = -40.0         # vanishing extent: only the 0.3 low-pass remains
    cloud.rots[160:170] = 0.0             # zero quaternion: normalise() divides by eps, rotation collapses
    cam = camera.camera_for(scenes.Intrinsics.synthetic(96, 64), scenes.front_view(0), 3)
    img, ref, r = _render_both(cloud, cam)
    assert np.isfinite(img).all()
    assert np.abs(img - ref["image"]).max() < 2e-4
    assert r.frame_info()["n_visible"] == int((ref["skip"] == 0).sum())


def test_workspace_grows_when_max_dup_is_too_small():
    cloud = scenes.synth_v1(5000, 21, scale_factor=1.0)
    cam = camera.camera_for(scenes.Intrinsics.synthetic(160, 112), scenes.front_view(0), 3)
    img, ref, r = _render_both(cloud, cam, dup_factor=0.01)  # first attempt overflows, the renderer redraws
    assert r.max_dup >= r.frame_info()["n_dup"] > 0.01 * cloud.n
    assert np.abs(img - ref["image"]).max() < 2e-4


def test_render_is_stream_ordered_and_reusable_across_sizes():
    from gsr_b200 import render as R

    r = R.Renderer()
    s = torch.cuda.Stream()
    for (n, w, h) in ((3000, 200, 120), (800, 64, 48), (3000, 200, 120)):
        cloud = scenes.synth_v1(n, n, scale_factor=0.7)
        cam = camera.camera_for(scenes.Intrinsics.synthetic(w, h), scenes.front_view(0), 3)
        with torch.cuda.stream(s):
            img = r.render(R.GaussianModel.from_cloud(cloud), cam)
        s.synchronize()
        assert np.abs(img.cpu().numpy() - c_oracle.render(cloud, cam)["image"]).max() < 2e-4


def test_ply_device_loader_matches_column_loader(tmp_path):
    """GaussianModel.from_ply (raw vertex block unpacked on the GPU) == the per-column host route, also for a file
    whose properties come in another order with extra columns."""
    import os

    from gsr_b200 import render as R
    from gsr_b200.ply import GS_PROPERTIES, write_gaussian_ply

    cloud = scenes.synth_v1(1234, 8)
    ref = R.GaussianModel.from_cloud(cloud)
    path = str(tmp_path / "a" / "point_cloud.ply")
    write_gaussian_ply(path, cloud.ply_table())
    m = R.GaussianModel.from_ply(path)
    for a, b in ((m.xyz, ref.xyz), (m.log_scales, ref.log_scales), (m.rots, ref.rots), (m.opacity, ref.opacity), (m.sh, ref.sh)):
        assert torch.equal(a, b)
    assert m.sh.stride() == (1, 1234, 16 * 1234)  # planar [3,16,N] storage
    # permuted property order + two foreign float columns
    rng = np.random.default_rng(0)
    names = list(GS_PROPERTIES) + ["extra_a", "extra_b"]
    perm = rng.permutation(len(names))
    table = np.concatenate([cloud.ply_table(), rng.normal(size=(cloud.n, 2)).astype(np.float32)], axis=1)[:, perm]
    path2 = str(tmp_path / "b.ply")
    head = ["ply", "format binary_little_endian 1.0", f"element vertex {cloud.n}"]
    head += [f"property float {names[j]}" for j in perm] + ["end_header", ""]
    with open(path2, "wb") as fh:
        fh.write("\n".join(head).encode("ascii")); fh.write(np.ascontiguousarray(table, dtype="<f4").tobytes())
    m2 = R.GaussianModel.from_ply(path2)
    for a, b in ((m2.xyz, ref.xyz), (m2.rots, ref.rots), (m2.opacity, ref.opacity), (m2.sh, ref.sh)):
        assert torch.equal(a, b)


_PX_SCRIPT = r"""
import os, sys, numpy as np, torch
root = sys.argv[1]
sys.path[:0] = [os.path.join(root, "torch-gaussian-splatting-rasterizer_b200"), root]
from gsr_b200 import camera, scenes, render as R
cloud = scenes.synth_v1(40_000, 11, scale_factor=0.5)
intr = scenes.Intrinsics.synthetic(403, 277)
cam = camera.camera_for(intr, scenes.orbit_views(1, total=7, first=3)[0], 3)
img = R.Renderer().render(R.GaussianModel.from_cloud(cloud), cam)
np.save(sys.argv[2], img.cpu().numpy())
"""


def test_packed_fp32_blend_is_bit_identical_to_the_scalar_kernel(tmp_path):
    """The 4-pixel kernel evaluates `power` with Blackwell's packed FP32 instructions (FFMA2 / FMUL2 / FADD2); the
    2-pixel instantiation is the scalar code.  Same pixels, same Gaussians, same order: the images must be equal bit
    for bit (a contracted multiply-add in the packed path would show up here)."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = {}
    for px in ("2", "4"):
        path = str(tmp_path / f"px{px}.npy")
        env = dict(os.environ, GSR_BLEND_PX=px)
        subprocess.run([sys.executable, "-c", _PX_SCRIPT, root, path], check=True, env=env, timeout=600)
        out[px] = np.load(path)
    assert out["4"].max() > 0.1
    assert np.array_equal(out["2"].view(np.uint32), out["4"].view(np.uint32))


def test_blend_staging_variants_are_bit_identical():
    """cp.async per CTA, cp.async.bulk + mbarrier (TMA unit), cp.async per warp: same records in the same order, so the
    same image bit for bit - also with contributor tracking and a rank window."""
    from gsr_b200 import _cabi, render as R

    cloud = scenes.synth_v1(60_000, 17, scale_factor=0.45)
    intr = scenes.Intrinsics.synthetic(389, 251)
    cam = camera.camera_for(intr, scenes.orbit_views(1, total=9, first=2)[0], 3)
    model = R.GaussianModel.from_cloud(cloud)
    pre = R.preprocess(model, cam)
    n_pairs = int((pre["tiles_touched"] & 0x7FFFFFFF).sum().item())
    binned = R.bin_tiles(pre, cam.width, cam.height, n_pairs + 1024)
    outs = []
    try:
        for stage in (0, 1, 2):
            _cabi.call("gsr_debug_blend_variant", stage, 0)
            a = R.blend(pre, binned, cam.width, cam.height)                              # tracked
            b = R.blend(pre, binned, cam.width, cam.height, track=False, last_rank=777)   # rank window
            outs.append((a["image"].clone(), a["final_T"].clone(), a["n_contrib"].clone(), b["image"].clone()))
    finally:
        _cabi.call("gsr_debug_blend_variant", -1, 0)
    assert outs[0][0].max() > 0.1 and not torch.equal(outs[0][0], outs[0][3])
    for o in outs[1:]:
        for x, y in zip(o, outs[0]):
            assert torch.equal(x, y)


def test_frame_above_65536_tiles_takes_three_tile_sort_passes():
    """An 8K frame (7680x4320 = 480 x 270 = 129600 tiles, 17 tile-id bits): the pair sort needs a third 8-bit pass and a
    third digit histogram; image against the oracle as at any other size."""
    cloud = scenes.synth_v1(30_000, 29, scale_factor=0.1)
    cam = camera.camera_for(scenes.Intrinsics.synthetic(7680, 4320), scenes.front_view(0), 3)
    img, ref, r = _render_both(cloud, cam)
    assert np.abs(img - ref["image"]).max() < 2e-4
    assert img.max() > 0.2 and r.frame_info()["n_visible"] == int((ref["skip"] == 0).sum())
