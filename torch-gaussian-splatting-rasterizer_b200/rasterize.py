"""Drop-in for the reference's ``rasterize`` module, backed by libgsr_b200.so (sm_100a).

Same module-level names, argument order and meaning as /root/reference/rasterize.py; what differs is where
the work happens.  Per-Gaussian and per-pixel math runs in CUDA kernels reached through the C ABI
(include/gsr_b200.h); only the O(1) camera set-up stays on the host, as in the reference.  There is no CPU
fallback: without a CUDA device every compute entry point raises ``gsr_b200._cabi.GsrError``.

New compared with the reference: ``render`` / ``render_views`` - the body the reference inlines in its click
command (:347-452) - which ``run_rasterization`` now calls.
"""
from __future__ import annotations

import ctypes
import logging
import math
import os
import shutil
import subprocess
from typing import Optional, Tuple

import click
import numpy as np
import torch

from gsr_b200 import camera as _camera
from gsr_b200.render import GaussianModel, Renderer, render, render_views  # noqa: F401  (public API)
from spherical_harmonics import sh_to_rgb  # noqa: F401
from utils import read_color_components, read_scene

logger = logging.getLogger(__name__)

# constants of the reference (rasterize.py:27-38); the kernels hard-code the same values
Z_FAR = 100.0
Z_NEAR = 0.01
GAUSSIAN_SPREAD = 3
BLOCK_SIZE = 16
MAX_GAUSSIAN_DENSITY = 0.99
MIN_ALPHA = 1 / 255


def _cabi():
    from gsr_b200 import _cabi as m

    m.require_cuda()
    return m


def _dev(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(device="cuda", dtype=torch.float32)


def quaternion_to_rotation_matrix(quaternion: torch.Tensor) -> torch.Tensor:
    """``[4,N]`` (w,x,y,z) -> ``[3,3,N]`` float32, no normalisation (reference :41-56).

    Camera poses (float64, a handful of columns) are evaluated on the host like every other per-view constant;
    per-Gaussian batches go through ``gsr_quat_to_rotmat``.
    """
    if quaternion.dtype == torch.float64 or (not quaternion.is_cuda and quaternion.shape[1] <= 16):
        return _camera.rotation_from_quaternion(quaternion)
    m = _cabi()
    q = _dev(quaternion)
    out = torch.empty((3, 3, q.shape[1]), dtype=torch.float32, device=q.device)
    m.call("gsr_quat_to_rotmat", q.shape[1], m.ptr(q), q.stride(0), q.stride(1), m.ptr(out), m.stream_handle())
    return out.to(quaternion.device)


def get_world_to_camera_matrix(normalized_qvec: torch.Tensor, tvec: torch.Tensor) -> torch.Tensor:
    """4x4 ``[R|t; 0 0 0 1]`` float32 (reference :59-77); callers transpose it (row-vector convention)."""
    return _camera.world_to_camera(normalized_qvec, tvec)


def get_projection_matrix(fov_x: float, fov_y: float) -> torch.Tensor:
    """Perspective matrix from the two fields of view and Z_NEAR / Z_FAR (reference :123-151)."""
    return _camera.perspective(fov_x, fov_y)


def project_to_camera_space(gaussian_means: torch.Tensor, world_to_camera: torch.Tensor) -> torch.Tensor:
    """``means @ M[:3,:3] + M[-1,:3]`` for the transposed world-to-camera matrix (reference :80-86)."""
    m = _cabi()
    x = _dev(gaussian_means)
    mat = (ctypes.c_float * 16)(*world_to_camera.detach().cpu().float().contiguous().reshape(-1).tolist())
    out = torch.empty((x.shape[0], 3), dtype=torch.float32, device=x.device)
    m.call("gsr_affine_fwd", x.shape[0], m.ptr(x), x.stride(0), x.stride(1), mat, 3, m.ptr(out), m.stream_handle())
    return out.to(gaussian_means.device)


def get_covariance_matrix_from_mesh(mesh) -> torch.Tensor:
    """Sigma = R S S^T R^T per Gaussian from the PLY's ``scale_*`` (log) and ``rot_*`` columns: ``[N,3,3]`` on the
    GPU (reference :89-120)."""
    m = _cabi()
    el = mesh.elements[0]
    scales = torch.from_numpy(np.stack([el["scale_0"], el["scale_1"], el["scale_2"]], axis=1).astype(np.float32)).cuda()
    rots = torch.from_numpy(np.stack([el[f"rot_{i}"] for i in range(4)], axis=1).astype(np.float32)).cuda()
    out = torch.empty((scales.shape[0], 3, 3), dtype=torch.float32, device="cuda")
    m.call("gsr_cov3d_fwd", scales.shape[0], m.ptr(scales), 3, 1, m.ptr(rots), 4, 1, m.ptr(out), m.stream_handle())
    return out


def compute_covering_bbox(screen_means: torch.Tensor, projected_covariances: torch.Tensor, width: float,
                          height: float) -> torch.Tensor:
    """Tile rect ``[N,4]`` int64 (tile units) from the eigen radius ``ceil(3 sqrt(lambda_max))`` (reference :154-198)."""
    m = _cabi()
    mu, cov = _dev(screen_means).contiguous(), _dev(projected_covariances).contiguous()
    out = torch.empty((mu.shape[0], 4), dtype=torch.int64, device=mu.device)
    m.call("gsr_bbox_fwd", mu.shape[0], m.ptr(mu), m.ptr(cov), int(width), int(height), m.ptr(out), m.stream_handle())
    return out.to(screen_means.device)


def compute_2d_covariance(cov_matrices, camera_space_points, tan_fov_x, tan_fov_y, focals, world_to_camera) -> torch.Tensor:
    """EWA projection ``(J R) Sigma (J R)^T`` + 0.3 low-pass: ``[N,2,2]`` (reference :201-252; ``focals`` is the
    full-resolution pair, halved inside like the reference does at :216)."""
    m = _cabi()
    cov, pts = _dev(cov_matrices).contiguous(), _dev(camera_space_points).contiguous()
    cam = _camera.CameraStruct()
    cam.w2c_t[:] = world_to_camera.detach().cpu().float().contiguous().reshape(-1).tolist()
    fx, fy = (np.asarray(focals, dtype=np.float64) / 2).tolist()
    cam.focal_x, cam.focal_y = float(np.float32(fx)), float(np.float32(fy))
    cam.limx, cam.limy = float(np.float32(1.3 * tan_fov_x)), float(np.float32(1.3 * tan_fov_y))
    out = torch.empty((cov.shape[0], 2, 2), dtype=torch.float32, device=cov.device)
    m.call("gsr_cov2d_fwd", cov.shape[0], m.ptr(cov), m.ptr(pts), ctypes.byref(cam), m.ptr(out), m.stream_handle())
    return out.to(cov_matrices.device)


def rasterize_gaussian(gaussian_index: int, bboxes: torch.Tensor, screen: torch.Tensor, screen_means: torch.Tensor,
                       sigmas: torch.Tensor, rgb: torch.Tensor, opacity_buffer: torch.Tensor,
                       opacity: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Blend ONE Gaussian into the x-major ``screen[W,H,3]`` / ``opacity_buffer[W,H]`` in place (reference :255-305).

    Kept for API compatibility; ``render`` is the fast path (one launch per frame instead of one per Gaussian).
    """
    m = _cabi()
    if not (screen.is_cuda and opacity_buffer.is_cuda):
        raise m.GsrError(-2, "rasterize_gaussian needs screen and opacity_buffer on the GPU (no CPU fallback)")
    g = int(gaussian_index)
    bb = (ctypes.c_int64 * 4)(*[int(v) for v in bboxes[g].tolist()])
    mu = (ctypes.c_float * 2)(*screen_means[g].tolist())
    con = (ctypes.c_float * 3)(*sigmas[g].tolist())
    col = (ctypes.c_float * 3)(*rgb[g].tolist())
    m.call("gsr_blend_one", bb, mu, con, col, float(opacity[g]), screen.shape[0], screen.shape[1], m.ptr(screen),
           m.ptr(opacity_buffer), m.stream_handle())
    return screen, opacity_buffer


# ----------------------------------------------------------------------------------------------
# scene loading + CLI
# ----------------------------------------------------------------------------------------------
def load_model(trained_model_path: str, device: str = "cuda") -> GaussianModel:
    """``<path>/point_cloud/iteration_30000/point_cloud.ply`` (reference :353) straight into the kernel layout."""
    ply_path = os.path.join(trained_model_path, "point_cloud/iteration_30000/point_cloud.ply")
    logger.info("Fetching trained model from: %s", ply_path)
    return GaussianModel.from_ply(ply_path, device=device)


def camera_from_scene(scene, cam_info, width: int, height: int, degree: int = 3) -> "_camera.CameraStruct":
    """Intrinsics of camera id 1 (full resolution), pose of ``scene`` - the reference's choices at :336-348."""
    fx, fy, _, _ = cam_info[1].params
    return _camera.make_camera(scene.qvec, scene.tvec, fx, fy, cam_info[1].width, cam_info[1].height, width, height,
                               degree)


def _save_png(image: torch.Tensor, path: str) -> None:
    from PIL import Image

    Image.fromarray((image.clamp(0, 1).cpu().numpy() * 255.0).astype(np.uint8)).save(path)


def psnr(image: torch.Tensor, reference: torch.Tensor) -> float:
    """PSNR (dB, peak 1.0) of two ``[H,W,3]`` images in [0,1]; ``inf`` for identical images."""
    mse = torch.mean((image.float().clamp(0, 1) - reference.to(image.device).float().clamp(0, 1)) ** 2).item()
    return float("inf") if mse == 0 else -10.0 * math.log10(mse)


def compare_with_ground_truth(image: torch.Tensor, gt_img_path: str, out_path: Optional[str] = None) -> dict:
    """What the reference shows in a matplotlib window (:468-476: the render above the ground-truth photo), as numbers
    and an optional PNG: PSNR / mean absolute error against the photo (resized to the render if needed), and the two
    images stacked vertically in ``out_path``."""
    from PIL import Image

    H, W = image.shape[:2]
    gt = Image.open(gt_img_path).convert("RGB")
    if gt.size != (W, H):
        gt = gt.resize((W, H), Image.BILINEAR)
    gt_t = torch.from_numpy(np.asarray(gt, dtype=np.float32) / 255.0)
    img = image.detach().float().clamp(0, 1).cpu()
    out = {"psnr_db": psnr(img, gt_t), "mae": float((img - gt_t).abs().mean())}
    if out_path:
        both = torch.cat([img, gt_t], dim=0)
        Image.fromarray((both.numpy() * 255.0).astype(np.uint8)).save(out_path)
    return out


def write_frame_sequence(frames: torch.Tensor, out_dir: str, prefix: str = "frame_", video: bool = False,
                         framerate: int = 20) -> int:
    """``frames[V,H,W,3]`` (e.g. ``gsr_b200.dist.render_orbit``) -> ``<out_dir>/<prefix>0000000.png`` ...; with
    ``video=True`` and ffmpeg on PATH also ``<out_dir>/video_render.mp4`` (same encoder settings as :465)."""
    os.makedirs(out_dir, exist_ok=True)
    for i in range(frames.shape[0]):
        _save_png(frames[i], os.path.join(out_dir, f"{prefix}{str(i).zfill(7)}.png"))
    if video:
        if shutil.which("ffmpeg"):
            H, W = frames.shape[1:3]
            subprocess.run(["ffmpeg", "-y", "-framerate", str(framerate), "-pattern_type", "glob", "-i",
                            os.path.join(out_dir, f"{prefix}*.png"), "-r", "10", "-vcodec", "libx264", "-s",
                            f"{W - W % 2}x{H - H % 2}", "-pix_fmt", "yuv420p",
                            os.path.join(out_dir, "video_render.mp4")], check=True)
        else:
            logger.warning("ffmpeg not found: %d frames written to %s, no video encoded", frames.shape[0], out_dir)
    return int(frames.shape[0])


@click.command()
@click.option("--input_dir", type=str, default="")
@click.option("--trained_model_path", type=str, default="")
@click.option("--output_path", type=str, default="")
@click.option("--scene-index", type=int, default=0)
@click.option("--scale-factor", type=int, default=2)
@click.option("--generate_video", is_flag=True, type=bool, default=False)
def run_rasterization(input_dir: str, trained_model_path, output_path: Optional[str], scene_index: int = 0,
                      scale_factor: int = 2, generate_video: bool = False) -> None:
    """Same options as the reference's command (:308-322).  Writes ``<output_path>/render.png`` (and the
    progressive-build frames / video when asked) instead of opening a matplotlib window."""
    from PIL import Image

    logger.info("Fetching scenes from: %s", input_dir)
    scenes, cam_info = read_scene(path_to_scene=input_dir)
    scene = scenes[scene_index]  # keyed by COLMAP image_id, like the reference (KeyError otherwise)
    width, height = Image.open(os.path.join(input_dir, f"images_{scale_factor}", scene.name)).size
    cam = camera_from_scene(scene, cam_info, width, height, degree=3)
    model = load_model(trained_model_path)
    renderer = Renderer()
    image = renderer.render(model, cam)
    run_rasterization.last_image = image  # handle for callers that drive the command programmatically
    gt_img_path = os.path.join(input_dir, f"images_{scale_factor}", scene.name)
    if output_path:
        os.makedirs(output_path, exist_ok=True)
        _save_png(image, os.path.join(output_path, "render.png"))
        # the reference's closing figure (:468-476), as a file and two numbers
        cmp = compare_with_ground_truth(image, gt_img_path, os.path.join(output_path, "comparison.png"))
        logger.info("render vs %s: PSNR %.2f dB, MAE %.4f", scene.name, cmp["psnr_db"], cmp["mae"])
        run_rasterization.last_comparison = cmp
    if generate_video:
        frames = os.path.join(output_path, "images")
        os.makedirs(frames, exist_ok=True)
        n_blend = renderer.frame_info()["n_visible"]
        step, snap = 0, image  # a scene in which nothing is blended still gets its (black) closing frames
        while step < n_blend:  # reference :448-452: a snapshot after blended Gaussian number step+1
            snap = renderer.render(model, cam, last_rank=step + 1)
            _save_png(snap, os.path.join(frames, f"image_iter_{str(step).zfill(7)}.png"))
            step += 1000
        framerate = 20
        for i in range(1, 2 * framerate + 1):  # two seconds of the last snapshot (:456-460)
            _save_png(snap, os.path.join(frames, f"image_iter_{str(n_blend + 1000 * i + 1).zfill(7)}.png"))
        if shutil.which("ffmpeg"):
            video = os.path.join(output_path, "video_render.mp4")
            if os.path.exists(video):
                os.remove(video)
            subprocess.run(["ffmpeg", "-framerate", str(framerate), "-pattern_type", "glob", "-i",
                            os.path.join(frames, "image_iter_*.png"), "-r", "10", "-vcodec", "libx264", "-s",
                            f"{width - width % 2}x{height - height % 2}", "-pix_fmt", "yuv420p", video], check=True)
        else:
            logger.warning("ffmpeg not found: frames written to %s, no video encoded", frames)


if __name__ == "__main__":
    logging.basicConfig(level=logging.INFO)
    run_rasterization()
