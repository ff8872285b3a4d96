"""TEST INFRASTRUCTURE ONLY - drives the UNMODIFIED reference on CPU to produce golden vectors.

This file never ships in the product path and is importable only where the reference
checkout exists (this build container: ``/root/reference``; the GPU box has no copy).
It is used by ``tests/golden/make_golden.py`` to pin ``oracle/gsr_oracle.c`` and
``oracle/torch_port.py`` against the real thing.

How: the reference imports two packages that are not installed here (``plyfile``,
``matplotlib``).  Tiny stand-ins are placed in ``sys.modules`` *before* importing
``rasterize``; ``plt.imshow`` captures the rendered ``[H,W,3]`` tensor
(/root/reference/rasterize.py:471).  Intermediates are captured by wrapping the
reference's own functions (``compute_2d_covariance``, ``compute_covering_bbox``,
``rasterize_gaussian``, ``torch.sort``) - the reference's glue code runs as is.

Tie order: the reference sorts depths with ``torch.sort`` (unstable,
rasterize.py:425).  ``stable=True`` is forced here because north_star defines ties
"by Gaussian index" (SURVEY.md Q13).
"""
from __future__ import annotations

import contextlib
import os
import sys
import types
from typing import Dict, Optional

import numpy as np
import torch

def _find_reference() -> str:
    """GSR_REFERENCE_DIR, else the read-only checkout of the build container, else the staged copy that travels to the
    GPU box for the benchmark's reference arm (git-ignored baseline/_ref, written by __graft_entry__.build())."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for cand in (os.environ.get("GSR_REFERENCE_DIR"), "/root/reference", os.path.join(root, "baseline", "_ref")):
        if cand and os.path.isfile(os.path.join(cand, "rasterize.py")):
            return cand
    return "/root/reference"


REF_DIR = _find_reference()
_PKG_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                        "torch-gaussian-splatting-rasterizer_b200")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "rasterize.py"))


_captured_images = []
_ref_modules: Dict[str, types.ModuleType] = {}


def _install_stubs() -> None:
    if _PKG_DIR not in sys.path:
        sys.path.append(_PKG_DIR)
    from gsr_b200 import ply as our_ply

    ply = types.ModuleType("plyfile")
    ply.PlyData, ply.PlyElement = our_ply.PlyData, our_ply.PlyElement
    mpl = types.ModuleType("matplotlib")
    mimg = types.ModuleType("matplotlib.image")
    mplt = types.ModuleType("matplotlib.pyplot")
    mimg.imread = lambda p: np.zeros((2, 2, 3), dtype=np.float32)
    for name in ("figure", "subplot", "title", "show"):
        setattr(mplt, name, lambda *a, **k: None)
    mplt.imshow = lambda x, *a, **k: _captured_images.append(x)
    mpl.image, mpl.pyplot = mimg, mplt
    sys.modules.update({"plyfile": ply, "matplotlib": mpl, "matplotlib.image": mimg, "matplotlib.pyplot": mplt})


def load_reference() -> Dict[str, types.ModuleType]:
    """Import the reference's modules under private names so they never shadow the drop-in ones."""
    if _ref_modules:
        return _ref_modules
    if not reference_available():
        raise RuntimeError(f"reference checkout not found at {REF_DIR}")
    _install_stubs()
    saved = {k: sys.modules.get(k) for k in ("rasterize", "spherical_harmonics", "utils", "data_reader")}
    for k in saved:
        sys.modules.pop(k, None)
    sys.path.insert(0, REF_DIR)
    try:
        import data_reader, spherical_harmonics, utils, rasterize  # noqa: E401  (the reference's, from REF_DIR)
        _ref_modules.update(rasterize=rasterize, spherical_harmonics=spherical_harmonics, utils=utils,
                            data_reader=data_reader)
    finally:
        sys.path.remove(REF_DIR)
        for k, v in saved.items():
            sys.modules.pop(k, None)
            if v is not None:
                sys.modules[k] = v
    return _ref_modules


@contextlib.contextmanager
def _stable_sort():
    orig = torch.sort

    def stable(x, *a, **k):
        k["stable"] = True
        return orig(x, *a, **k)

    torch.sort = stable
    try:
        yield
    finally:
        torch.sort = orig


class _StopAfter(Exception):
    """Raised by the timing wrapper to leave the reference's loop after a bounded number of blended Gaussians."""


def time_reference(scene_dir: str, scene_index: int = 0, scale_factor: int = 2, max_blended: Optional[int] = None,
                   threads: Optional[int] = None) -> Dict[str, float]:
    """Wall-clock of the unmodified reference's ``run_rasterization`` on CPU, optionally left after `max_blended`
    calls of its own ``rasterize_gaussian`` (the loop is the whole cost and is linear in the number of blended
    Gaussians): returns seconds before the loop (load + vectorised set-up), seconds inside it, Gaussians blended and
    how many the loop would blend in total."""
    import time

    ref = load_reference()["rasterize"]
    orig_blend, orig_sort, orig_set_threads = ref.rasterize_gaussian, torch.sort, torch.set_num_threads
    st = dict(t_first=None, blended=0, blendable=None)

    def sort_wrap(x, *a, **k):
        k["stable"] = True
        return orig_sort(x, *a, **k)

    def blend_wrap(gaussian_index, bboxes, screen, screen_means, sigmas, rgb, opacity_buffer, opacity):
        if st["t_first"] is None:
            st["t_first"] = time.perf_counter()
            area = (bboxes[:, 2] - bboxes[:, 0]) * (bboxes[:, 3] - bboxes[:, 1])
            st["blendable"] = int(((area != 0) & ~(sigmas == 0).any(dim=1)).sum())  # the loop's own skip rule (:441)
        out = orig_blend(gaussian_index, bboxes, screen, screen_means, sigmas, rgb, opacity_buffer, opacity)
        st["blended"] += 1
        if max_blended is not None and st["blended"] >= max_blended:
            raise _StopAfter()
        return out

    _captured_images.clear()
    t0 = time.perf_counter()
    try:
        ref.rasterize_gaussian = blend_wrap
        torch.sort = sort_wrap
        if threads is not None:
            torch.set_num_threads = lambda n: orig_set_threads(threads)
        try:
            ref.run_rasterization.callback(input_dir=scene_dir, trained_model_path=scene_dir, output_path="",
                                           scene_index=scene_index, scale_factor=scale_factor, generate_video=False)
        except _StopAfter:
            pass
    finally:
        ref.rasterize_gaussian, torch.sort, torch.set_num_threads = orig_blend, orig_sort, orig_set_threads
    t1 = time.perf_counter()
    t_first = st["t_first"] if st["t_first"] is not None else t1
    return dict(setup_s=t_first - t0, loop_s=t1 - t_first, blended=st["blended"], blendable=st["blendable"] or 0,
                threads=torch.get_num_threads())


def run_reference(scene_dir: str, scene_index: int = 0, scale_factor: int = 2, stable: bool = True,
                  capture: bool = True, threads: Optional[int] = None) -> Dict[str, np.ndarray]:
    """Run ``run_rasterization.callback`` of the unmodified reference; return image + captured intermediates."""
    ref = load_reference()["rasterize"]
    cap: Dict[str, np.ndarray] = {}
    originals = {n: getattr(ref, n) for n in ("compute_2d_covariance", "compute_covering_bbox", "rasterize_gaussian")}
    orig_sort = torch.sort
    orig_set_threads = torch.set_num_threads

    def cov2d_wrap(cov_matrices, camera_space_points, *a, **k):
        out = originals["compute_2d_covariance"](cov_matrices, camera_space_points, *a, **k)
        cap["cov3d"] = cov_matrices.detach().clone().numpy()
        cap["cam_means"] = camera_space_points.detach().clone().numpy()
        cap["cov2d_unculled"] = out.detach().clone().numpy()
        return out

    def bbox_wrap(screen_means, projected_covariances, width, height):
        out = originals["compute_covering_bbox"](screen_means, projected_covariances, width, height)
        cap["mean2d"] = screen_means.detach().clone().numpy()
        cap["cov2d"] = projected_covariances.detach().clone().numpy()
        cap["tile_rect"] = out.detach().clone().numpy()
        return out

    def blend_wrap(gaussian_index, bboxes, screen, screen_means, sigmas, rgb, opacity_buffer, opacity):
        if "conic" not in cap:
            cap["pixel_bbox"] = bboxes.detach().clone().numpy()
            cap["conic"] = sigmas.detach().clone().numpy()
            cap["rgb"] = rgb.detach().clone().numpy()
            cap["opacity"] = opacity.detach().clone().numpy()
        cap["n_blended"] = cap.get("n_blended", 0) + 1
        out = originals["rasterize_gaussian"](gaussian_index, bboxes, screen, screen_means, sigmas, rgb,
                                              opacity_buffer, opacity)
        cap["_T"] = out[1]
        return out

    def sort_wrap(x, *a, **k):
        if stable:
            k["stable"] = True
        out = orig_sort(x, *a, **k)
        if x.dim() == 1 and "depth" not in cap:
            cap["depth"] = x.detach().clone().numpy()
            cap["order"] = out.indices.detach().clone().numpy()
        return out

    _captured_images.clear()
    try:
        if capture:
            ref.compute_2d_covariance, ref.compute_covering_bbox = cov2d_wrap, bbox_wrap
            ref.rasterize_gaussian = blend_wrap
        torch.sort = sort_wrap
        if threads is not None:
            torch.set_num_threads = lambda n: orig_set_threads(threads)
        ref.run_rasterization.callback(input_dir=scene_dir, trained_model_path=scene_dir, output_path="",
                                       scene_index=scene_index, scale_factor=scale_factor, generate_video=False)
    finally:
        for n, f in originals.items():
            setattr(ref, n, f)
        torch.sort = orig_sort
        torch.set_num_threads = orig_set_threads
    img = _captured_images[0]
    cap["image"] = np.ascontiguousarray(img.detach().numpy() if isinstance(img, torch.Tensor) else np.asarray(img))
    if "_T" in cap:
        cap["final_T"] = np.ascontiguousarray(cap.pop("_T").detach().numpy().T)  # [W,H] -> [H,W]
    return cap


def reference_sh_to_rgb(xyz, sh, w2c_t, degree):
    return load_reference()["spherical_harmonics"].sh_to_rgb(xyz, sh, w2c_t, degree=degree)


def reference_autograd(cloud, intr, view, degree: int, grad_seed: int = 0) -> Dict[str, np.ndarray]:
    """Gradients of ``sum(image * G)`` (G ~ U(0,1) seeded) through the reference's own functions.

    The reference has no backward; torch.autograd flows through its helper functions when
    leaf tensors replace the PLY-derived ones.  Only the glue that the
    reference inlines in its click command is re-stated here: Sigma-from-tensors
    (rasterize.py:112-120) and lines 371-425.
    """
    ref = load_reference()["rasterize"]
    sh_fn = load_reference()["spherical_harmonics"].sh_to_rgb
    import math

    leaf = lambda a: torch.tensor(np.asarray(a, dtype=np.float32), requires_grad=True)
    means, log_s, quat, op_logit = leaf(cloud.xyz), leaf(cloud.scales), leaf(cloud.rots), leaf(cloud.opacity)
    sh = leaf(np.ascontiguousarray(cloud.sh_planar().transpose(2, 1, 0)))  # [N,16,3]
    W, H = intr.width, intr.height
    focals = np.array([intr.fx, intr.fy])
    fov_x = 2 * np.arctan(intr.cam_width / (2 * intr.fx)); fov_y = 2 * np.arctan(intr.cam_height / (2 * intr.fy))
    tan_x, tan_y = math.tan(fov_x * 0.5), math.tan(fov_y * 0.5)

    scales = torch.exp(log_s.T)
    uq = torch.nn.functional.normalize(quat.T, p=2.0, dim=0)
    R = ref.quaternion_to_rotation_matrix(uq).permute(2, 0, 1)
    M = R @ torch.diag_embed(scales.T)
    cov3d = M @ M.permute(0, 2, 1)
    opacity = torch.sigmoid(op_logit)

    w2c = ref.get_world_to_camera_matrix(torch.tensor(view.qvec), torch.tensor(view.tvec)).transpose(0, 1)
    proj = ref.get_projection_matrix(fov_x, fov_y).transpose(0, 1)
    full = (w2c.unsqueeze(0).bmm(proj.unsqueeze(0))).squeeze(0)
    rgb = sh_fn(means, sh, w2c, degree=degree)
    cam = ref.project_to_camera_space(means, w2c)
    pts = means @ full[:3, :] + full[-1, :]
    cull = cam[:, 2] < 0.2
    pts = torch.where(cull[:, None], torch.zeros_like(pts), pts)
    p_w = 1.0 / (pts[:, -1] + 0.0000001)
    p_proj = pts[:, :-1] * p_w[:, None]
    cov2d = ref.compute_2d_covariance(cov3d, cam, tan_x, tan_y, focals, w2c)
    cov2d = torch.where(cull[:, None, None], torch.zeros_like(cov2d), cov2d)
    mean2d = ((p_proj[:, :2] + 1.0) * torch.tensor([W, H]) - 1.0) / 2
    rect = ref.compute_covering_bbox(mean2d, cov2d, W, H)
    det = cov2d[:, 0, 0] * cov2d[:, 1, 1] - cov2d[:, 1, 0] * cov2d[:, 0, 1]
    det_inv = torch.where(det == 0, torch.zeros_like(det), 1 / det)
    conic = torch.stack([cov2d[:, 1, 1] * det_inv, cov2d[:, 0, 0] * det_inv, -cov2d[:, 0, 1] * det_inv], dim=-1)
    B = ref.BLOCK_SIZE
    bb = torch.stack([torch.clamp(rect[:, 0] * B, 0, W - 1), torch.clamp(rect[:, 1] * B, 0, H - 1),
                      torch.clamp(rect[:, 2] * B, 0, W - 1), torch.clamp(rect[:, 3] * B, 0, H - 1)], dim=-1)
    area = (bb[:, 2] - bb[:, 0]) * (bb[:, 3] - bb[:, 1])
    order = torch.sort(cam[:, 2].detach(), stable=True).indices
    screen = torch.zeros((W, H, 3)); T = torch.ones((W, H))
    for g in order:
        if area[g] == 0 or torch.any(conic[g] == 0):
            continue
        # rasterize_gaussian mutates in place; clone so autograd keeps every version
        screen, T = ref.rasterize_gaussian(g, bb, screen.clone(), mean2d, conic, rgb, T.clone(), opacity)
    image = screen.transpose(1, 0)
    G = torch.tensor(np.random.default_rng(grad_seed).uniform(0, 1, (H, W, 3)).astype(np.float32))
    (image * G).sum().backward()
    return dict(image=image.detach().numpy().copy(), grad_image=G.numpy(),
                d_xyz=means.grad.numpy(), d_scales=log_s.grad.numpy(), d_rots=quat.grad.numpy(),
                d_opacity=op_logit.grad.numpy(), d_sh=sh.grad.numpy())
