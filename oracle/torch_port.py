"""Differentiable torch-CPU port of the reference render path.  TEST INFRASTRUCTURE ONLY.

Two jobs, both as a checker / reported baseline, never as the product:
  * ``bench.py``'s ``cpu_baseline`` and ``--impl reference`` legs time :func:`render` - it keeps the
    reference's execution model (vectorised per-Gaussian set-up, then ONE eager-torch blend step per
    Gaussian in global depth order, /root/reference/rasterize.py:436-446) so its cost per Gaussian is
    of the reference's kind (PINNED: tests/test_oracle_cpu.py compares it with the goldens).
  * ``tests/`` use ``torch.autograd`` through it as the gradient oracle for the backward kernels on
    the GPU box, where the reference checkout does not exist (pinned against tests/golden/grad_small.npz,
    which was produced by autograd through the reference's own functions).

Differences in form from the reference (results identical to float32 rounding): the blend step works on
a rectangular slice of the screen with broadcasting and ``torch.where`` instead of index lists.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch

BLOCK = 16
SH_0 = 0.28209479177387814
SH_C1 = 0.4886025119029199
SH_C2 = (1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396)
SH_C3 = (-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
         1.445305721320277, -0.5900435899266435)


def _cam_tensors(cam):
    w2c_t = torch.tensor(list(cam.w2c_t), dtype=torch.float32).reshape(4, 4)
    full_t = torch.tensor(list(cam.full_t), dtype=torch.float32).reshape(4, 4)
    centre = torch.tensor(list(cam.cam_center), dtype=torch.float32)
    return w2c_t, full_t, centre


def sh_colour(xyz, sh, centre, degree):
    """spherical_harmonics.py:36-71.  sh is [N,16,3]."""
    d = xyz - centre
    d = d / torch.linalg.vector_norm(d, dim=1, keepdim=True)
    x, y, z = d[:, 0:1], d[:, 1:2], d[:, 2:3]
    col = sh[:, 0, :] * SH_0
    if degree > 0:
        col = col + (-SH_C1 * y * sh[:, 1, :] + SH_C1 * z * sh[:, 2, :] - SH_C1 * x * sh[:, 3, :])
    if degree > 1:
        xx, yy, zz = x * x, y * y, z * z
        col = col + (SH_C2[0] * x * y * sh[:, 4, :] + SH_C2[1] * y * z * sh[:, 5, :]
                     + SH_C2[2] * (2 * zz - xx - yy) * sh[:, 6, :] + SH_C2[3] * x * z * sh[:, 7, :]
                     + SH_C2[4] * (xx - yy) * sh[:, 8, :])
    if degree > 2:
        col = col + (SH_C3[0] * y * (3 * xx - yy) * sh[:, 9, :] + SH_C3[1] * x * y * z * sh[:, 10, :]
                     + SH_C3[2] * y * (4 * zz - xx - yy) * sh[:, 11, :]
                     + SH_C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * sh[:, 12, :]
                     + SH_C3[4] * x * (4 * zz - xx - yy) * sh[:, 13, :] + SH_C3[5] * z * (xx - yy) * sh[:, 14, :]
                     + SH_C3[6] * x * (xx - 3 * yy) * sh[:, 15, :])
    return torch.clamp(col + 0.5, 0, 1)


def covariance3d(log_scales, rots):
    """rasterize.py:97-120 on [N,3] log-scales and [N,4] quaternions."""
    s = torch.exp(log_scales)
    q = torch.nn.functional.normalize(rots, p=2.0, dim=1)
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = torch.stack([
        1 - 2 * y * y - 2 * z * z, 2 * x * y - 2 * z * w, 2 * x * z + 2 * y * w,
        2 * x * y + 2 * z * w, 1 - 2 * x * x - 2 * z * z, 2 * y * z - 2 * x * w,
        2 * x * z - 2 * y * w, 2 * y * z + 2 * x * w, 1 - 2 * x * x - 2 * y * y], dim=1).reshape(-1, 3, 3)
    M = R * s[:, None, :]
    return M @ M.transpose(1, 2)


def preprocess(xyz, log_scales, rots, opacity_logit, sh, cam) -> Dict[str, torch.Tensor]:
    """Everything before the blend loop (rasterize.py:354-425), vectorised over Gaussians."""
    W, H = cam.width, cam.height
    w2c_t, full_t, centre = _cam_tensors(cam)
    cov3d = covariance3d(log_scales, rots)
    opacity = torch.sigmoid(opacity_logit)
    rgb = sh_colour(xyz, sh, centre, cam.sh_degree)
    cam_p = xyz @ w2c_t[:3, :3] + w2c_t[3, :3]
    clip = xyz @ full_t[:3, :] + full_t[3, :]
    z = cam_p[:, 2]
    culled = z < 0.2
    clip = torch.where(culled[:, None], torch.zeros_like(clip), clip)
    pw = 1.0 / (clip[:, 3] + 0.0000001)
    ndc = clip[:, :2] * pw[:, None]
    mean2d = ((ndc + 1.0) * torch.tensor([W, H]) - 1.0) / 2
    # EWA projection, rasterize.py:201-252
    fx, fy, limx, limy = cam.focal_x, cam.focal_y, cam.limx, cam.limy
    tx = torch.clamp(cam_p[:, 0] / z, -limx, limx) * z
    ty = torch.clamp(cam_p[:, 1] / z, -limy, limy) * z
    zero = torch.zeros_like(z)
    J = torch.stack([fx / z, zero, -(fx * tx) / (z * z), zero, fy / z, -(fy * ty) / (z * z), zero, zero, zero],
                    dim=1).reshape(-1, 3, 3)
    Rc = w2c_t[:3, :3].T
    T = J @ Rc
    iu = torch.triu(torch.ones(3, 3, dtype=torch.bool))
    vrk = torch.where(iu, cov3d, cov3d.transpose(1, 2))
    cov2d = (T @ vrk @ T.transpose(1, 2))[:, :2, :2] + torch.tensor([[0.3, 0.0], [0.0, 0.3]])
    cov2d = torch.where(culled[:, None, None], torch.zeros_like(cov2d), cov2d)
    det = cov2d[:, 0, 0] * cov2d[:, 1, 1] - cov2d[:, 1, 0] * cov2d[:, 0, 1]
    tr = cov2d[:, 0, 0] + cov2d[:, 1, 1]
    lam = tr / 2.0 + torch.sqrt(torch.clamp_min(tr * tr / 4.0 - det, 0.1))
    radius = torch.ceil(3 * torch.sqrt(lam)).detach()
    m = mean2d.detach()
    rect = torch.stack([torch.clamp((m[:, 0] - radius) / BLOCK, 0, W - 1), torch.clamp((m[:, 1] - radius) / BLOCK, 0, H - 1),
                        torch.clamp((m[:, 0] + (radius + BLOCK - 1)) / BLOCK, 0, W - 1),
                        torch.clamp((m[:, 1] + (radius + BLOCK - 1)) / BLOCK, 0, H - 1)], dim=1).floor().to(torch.int64)
    det_inv = torch.where(det == 0, torch.zeros_like(det), 1 / det)
    conic = torch.stack([cov2d[:, 1, 1] * det_inv, cov2d[:, 0, 0] * det_inv, -cov2d[:, 0, 1] * det_inv], dim=1)
    hi = torch.tensor([W - 1, H - 1, W - 1, H - 1])
    bbox = torch.minimum(torch.clamp_min(rect * BLOCK, 0), hi)
    area = (bbox[:, 2] - bbox[:, 0]) * (bbox[:, 3] - bbox[:, 1])
    skip = (area == 0) | (conic.detach() == 0).any(dim=1)
    order = torch.sort(z.detach(), stable=True).indices
    return dict(cov3d=cov3d, opacity=opacity, rgb=rgb, cam_means=cam_p, depth=z, mean2d=mean2d, cov2d=cov2d,
                tile_rect=rect, conic=conic, pixel_bbox=bbox, skip=skip, order=order)


def blend_step(g: int, pre, screen, T, functional: bool):
    """rasterize.py:255-305 on the rectangular slice of the x-major screen covered by Gaussian g."""
    x0, y0, x1, y1 = pre["pixel_bbox"][g].tolist()
    a, c, b = pre["conic"][g]
    dx = (pre["mean2d"][g, 0] - torch.arange(x0, x1))[:, None]
    dy = (pre["mean2d"][g, 1] - torch.arange(y0, y1))[None, :]
    power = -0.5 * (a * dx * dx + c * dy * dy) - b * dx * dy
    alpha = torch.clamp_max(pre["opacity"][g] * torch.exp(power), 0.99)
    keep = (alpha > 1 / 255) & (power <= 0)
    if functional:  # autograd-friendly: dense update of the slice on fresh copies
        alpha = torch.where(keep, alpha, torch.zeros_like(alpha))
        t_old = T[x0:x1, y0:y1]
        add = (alpha[:, :, None] * pre["rgb"][g]) * t_old[:, :, None]
        screen = screen.clone(); T = T.clone()
        screen[x0:x1, y0:y1] = screen[x0:x1, y0:y1] + add
        T[x0:x1, y0:y1] = t_old * (1 - alpha)
        return screen, T
    # timing path: touch only the pixels that pass the test, in place, like the reference's masked update
    ix, iy = torch.nonzero(keep, as_tuple=True)
    ix += x0; iy += y0
    av = alpha[keep]
    tv = T[ix, iy]
    screen[ix, iy] += (av[:, None] * pre["rgb"][g]) * tv[:, None]
    T[ix, iy] = tv * (1 - av)
    return screen, T


def render(xyz, log_scales, rots, opacity_logit, sh, cam, max_blended: Optional[int] = None,
           differentiable: bool = False) -> Dict[str, torch.Tensor]:
    """Full port render.  ``max_blended`` stops after that many *blended* Gaussians (bounded timing samples)."""
    ctx = torch.enable_grad() if differentiable else torch.no_grad()
    with ctx:
        pre = preprocess(xyz, log_scales, rots, opacity_logit, sh, cam)
        W, H = cam.width, cam.height
        screen, T = torch.zeros((W, H, 3)), torch.ones((W, H))
        blended = 0
        skip = pre["skip"].tolist()
        for g in pre["order"].tolist():
            if skip[g]:
                continue
            screen, T = blend_step(g, pre, screen, T, differentiable)
            blended += 1
            if max_blended is not None and blended >= max_blended:
                break
        pre.update(image=screen.transpose(0, 1), final_T=T.transpose(0, 1), n_blended=blended,
                   n_blendable=int((~pre["skip"]).sum()))
    return pre


def cloud_tensors(cloud, requires_grad: bool = False):
    mk = lambda a: torch.tensor(np.ascontiguousarray(a, dtype=np.float32), requires_grad=requires_grad)
    sh = np.ascontiguousarray(cloud.sh_planar().transpose(2, 1, 0))
    return mk(cloud.xyz), mk(cloud.scales), mk(cloud.rots), mk(cloud.opacity), mk(sh)


def gradients(cloud, cam, grad_image: np.ndarray) -> Dict[str, np.ndarray]:
    """d sum(image*G) / d(parameters) by autograd through the port."""
    xyz, ls, rot, op, sh = cloud_tensors(cloud, requires_grad=True)
    out = render(xyz, ls, rot, op, sh, cam, differentiable=True)
    (out["image"] * torch.tensor(grad_image)).sum().backward()
    return dict(image=out["image"].detach().numpy().copy(), d_xyz=xyz.grad.numpy(), d_scales=ls.grad.numpy(),
                d_rots=rot.grad.numpy(), d_opacity=op.grad.numpy(), d_sh=sh.grad.numpy())
