"""Spherical-harmonics colour evaluation (drop-in for the reference's ``spherical_harmonics``).

``sh_to_rgb`` runs the sm_100a kernel behind ``gsr_sh_fwd`` (include/gsr_b200.h); there is no CPU path.
"""
from __future__ import annotations

import ctypes

import torch

# real SH basis constants in Cartesian form (reference spherical_harmonics.py:4-24)
SH_0 = 0.28209479177387814
SH_C1 = 0.4886025119029199
SH_C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396]
SH_C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
         1.445305721320277, -0.5900435899266435]


def sh_to_rgb(xyz: torch.Tensor, sh: torch.Tensor, world_view_transform: torch.Tensor, degree: int = 0) -> torch.Tensor:
    """View-dependent colour of every Gaussian: ``[N,3]`` in [0,1] (+0.5 offset, clamped).

    ``sh`` is ``[N,16,3]`` with any strides; the camera centre is row 3 of the inverse of the (transposed)
    world-to-camera matrix, computed on the host with the same torch call as the reference.
    """
    from gsr_b200 import _cabi

    _cabi.require_cuda()
    if not 0 <= int(degree) <= 3:
        raise ValueError("degree must be 0..3")
    origin = xyz.device
    centre = world_view_transform.detach().cpu().float().inverse()[3, :3].contiguous()
    xyz_d = xyz.detach().float().cuda()
    sh_d = sh.detach().float().cuda()
    out = torch.empty((xyz_d.shape[0], 3), dtype=torch.float32, device=xyz_d.device)
    centre_c = (ctypes.c_float * 3)(*centre.tolist())
    _cabi.call("gsr_sh_fwd", xyz_d.shape[0], _cabi.ptr(xyz_d), xyz_d.stride(0), xyz_d.stride(1), _cabi.ptr(sh_d),
               sh_d.stride(0), sh_d.stride(1), sh_d.stride(2), centre_c, int(degree), _cabi.ptr(out),
               _cabi.stream_handle())
    return out.to(origin)
