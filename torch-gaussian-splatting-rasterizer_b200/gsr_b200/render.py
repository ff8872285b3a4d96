"""Host side of the B200 render path: device containers, workspace management and the ``render`` calls.

PyTorch is used for device memory, streams and (elsewhere) torch.distributed only; every per-Gaussian or
per-pixel operation happens in libgsr_b200.so through the C ABI (``_cabi``).  The reference inlines this whole
path in its click command (/root/reference/rasterize.py:347-452); ``render`` is that body as a function.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _cabi
from .camera import CameraStruct, make_camera

U32_MAX = 0xFFFFFFFF


# ----------------------------------------------------------------------------------------------
# device container
# ----------------------------------------------------------------------------------------------
@dataclass
class GaussianModel:
    """Raw 3DGS parameters resident in HBM.

    Layout (DESIGN.md "Data layout"): xyz [N,3], log_scales [N,3], rots [N,4], opacity [N] contiguous; SH as the
    planar block [3,16,N] the reference itself builds (utils.py:21-31), exposed as its [N,16,3] view - with it
    every load instruction of a warp in the preprocess kernel reads one full 128-byte line.
    """

    xyz: torch.Tensor
    log_scales: torch.Tensor
    rots: torch.Tensor
    opacity: torch.Tensor
    sh: torch.Tensor  # logical [N,16,3], any strides
    _prepared: Optional[torch.Tensor] = None  # view-independent half, blocked layout (include/gsr_b200.h: gsr_prepare)

    @property
    def n(self) -> int:
        return int(self.xyz.shape[0])

    def prepare(self, force: bool = False) -> torch.Tensor:
        """Compute (once) the prepared model a resident scene is rendered from: Sigma, sigmoid(opacity) and the record's
        logarithms plus xyz / SH in 16-byte blocked rows.  Call ``prepare(force=True)`` after changing a parameter
        tensor in place; the differentiable path never uses it (parameters change every step there)."""
        if self._prepared is None or force:
            _cabi.require_cuda()
            need = ctypes.c_size_t()
            _cabi.call("gsr_prepared_bytes", self.n, ctypes.byref(need))
            buf = self._prepared
            if buf is None or buf.numel() < need.value + 256:
                buf = torch.empty(need.value + 256, dtype=torch.uint8, device=self.device)
            g = self.c_struct()
            al = _align256(buf)
            with torch.cuda.device(self.device):
                _cabi.call("gsr_prepare", ctypes.byref(g), _cabi.ptr(al), al.numel(), _cabi.stream_handle())
            self._prepared = buf
        return _align256(self._prepared)

    @property
    def device(self):
        return self.xyz.device

    @staticmethod
    def from_tensors(means, log_scales, quats, opacity_logits, sh, device="cuda", planar_sh: bool = True,
                     non_blocking: bool = False) -> "GaussianModel":
        def up(t, shape):
            t = torch.as_tensor(t)
            if t.dtype != torch.float32:
                t = t.float()
            assert tuple(t.shape[1:]) == shape, (t.shape, shape)
            return t.to(device, non_blocking=non_blocking)

        sh = torch.as_tensor(sh)
        assert sh.dim() == 3 and tuple(sh.shape[1:]) == (16, 3), "sh must be [N,16,3]"
        sh = sh.float() if sh.dtype != torch.float32 else sh
        sh_d = sh.to(device, non_blocking=non_blocking)  # keeps the strides of the source (planar stays planar)
        if planar_sh and sh_d.stride() != (1, sh_d.shape[0], 16 * sh_d.shape[0]):
            sh_d = sh_d.permute(2, 1, 0).contiguous().permute(2, 1, 0)
        op = torch.as_tensor(opacity_logits).float().reshape(-1).to(device, non_blocking=non_blocking)
        return GaussianModel(up(means, (3,)), up(log_scales, (3,)), up(quats, (4,)), op, sh_d)

    @staticmethod
    def from_cloud(cloud, device="cuda") -> "GaussianModel":
        """From a host ``scenes.GaussianCloud`` (PLY columns)."""
        sh = torch.from_numpy(cloud.sh_planar()).permute(2, 1, 0)  # [N,16,3] view of planar storage
        return GaussianModel.from_tensors(torch.from_numpy(cloud.xyz), torch.from_numpy(cloud.scales),
                                          torch.from_numpy(cloud.rots), torch.from_numpy(cloud.opacity), sh, device)

    @staticmethod
    def from_ply(path: str, device="cuda") -> "GaussianModel":
        """3DGS PLY file -> HBM in the kernel layout: the raw vertex block is uploaded once (pinned staging) and
        de-interleaved on the GPU by ``gsr_ply_unpack``; no per-column work on the host."""
        from .ply import read_vertex_block

        _cabi.require_cuda()
        block, names = read_vertex_block(path)  # [n, P] float32 view of the file's vertex records
        want = (["x", "y", "z", "scale_0", "scale_1", "scale_2", "rot_0", "rot_1", "rot_2", "rot_3", "opacity",
                 "f_dc_0", "f_dc_1", "f_dc_2"] + [f"f_rest_{j}" for j in range(45)])
        missing = [w for w in want if w not in names]
        if missing:
            raise ValueError(f"{path}: missing 3DGS properties {missing[:4]}")
        cols = (ctypes.c_int32 * 59)(*[names.index(w) for w in want])
        n, stride = block.shape
        raw = torch.from_numpy(block).pin_memory().to(device, non_blocking=True)
        f = lambda *shape: torch.empty(shape, dtype=torch.float32, device=device)
        xyz, ls, rot, op, planar = f(n, 3), f(n, 3), f(n, 4), f(n), f(3, 16, n)
        _cabi.call("gsr_ply_unpack", _cabi.ptr(raw), n, stride, cols, _cabi.ptr(xyz), _cabi.ptr(ls), _cabi.ptr(rot),
                   _cabi.ptr(op), _cabi.ptr(planar), _cabi.stream_handle())
        return GaussianModel(xyz, ls, rot, op, planar.permute(2, 1, 0))

    def c_struct(self) -> _cabi.Gaussians:
        g = _cabi.Gaussians()
        g.n = self.n
        for name, t in (("xyz", self.xyz), ("log_scales", self.log_scales), ("rots", self.rots)):
            assert t.dtype == torch.float32 and t.is_cuda
            setattr(g, name, t.data_ptr())
        g.xyz_sn, g.xyz_sc = self.xyz.stride()
        g.scale_sn, g.scale_sc = self.log_scales.stride()
        g.rot_sn, g.rot_sc = self.rots.stride()
        g.opacity_logit, g.opacity_sn = self.opacity.data_ptr(), self.opacity.stride(0)
        g.sh = self.sh.data_ptr()
        g.sh_sn, g.sh_sk, g.sh_sc = self.sh.stride()
        return g


def _align256(t: torch.Tensor) -> torch.Tensor:
    off = (-t.data_ptr()) % 256
    return t[off:]


class Renderer:
    """Owns the workspace for frames of one size class and enqueues whole frames on the current stream.

    Pair capacity (``max_dup``, the room for (Gaussian, tile) pairs): starts at ``dup_factor * n``; a frame that
    overflows is detected on the device (sticky ``status`` in ``info_dev``) and redrawn once with the capacity it
    reported; after a checked frame the capacity is tightened to 1.5x what that frame needed and then STAYS there for
    frames of the same (n, W, H) - grids, memsets and look-back tables of the pair sort are sized by capacity.
    """

    def __init__(self, device="cuda", dup_factor: float = 8.0, use_prepared: bool = True):
        _cabi.require_cuda()
        self.device = torch.device(device)
        self.dup_factor = dup_factor
        self.use_prepared = use_prepared
        self._ws: Optional[torch.Tensor] = None
        self._ws_key = None
        self._cap_key = None    # (n, W, H) the current capacity was established for
        self.max_dup = 0
        self.info_dev = torch.zeros(8, dtype=torch.int32, device=self.device)  # gsr_frame_info (32 bytes), sticky status
        self.adaptive_capacity = True
        self.frame_serial = 0   # bumped by every frame: backward must replay the workspace of ITS forward
        self._fwd_max_dup = 0   # capacity the last forward was enqueued with (what its backward must pass)

    # -- workspace ---------------------------------------------------------------------------
    def reserve(self, n: int, width: int, height: int, max_dup: Optional[int] = None) -> None:
        if max_dup is None:
            if self._cap_key == (n, width, height) and self.max_dup > 0:
                max_dup = self.max_dup          # established (possibly tightened) capacity is authoritative
            else:
                max_dup = max(int(self.dup_factor * n), 1 << 16)
        max_dup = min(int(max_dup), (1 << 30) - 1)
        key = (n, width, height, max_dup)
        if self._ws is not None and self._ws_key == key:
            return
        need = ctypes.c_size_t()
        _cabi.call("gsr_render_workspace_bytes", n, width, height, max_dup, ctypes.byref(need))
        if self._ws is None or self._ws.numel() < need.value + 256:
            self._ws = None
            self._ws = torch.empty(need.value + 256, dtype=torch.uint8, device=self.device)
        self._ws_key, self._cap_key, self.max_dup = key, (n, width, height), max_dup

    def _enqueue(self, model: GaussianModel, cam: CameraStruct, out, final_T, n_contrib, first_rank, last_rank, rb, re,
                 prepared: bool) -> None:
        self.frame_serial += 1
        self._fwd_max_dup = self.max_dup
        ws = _align256(self._ws)
        if prepared:
            _cabi.call("gsr_render_prepared_fwd", _cabi.ptr(model.prepare()), model.n, ctypes.byref(cam), self.max_dup,
                       _cabi.ptr(ws), ws.numel(), _cabi.ptr(out), _cabi.ptr(final_T), _cabi.ptr(n_contrib), first_rank,
                       last_rank, rb, re, _cabi.ptr(self.info_dev), _cabi.stream_handle())
        else:
            g = model.c_struct()
            _cabi.call("gsr_render_fwd", ctypes.byref(g), ctypes.byref(cam), self.max_dup, _cabi.ptr(ws),
                       ws.numel(), _cabi.ptr(out), _cabi.ptr(final_T), _cabi.ptr(n_contrib), first_rank, last_rank,
                       rb, re, _cabi.ptr(self.info_dev), _cabi.stream_handle())

    # -- one frame ---------------------------------------------------------------------------
    def render(self, model: GaussianModel, cam: CameraStruct, out: Optional[torch.Tensor] = None,
               final_T: Optional[torch.Tensor] = None, n_contrib: Optional[torch.Tensor] = None,
               first_rank: int = 0, last_rank: int = U32_MAX, tile_rows: Optional[Sequence[int]] = None,
               check: bool = True, prepared: Optional[bool] = None) -> torch.Tensor:
        """Enqueue one frame; returns image ``[H,W,3]`` float32 (the reference's ``screen.transpose(1,0)``).

        ``prepared`` (default: the renderer's ``use_prepared``) renders from ``model.prepare()``; ``False`` runs the
        fused raw-parameter kernel (one-shot frames, the differentiable path).  The two are bit-identical.
        ``check=False`` does not read the status back: it is sticky on the device, call ``check_status()`` after the
        last frame of a sequence."""
        W, H = cam.width, cam.height
        self.reserve(model.n, W, H)
        if out is None:
            out = torch.empty((H, W, 3), dtype=torch.float32, device=self.device)
        assert out.is_contiguous() and out.dtype == torch.float32 and tuple(out.shape) == (H, W, 3)
        rb, re = (0, -1) if tile_rows is None else (int(tile_rows[0]), int(tile_rows[1]))
        prepared = self.use_prepared if prepared is None else prepared
        while True:
            self._enqueue(model, cam, out, final_T, n_contrib, first_rank, last_rank, rb, re, prepared)
            if not check:
                return out
            info = self.frame_info()
            if info["status"] == _cabi.GSR_ERR_DUP_OVERFLOW:  # grow once to the reported need and redo the frame
                self.clear_status()
                before = self.max_dup
                self.reserve(model.n, W, H, max_dup=int(max(info["n_dup"], info["n_dup_peak"]) * 1.1) + 1024)
                if self.max_dup <= before:  # already at the 2^30 - 1 ceiling of the pair sort: nothing left to grow
                    raise _cabi.GsrError(info["status"], f"the frame needs {info['n_dup_peak']} (Gaussian, tile) pairs, "
                                                         f"above the supported capacity {self.max_dup}")
                continue
            if info["status"] != _cabi.GSR_OK:
                self.clear_status()
                raise _cabi.GsrError(info["status"], "frame failed")
            # Tighten the pair capacity to what frames of this scene need (+50 %); the allocation stays, and the
            # tightened value is what reserve() keeps using.  An overflow on a later view is detected on the device.
            tight = max(int(1.5 * info["n_dup"]) + 4096, 1 << 16)
            if self.adaptive_capacity and tight < self.max_dup and tile_rows is None:
                self.max_dup = tight
                self._ws_key = (model.n, W, H, tight)
            return out

    def frame_info(self) -> Dict[str, int]:
        """Synchronising read-back of the counters: n_visible / n_dup of the last frame, sticky status / n_dup_peak."""
        raw = self.info_dev.cpu().numpy().view(np.uint32)
        return dict(status=int(raw[0]), n_visible=int(raw[1]), n_dup=int(raw[2]), n_dup_peak=int(raw[3]),
                    warp_evals=int(raw[4]) | (int(raw[5]) << 32), pixel_evals=int(raw[6]) | (int(raw[7]) << 32))

    def clear_status(self) -> None:
        self.info_dev.zero_()

    def check_status(self) -> Dict[str, int]:
        """After a sequence of ``check=False`` frames: raise if ANY of them overflowed the pair capacity (their images
        are invalid); the capacity is raised to what the worst frame needed, so the caller can simply redraw."""
        info = self.frame_info()
        if info["status"] != _cabi.GSR_OK:
            self.clear_status()
            if info["status"] == _cabi.GSR_ERR_DUP_OVERFLOW and self._cap_key is not None:
                self.reserve(*self._cap_key, max_dup=int(info["n_dup_peak"] * 1.1) + 1024)
            raise _cabi.GsrError(info["status"], f"a frame needed {info['n_dup_peak']} (Gaussian, tile) pairs, more than "
                                                 f"the capacity it was drawn with; capacity raised to {self.max_dup}: redraw")
        return info

    def backward(self, model: GaussianModel, cam: CameraStruct, final_T: torch.Tensor, n_contrib: torch.Tensor,
                 grad_image: torch.Tensor, tile_rows: Optional[Sequence[int]] = None):
        """Gradients of the last frame rendered with this renderer (same model, same camera)."""
        n, dev = model.n, self.device
        grad_image = grad_image.contiguous().float()
        scratch = torch.empty((max(n, 1), 9), dtype=torch.float32, device=dev)
        d_xyz = torch.empty((n, 3), dtype=torch.float32, device=dev)
        d_ls = torch.empty((n, 3), dtype=torch.float32, device=dev)
        d_rot = torch.empty((n, 4), dtype=torch.float32, device=dev)
        d_op = torch.empty((n,), dtype=torch.float32, device=dev)
        d_sh = torch.empty_strided(tuple(model.sh.shape), model.sh.stride(), dtype=torch.float32, device=dev)
        rb, re = (0, -1) if tile_rows is None else (int(tile_rows[0]), int(tile_rows[1]))
        ws = _align256(self._ws)
        g = model.c_struct()
        # the capacity the FORWARD was enqueued with (render() may have tightened self.max_dup since)
        _cabi.call("gsr_render_bwd", ctypes.byref(g), ctypes.byref(cam), self._fwd_max_dup, _cabi.ptr(ws), ws.numel(),
                   _cabi.ptr(final_T), _cabi.ptr(n_contrib), _cabi.ptr(grad_image), _cabi.ptr(scratch), _cabi.ptr(d_xyz),
                   _cabi.ptr(d_ls), _cabi.ptr(d_rot), _cabi.ptr(d_op), _cabi.ptr(d_sh), *d_sh.stride(), rb, re,
                   _cabi.stream_handle())
        return d_xyz, d_ls, d_rot, d_op, d_sh


class _RenderFunction(torch.autograd.Function):
    """image = render(parameters): forward through gsr_render_fwd, backward through gsr_render_bwd."""

    @staticmethod
    def forward(ctx, xyz, log_scales, rots, opacity, sh, renderer, cam, tile_rows):
        model = GaussianModel(xyz.detach(), log_scales.detach(), rots.detach(), opacity.detach(), sh.detach())
        H, W = cam.height, cam.width
        final_T = torch.empty((H, W), dtype=torch.float32, device=renderer.device)
        n_contrib = torch.zeros((H, W), dtype=torch.int32, device=renderer.device)
        out = torch.zeros((H, W, 3), dtype=torch.float32, device=renderer.device)
        image = renderer.render(model, cam, out=out, final_T=final_T, n_contrib=n_contrib, tile_rows=tile_rows,
                                prepared=False)  # parameters change every step: nothing view-independent to keep
        ctx.model, ctx.renderer, ctx.cam, ctx.tile_rows = model, renderer, cam, tile_rows
        ctx.serial = renderer.frame_serial
        ctx.save_for_backward(final_T, n_contrib)
        return image

    @staticmethod
    def backward(ctx, grad_image):
        if ctx.renderer.frame_serial != ctx.serial:
            raise RuntimeError("the renderer drew another frame since this forward pass: its workspace (records, tile "
                               "lists) was overwritten; call backward before rendering again or use one Renderer per graph")
        final_T, n_contrib = ctx.saved_tensors
        grads = ctx.renderer.backward(ctx.model, ctx.cam, final_T, n_contrib, grad_image, ctx.tile_rows)
        return (*grads, None, None, None)


def render_differentiable(xyz, log_scales, rots, opacity_logits, sh, cam: CameraStruct,
                          renderer: Optional[Renderer] = None, tile_rows=None) -> torch.Tensor:
    """Differentiable frame: CUDA float32 tensors in (raw 3DGS parameters, sh [N,16,3] any strides), image out."""
    _cabi.require_cuda()
    r = renderer or default_renderer()
    for t in (xyz, log_scales, rots, opacity_logits, sh):
        if not (t.is_cuda and t.dtype == torch.float32):
            raise ValueError("render_differentiable needs CUDA float32 tensors")
    return _RenderFunction.apply(xyz.contiguous(), log_scales.contiguous(), rots.contiguous(),
                                 opacity_logits.contiguous(), sh, r, cam, tile_rows)


_default_renderer: Optional[Renderer] = None


def default_renderer() -> Renderer:
    global _default_renderer
    if _default_renderer is None:
        _default_renderer = Renderer()
    return _default_renderer


def render(means, log_scales, quats, opacity_logits, sh, qvec, tvec, fx, fy, cam_width, cam_height, width, height,
           degree: int = 3, renderer: Optional[Renderer] = None) -> torch.Tensor:
    """The reference's render path as a function (rasterize.py:347-452).

    Tensors may live on the host (they are uploaded - this is the end-to-end path the benchmark's ``e2e`` times)
    or on the GPU.  Returns the ``[H,W,3]`` float32 image on the GPU; no background term, black where nothing
    was blended, last row / column zero like the reference.
    """
    _cabi.require_cuda()
    r = renderer or default_renderer()
    model = GaussianModel.from_tensors(means, log_scales, quats, opacity_logits, sh, device=r.device)
    cam = make_camera(qvec, tvec, fx, fy, cam_width, cam_height, width, height, degree)
    return r.render(model, cam, prepared=False)  # one frame of this model: the fused raw kernel, no prepared copy


_lane_pool: Dict[int, list] = {}


def frame_lanes(count: int = 2, renderer: Optional[Renderer] = None):
    """`count` (renderer, stream) pairs for rendering independent frames concurrently: the latency-bound binning
    kernels of one frame overlap the issue-bound blend of another (+10 % frames/s with two lanes on B200)."""
    first = renderer or default_renderer()
    pool = _lane_pool.setdefault(id(first), [(first, None)])
    while len(pool) < count:
        pool.append((Renderer(first.device, first.dup_factor), torch.cuda.Stream(device=first.device)))
    return pool[:count]


def render_views(model: GaussianModel, intr, views, degree: int = 3, renderer: Optional[Renderer] = None,
                 out: Optional[torch.Tensor] = None, lanes: int = 3, _retry: bool = False) -> torch.Tensor:
    """Render a list of ``scenes.View`` with one resident model: ``[V,H,W,3]`` (video / orbit path).

    Frames are independent, so they are spread round-robin over `lanes` renderers/streams; the call returns with
    the current stream waiting on all of them.  A frame that overflows the pair capacity of its lane is detected (sticky
    device status, one read per lane) and the whole batch is redrawn once with the capacity the worst frame reported."""
    from .camera import camera_for

    pool = frame_lanes(max(1, lanes), renderer)
    dev = pool[0][0].device
    if out is None:
        out = torch.empty((len(views), intr.height, intr.width, 3), dtype=torch.float32, device=dev)
    main = torch.cuda.current_stream()
    start = torch.cuda.Event(); start.record(main)
    for i, v in enumerate(views):
        rend, strm = pool[i % len(pool)]
        cam = camera_for(intr, v, degree)
        if strm is None:
            rend.render(model, cam, out=out[i], check=False)
        else:
            if i < len(pool):
                strm.wait_event(start)
            with torch.cuda.stream(strm):
                rend.render(model, cam, out=out[i], check=False)
    for rend, strm in pool:
        if strm is not None:
            main.wait_stream(strm)
    # one synchronising status check per lane: the device-side status is sticky, so it covers EVERY frame of the lane
    overflow = None
    for rend, _ in pool[: min(len(pool), len(views))]:
        try:
            rend.check_status()
        except _cabi.GsrError as exc:
            if exc.status != _cabi.GSR_ERR_DUP_OVERFLOW:
                raise
            overflow = exc
    if overflow is not None:
        if _retry:
            raise overflow
        cap = max(rend.max_dup for rend, _ in pool)  # check_status() raised the capacity of the lanes that overflowed
        for rend, _ in pool:
            rend.reserve(model.n, intr.width, intr.height, max_dup=cap)
        return render_views(model, intr, views, degree, renderer, out, lanes, _retry=True)
    return out


def _host_struct(means, log_scales, quats, opacity_logits, sh):
    """``gsr_gaussians`` over HOST tensors (float32; dense rows, SH planar [3,16,N]-viewed-as-[N,16,3] or dense)."""
    keep = []

    def dense(t, inner):
        t = torch.as_tensor(t)
        if t.dtype != torch.float32 or t.is_cuda or not t.is_contiguous():
            t = t.detach().to("cpu", torch.float32).contiguous()
        assert t.numel() % inner == 0
        keep.append(t)
        return t

    xyz, ls, rot, op = dense(means, 3), dense(log_scales, 3), dense(quats, 4), dense(opacity_logits, 1).reshape(-1)
    n = xyz.shape[0]
    sh = torch.as_tensor(sh)
    assert sh.dim() == 3 and tuple(sh.shape[1:]) == (16, 3), "sh must be [N,16,3]"
    if sh.dtype != torch.float32 or sh.is_cuda or sh.stride() not in ((1, n, 16 * n), (48, 3, 1)):
        sh = sh.detach().to("cpu", torch.float32).contiguous()
    keep.append(sh)
    g = _cabi.Gaussians()
    g.n = n
    g.xyz, g.xyz_sn, g.xyz_sc = xyz.data_ptr(), 3, 1
    g.log_scales, g.scale_sn, g.scale_sc = ls.data_ptr(), 3, 1
    g.rots, g.rot_sn, g.rot_sc = rot.data_ptr(), 4, 1
    g.opacity_logit, g.opacity_sn = op.data_ptr(), 1
    g.sh = sh.data_ptr()
    g.sh_sn, g.sh_sk, g.sh_sc = sh.stride()
    return g, keep


class HostRenderer:
    """Frames from HOST tensors through the C-level pipeline ``gsr_host_renderer`` (include/gsr_b200.h): chunked upload
    overlapped with the per-chunk preprocess, binning + blend, image read-back on its own stream, `depth` frames in
    flight.  ``render`` is synchronous; ``submit`` / ``wait`` pipeline consecutive frames."""

    def __init__(self, n_max: int, width: int, height: int, max_dup: Optional[int] = None, depth: int = 2, device="cuda"):
        _cabi.require_cuda()
        self.device = torch.device(device)
        self.n_max, self.width, self.height, self.depth = int(n_max), int(width), int(height), int(depth)
        self.max_dup = int(max_dup if max_dup is not None else max(8 * n_max, 1 << 16))
        self._h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _cabi.call("gsr_host_renderer_create", self.n_max, self.width, self.height, self.max_dup, self.depth,
                       ctypes.byref(self._h))
        self._alive = {}

    def close(self) -> None:
        if self._h:
            _cabi.call("gsr_host_renderer_destroy", self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001 - interpreter shutdown
            pass

    def submit(self, means, log_scales, quats, opacity_logits, sh, cam: CameraStruct, image_host: torch.Tensor) -> int:
        assert not image_host.is_cuda and image_host.is_contiguous() and image_host.dtype == torch.float32
        assert tuple(image_host.shape) == (self.height, self.width, 3)
        g, keep = _host_struct(means, log_scales, quats, opacity_logits, sh)
        t = ctypes.c_int32()
        _cabi.call("gsr_host_renderer_submit", self._h, ctypes.byref(g), ctypes.byref(cam), _cabi.ptr(image_host),
                   ctypes.byref(t))
        self._alive[t.value] = (keep, image_host)  # host buffers must outlive the frame
        return t.value

    def wait(self, ticket: int) -> Dict[str, int]:
        info = _cabi.FrameInfo()
        try:
            _cabi.call("gsr_host_renderer_wait", self._h, ticket, ctypes.byref(info))
        finally:
            self._alive.pop(ticket, None)
        return dict(status=info.status, n_visible=info.n_visible, n_dup=info.n_dup, n_dup_peak=info.n_dup_peak)

    def render(self, means, log_scales, quats, opacity_logits, sh, cam: CameraStruct,
               image_host: Optional[torch.Tensor] = None) -> torch.Tensor:
        if image_host is None:
            image_host = torch.empty((self.height, self.width, 3), dtype=torch.float32).pin_memory()
        self.wait(self.submit(means, log_scales, quats, opacity_logits, sh, cam, image_host))
        return image_host


class BatchRenderer:
    """A batch of views of one resident (prepared) model through ONE C-ABI call, ``gsr_render_views_fwd``: the library
    round-robins the frames over `lanes` caller-owned streams / workspaces, so consecutive frames overlap on the GPU
    and the host pays one foreign call for the whole batch instead of one per frame."""

    def __init__(self, lanes: int = 3, device="cuda", dup_factor: float = 8.0):
        _cabi.require_cuda()
        self.device = torch.device(device)
        self.lanes = max(1, int(lanes))
        self.dup_factor = dup_factor
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(self.lanes)]
        self.info_dev = torch.zeros((self.lanes, 8), dtype=torch.int32, device=self.device)
        self._ws: list = []
        self._key = None
        self.max_dup = 0
        self.ws_bytes = 0

    def reserve(self, n: int, width: int, height: int, max_dup: Optional[int] = None) -> None:
        if max_dup is None:
            max_dup = self.max_dup if (self._key is not None and self._key[:3] == (n, width, height)) else \
                max(int(self.dup_factor * n), 1 << 16)
        max_dup = min(int(max_dup), (1 << 30) - 1)
        key = (n, width, height, max_dup)
        if key == self._key:
            return
        need = ctypes.c_size_t()
        _cabi.call("gsr_render_workspace_bytes", n, width, height, max_dup, ctypes.byref(need))
        if not self._ws or self._ws[0].numel() < need.value + 256:
            self._ws = []
            self._ws = [torch.empty(need.value + 256, dtype=torch.uint8, device=self.device) for _ in range(self.lanes)]
        self._key, self.max_dup, self.ws_bytes = key, max_dup, need.value

    def render(self, model: GaussianModel, cams: Sequence[CameraStruct], out: Optional[torch.Tensor] = None,
               check: bool = True) -> torch.Tensor:
        """``[V,H,W,3]``; returns with the current stream waiting on every lane.  ``check`` reads the sticky status of
        each lane once after the batch and redraws the batch with a larger capacity if any frame overflowed."""
        V = len(cams)
        W, H = cams[0].width, cams[0].height
        assert all(c.width == W and c.height == H for c in cams)
        self.reserve(model.n, W, H)
        if out is None:
            out = torch.empty((V, H, W, 3), dtype=torch.float32, device=self.device)
        assert out.is_contiguous() and tuple(out.shape) == (V, H, W, 3) and out.dtype == torch.float32
        prepared = model.prepare()
        cam_arr = (CameraStruct * V)(*cams)
        for attempt in range(2):
            ws_arr = (ctypes.c_void_p * self.lanes)(*[_align256(w).data_ptr() for w in self._ws])
            st_arr = (ctypes.c_void_p * self.lanes)(*[s.cuda_stream for s in self.streams])
            main = torch.cuda.current_stream()
            ev = torch.cuda.Event(); ev.record(main)
            for s_ in self.streams:
                s_.wait_event(ev)
            _cabi.call("gsr_render_views_fwd", _cabi.ptr(prepared), model.n, cam_arr, V, self.max_dup, ws_arr,
                       self.ws_bytes, self.lanes, st_arr, _cabi.ptr(out), 3 * W * H, _cabi.ptr(self.info_dev))
            for s_ in self.streams:
                main.wait_stream(s_)
            if not check:
                return out
            raw = self.info_dev.cpu().numpy().view(np.uint32)
            if not raw[:, 0].any():
                return out
            self.info_dev.zero_()
            if attempt == 1 or not (raw[:, 0] == _cabi.GSR_ERR_DUP_OVERFLOW).any():
                raise _cabi.GsrError(int(raw[:, 0].max()), f"a frame of the batch failed (needs {int(raw[:, 3].max())} pairs)")
            self.reserve(model.n, W, H, max_dup=int(int(raw[:, 3].max()) * 1.1) + 1024)
        return out


# ----------------------------------------------------------------------------------------------
# stage-level wrappers (tests and profiling drive the stages one by one)
# ----------------------------------------------------------------------------------------------
def preprocess(model: GaussianModel, cam: CameraStruct, debug: bool = False, tile_rows=None,
               prepared: bool = False) -> Dict[str, torch.Tensor]:
    """Per-view preprocess stage.  ``prepared=True`` runs the kernel of the resident-scene path over ``model.prepare()``
    (no debug outputs); its outputs are bit-identical to the raw kernel's."""
    _cabi.require_cuda()
    n, dev = model.n, model.device
    rb, re = (0, 1 << 30) if tile_rows is None else tile_rows
    out = dict(depth_key=torch.empty(n, dtype=torch.int32, device=dev),
               tiles_touched=torch.empty(n, dtype=torch.int32, device=dev),
               tile_span=torch.zeros((n, 4), dtype=torch.int32, device=dev),
               rec=torch.zeros((n, _cabi.REC_FLOATS), dtype=torch.float32, device=dev),
               depth_hist=torch.empty(1024, dtype=torch.int32, device=dev))
    if debug:
        out.update(cov3d=torch.empty((n, 3, 3), dtype=torch.float32, device=dev),
                   cov2d=torch.empty((n, 2, 2), dtype=torch.float32, device=dev),
                   tile_rect=torch.empty((n, 4), dtype=torch.int32, device=dev),
                   mean2d=torch.empty((n, 2), dtype=torch.float32, device=dev))
    if prepared:
        assert not debug, "debug outputs come from the raw kernel"
        _cabi.call("gsr_preprocess_prepared_fwd", _cabi.ptr(model.prepare()), n, ctypes.byref(cam), rb, re,
                   _cabi.ptr(out["depth_key"]), _cabi.ptr(out["tiles_touched"]), _cabi.ptr(out["tile_span"]),
                   _cabi.ptr(out["rec"]), _cabi.ptr(out["depth_hist"]), _cabi.stream_handle())
    else:
        g = model.c_struct()
        _cabi.call("gsr_preprocess_fwd", ctypes.byref(g), ctypes.byref(cam), rb, re, _cabi.ptr(out["depth_key"]),
                   _cabi.ptr(out["tiles_touched"]), _cabi.ptr(out["tile_span"]), _cabi.ptr(out["rec"]),
                   _cabi.ptr(out.get("cov3d")), _cabi.ptr(out.get("cov2d")), _cabi.ptr(out.get("tile_rect")),
                   _cabi.ptr(out.get("mean2d")), _cabi.ptr(out["depth_hist"]), _cabi.stream_handle())
    # views of the span: the clipped (and culling-shrunk) tile rect as u16 x0,y0,x1,y1 and the 64-bit keep mask
    out["rect"] = out["tile_span"].view(torch.int16)[:, :4]
    out["tile_mask"] = out["tile_span"][:, 2:4]
    return out


class Stages:
    """The three stages of a frame over PREALLOCATED buffers, one C-ABI call each and nothing else on the stream:
    what the benchmark times per stage with CUDA events (the dict-returning wrappers above allocate and zero-fill their
    outputs on every call)."""

    def __init__(self, model: GaussianModel, width: int, height: int, max_dup: int):
        _cabi.require_cuda()
        dev, n = model.device, model.n
        self.model, self.n, self.W, self.H, self.max_dup = model, n, width, height, int(max_dup)
        i32 = lambda *shape: torch.empty(shape, dtype=torch.int32, device=dev)
        self.depth_key, self.tiles_touched, self.tile_span = i32(max(n, 1)), i32(max(n, 1)), i32(max(n, 1), 4)
        self.rec = torch.zeros((max(n, 1), _cabi.REC_FLOATS), dtype=torch.float32, device=dev)
        self.depth_hist, self.order, self.blend_rank = i32(1024), i32(max(n, 1)), i32(max(n, 1))
        self.tile_gids = i32(self.max_dup)
        self.tile_ranges = i32(((width + 15) // 16) * ((height + 15) // 16), 2)
        self.info = torch.zeros(8, dtype=torch.int32, device=dev)
        need = ctypes.c_size_t()
        _cabi.call("gsr_bin_workspace_bytes", n, width, height, self.max_dup, ctypes.byref(need))
        self._ws = torch.empty(need.value + 256, dtype=torch.uint8, device=dev)
        self.image = torch.empty((height, width, 3), dtype=torch.float32, device=dev)

    def preprocess(self, cam: CameraStruct, prepared: bool = True, rows: Tuple[int, int] = (0, 1 << 30)) -> None:
        """`rows`: band of tile rows (prepared path), as a rank of a band-sharded render would pass it."""
        if prepared:
            _cabi.call("gsr_preprocess_prepared_fwd", _cabi.ptr(self.model.prepare()), self.n, ctypes.byref(cam), rows[0], rows[1],
                       _cabi.ptr(self.depth_key), _cabi.ptr(self.tiles_touched), _cabi.ptr(self.tile_span),
                       _cabi.ptr(self.rec), _cabi.ptr(self.depth_hist), _cabi.stream_handle())
        else:
            g = self.model.c_struct()
            _cabi.call("gsr_preprocess_fwd", ctypes.byref(g), ctypes.byref(cam), 0, 1 << 30, _cabi.ptr(self.depth_key),
                       _cabi.ptr(self.tiles_touched), _cabi.ptr(self.tile_span), _cabi.ptr(self.rec), None, None, None,
                       None, _cabi.ptr(self.depth_hist), _cabi.stream_handle())

    def bin(self) -> None:
        ws = _align256(self._ws)
        _cabi.call("gsr_bin", self.n, _cabi.ptr(self.depth_key), _cabi.ptr(self.depth_hist), _cabi.ptr(self.tiles_touched),
                   _cabi.ptr(self.tile_span), self.W, self.H, self.max_dup, _cabi.ptr(self.order), None,
                   _cabi.ptr(self.tile_gids), _cabi.ptr(self.tile_ranges), _cabi.ptr(self.info), _cabi.ptr(ws), ws.numel(),
                   _cabi.stream_handle())

    def blend(self, count: Optional[torch.Tensor] = None) -> None:
        _cabi.call("gsr_blend_fwd", _cabi.ptr(self.rec), _cabi.ptr(self.tile_gids), _cabi.ptr(self.tile_ranges), self.W,
                   self.H, 0, -1, None, 0, U32_MAX, _cabi.ptr(self.image), None, None, _cabi.ptr(count),
                   _cabi.stream_handle())


def sort_pairs(keys: torch.Tensor, values: Optional[torch.Tensor], begin_bit: int = 0, end_bit: int = 32,
               n_dev: Optional[torch.Tensor] = None):
    _cabi.require_cuda()
    n = keys.numel()
    need = ctypes.c_size_t()
    _cabi.call("gsr_sort_workspace_bytes", n, ctypes.byref(need))
    ws = _align256(torch.empty(need.value + 256, dtype=torch.uint8, device=keys.device))
    ko, vo, kt, vt = (torch.empty_like(keys) for _ in range(4))
    _cabi.call("gsr_sort_pairs", _cabi.ptr(keys), _cabi.ptr(values), _cabi.ptr(ko), _cabi.ptr(vo), _cabi.ptr(kt),
               _cabi.ptr(vt), n, _cabi.ptr(n_dev), begin_bit, end_bit, _cabi.ptr(ws), ws.numel(), _cabi.stream_handle())
    return ko, vo


def bin_tiles(pre: Dict[str, torch.Tensor], width: int, height: int, max_dup: int) -> Dict[str, torch.Tensor]:
    """Tile lists from the preprocess outputs, plus the global depth order and the blend ranks."""
    _cabi.require_cuda()
    dev = pre["depth_key"].device
    n = pre["depth_key"].numel()
    tiles = ((width + 15) // 16) * ((height + 15) // 16)
    need = ctypes.c_size_t()
    _cabi.call("gsr_bin_workspace_bytes", n, width, height, max_dup, ctypes.byref(need))
    ws = _align256(torch.empty(need.value + 256, dtype=torch.uint8, device=dev))
    out = dict(order=torch.empty(n, dtype=torch.int32, device=dev),
               blend_rank=torch.zeros(n, dtype=torch.int32, device=dev),
               tile_gids=torch.empty(max_dup, dtype=torch.int32, device=dev),
               tile_ranges=torch.empty((tiles, 2), dtype=torch.int32, device=dev),
               info=torch.zeros(8, dtype=torch.int32, device=dev))
    _cabi.call("gsr_bin", n, _cabi.ptr(pre["depth_key"]), _cabi.ptr(pre.get("depth_hist")),
               _cabi.ptr(pre["tiles_touched"]), _cabi.ptr(pre["tile_span"]), width, height, max_dup,
               _cabi.ptr(out["order"]), _cabi.ptr(out["blend_rank"]), _cabi.ptr(out["tile_gids"]),
               _cabi.ptr(out["tile_ranges"]), _cabi.ptr(out["info"]), _cabi.ptr(ws), ws.numel(), _cabi.stream_handle())
    out["_ws"] = ws
    return out


def blend(pre, binned, width: int, height: int, count: bool = False, first_rank: int = 0, last_rank: int = U32_MAX,
          tile_rows=None, track: bool = True):
    """``track=False`` omits ``n_contrib`` (only a backward pass needs it) - the instantiation a plain render uses."""
    _cabi.require_cuda()
    dev = pre["rec"].device
    image = torch.empty((height, width, 3), dtype=torch.float32, device=dev)
    final_T = torch.empty((height, width), dtype=torch.float32, device=dev)
    n_contrib = torch.empty((height, width), dtype=torch.int32, device=dev) if track else None
    info = torch.zeros(8, dtype=torch.int32, device=dev) if count else None
    rb, re = (0, -1) if tile_rows is None else tile_rows
    _cabi.call("gsr_blend_fwd", _cabi.ptr(pre["rec"]), _cabi.ptr(binned["tile_gids"]), _cabi.ptr(binned["tile_ranges"]),
               width, height, rb, re, _cabi.ptr(binned.get("blend_rank")), first_rank, last_rank, _cabi.ptr(image),
               _cabi.ptr(final_T), _cabi.ptr(n_contrib), _cabi.ptr(info), _cabi.stream_handle())
    return dict(image=image, final_T=final_T, n_contrib=n_contrib, info=info)
