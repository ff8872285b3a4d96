"""Out-of-bounds and uninitialised-memory checks of the C-ABI entry points, with guard zones and poison patterns.

compute-sanitizer is closed on the B200 pool (SURVEY section 5 planned memcheck / racecheck runs), so the same questions
are asked with the library's own contract: every buffer the caller hands over (workspace, prepared model, image,
final_T, n_contrib, gradients) sits between two guard zones holding a byte pattern, and is itself pre-filled with a
poison pattern (0x00, 0xFF = NaN / huge indices, 0xA5).  After every call the guards must be untouched (no write outside
the stated sizes) and the results must be bit-identical whatever the poison was (no read of memory the library did not
write itself in that frame).  Sizes make every multi-CTA path run: several one-sweep tiles per sort pass, the decoupled
look-back of the emission, ragged last block of 32, partial tiles on both image edges.
"""
import ctypes

import numpy as np
import pytest
import torch

from gsr_b200 import camera, scenes

pytestmark = pytest.mark.gpu

GUARD = 1 << 16
GUARD_BYTE = 0x5C
N, W, H = 40_013, 333, 217


class Arena:
    """Caller-owned device buffers between guard zones."""

    def __init__(self):
        self.blocks = []

    def take(self, nbytes: int, fill: int) -> torch.Tensor:
        """`nbytes` of uint8 (256-byte aligned start) pre-filled with `fill`, guard zones on both sides."""
        nbytes = int(nbytes)
        big = torch.full((GUARD + 256 + nbytes + GUARD,), GUARD_BYTE, dtype=torch.uint8, device="cuda")
        off = GUARD + (-(big.data_ptr() + GUARD)) % 256
        body = big[off:off + nbytes]
        body.fill_(fill)
        self.blocks.append((big, off, nbytes))
        return body

    def f32(self, shape, fill: int) -> torch.Tensor:
        return self.take(4 * int(np.prod(shape)), fill).view(torch.float32).view(*shape)

    def i32(self, shape, fill: int) -> torch.Tensor:
        return self.take(4 * int(np.prod(shape)), fill).view(torch.int32).view(*shape)

    def check(self, what: str) -> None:
        torch.cuda.synchronize()
        for k, (big, off, nbytes) in enumerate(self.blocks):
            assert bool((big[:off] == GUARD_BYTE).all()), f"{what}: block {k} written BEFORE its start"
            assert bool((big[off + nbytes:] == GUARD_BYTE).all()), f"{what}: block {k} ({nbytes} bytes) written PAST its end"


@pytest.fixture(scope="module")
def scene():
    from gsr_b200 import render as R

    cloud = scenes.synth_v1(N, 7, scale_factor=0.3)
    intr = scenes.Intrinsics.synthetic(W, H)
    cams = [camera.camera_for(intr, v, 3) for v in scenes.orbit_views(2, total=16, first=2)]
    model = R.GaussianModel.from_cloud(cloud)
    r = R.Renderer()
    r.adaptive_capacity = False
    want = [r.render(model, c).clone() for c in cams]
    info = r.frame_info()
    assert info["n_dup"] > 100_000 and info["n_visible"] > 24_000   # multi-tile sorts, multi-CTA emission
    return cloud, cams, model, want, r.max_dup


def _ws_bytes(n, max_dup):
    from gsr_b200 import _cabi

    need = ctypes.c_size_t()
    _cabi.call("gsr_render_workspace_bytes", n, W, H, max_dup, ctypes.byref(need))
    return need.value


@pytest.mark.parametrize("poison", [0x00, 0xFF, 0xA5])
def test_forward_paths_stay_inside_their_buffers_and_ignore_stale_memory(scene, poison):
    from gsr_b200 import _cabi, render as R

    cloud, cams, model, want, max_dup = scene
    arena = Arena()
    ws = arena.take(_ws_bytes(N, max_dup), poison)
    need = ctypes.c_size_t()
    _cabi.call("gsr_prepared_bytes", N, ctypes.byref(need))
    prepared = arena.take(need.value, poison)
    info = arena.i32((8,), 0)
    g = model.c_struct()
    _cabi.call("gsr_prepare", ctypes.byref(g), _cabi.ptr(prepared), prepared.numel(), _cabi.stream_handle())
    arena.check("gsr_prepare")
    assert torch.equal(prepared, model.prepare().view(torch.uint8)[:prepared.numel()]), "prepared model depends on stale bytes"

    def frame(cam, use_prepared, first=0, last=R.U32_MAX, rows=(0, -1), track=False):
        image = arena.f32((H, W, 3), poison)
        final_T = arena.f32((H, W), poison) if track else None
        n_contrib = arena.i32((H, W), poison) if track else None
        ws.fill_(poison)  # stale contents of a previous frame must not matter either
        if use_prepared:
            _cabi.call("gsr_render_prepared_fwd", _cabi.ptr(prepared), N, ctypes.byref(cam), max_dup, _cabi.ptr(ws),
                       ws.numel(), _cabi.ptr(image), _cabi.ptr(final_T), _cabi.ptr(n_contrib), first, last, rows[0],
                       rows[1], _cabi.ptr(info), _cabi.stream_handle())
        else:
            _cabi.call("gsr_render_fwd", ctypes.byref(g), ctypes.byref(cam), max_dup, _cabi.ptr(ws), ws.numel(),
                       _cabi.ptr(image), _cabi.ptr(final_T), _cabi.ptr(n_contrib), first, last, rows[0], rows[1],
                       _cabi.ptr(info), _cabi.stream_handle())
        arena.check(f"frame(prepared={use_prepared}, window=({first},{last}), rows={rows}, track={track})")
        return image, final_T, n_contrib

    for cam, img in zip(cams, want):
        for use_prepared in (True, False):
            got, _, _ = frame(cam, use_prepared)
            assert torch.equal(got, img), f"image depends on stale memory (prepared={use_prepared}, poison {poison:#x})"
    # contributor tracking (the training forward), a rank window covering everything, a band of tile rows
    got, final_T, n_contrib = frame(cams[0], False, track=True)
    assert torch.equal(got, want[0]) and bool(torch.isfinite(final_T).all()) and int(n_contrib.min()) >= 0
    got, _, _ = frame(cams[0], True, first=0, last=N + 7)
    assert torch.equal(got, want[0])
    rows = (3, 9)
    got, _, _ = frame(cams[0], True, rows=rows)
    assert torch.equal(got[rows[0] * 16:rows[1] * 16], want[0][rows[0] * 16:rows[1] * 16])
    outside = torch.cat([got[:rows[0] * 16].reshape(-1), got[rows[1] * 16:].reshape(-1)]).view(torch.uint8)
    assert bool((outside == poison).all()), "a band render wrote image rows outside its band"
    assert int(info[0]) == 0


@pytest.mark.parametrize("poison", [0x00, 0xFF])
def test_backward_stays_inside_its_buffers_and_writes_every_gradient(scene, poison):
    from gsr_b200 import _cabi

    cloud, cams, model, want, max_dup = scene
    cam = cams[1]
    g = model.c_struct()

    def grads(fill):
        arena = Arena()
        ws = arena.take(_ws_bytes(N, max_dup), fill)
        image, final_T, n_contrib = arena.f32((H, W, 3), fill), arena.f32((H, W), fill), arena.i32((H, W), fill)
        info = arena.i32((8,), 0)
        _cabi.call("gsr_render_fwd", ctypes.byref(g), ctypes.byref(cam), max_dup, _cabi.ptr(ws), ws.numel(),
                   _cabi.ptr(image), _cabi.ptr(final_T), _cabi.ptr(n_contrib), 0, 0xFFFFFFFF, 0, -1, _cabi.ptr(info),
                   _cabi.stream_handle())
        dL = torch.linspace(0.5, 1.5, H * W * 3, device="cuda").reshape(H, W, 3).contiguous()
        out = dict(scratch=arena.f32((N, 9), fill), d_xyz=arena.f32((N, 3), fill), d_ls=arena.f32((N, 3), fill),
                   d_rot=arena.f32((N, 4), fill), d_op=arena.f32((N,), fill), d_sh=arena.f32((3, 16, N), fill))
        d_sh = out["d_sh"].permute(2, 1, 0)  # [N,16,3] view of planar storage, like the model's SH block
        _cabi.call("gsr_render_bwd", ctypes.byref(g), ctypes.byref(cam), max_dup, _cabi.ptr(ws), ws.numel(),
                   _cabi.ptr(final_T), _cabi.ptr(n_contrib), _cabi.ptr(dL), _cabi.ptr(out["scratch"]),
                   _cabi.ptr(out["d_xyz"]), _cabi.ptr(out["d_ls"]), _cabi.ptr(out["d_rot"]), _cabi.ptr(out["d_op"]),
                   _cabi.ptr(out["d_sh"]), *d_sh.stride(), 0, -1, _cabi.stream_handle())
        arena.check("forward + backward")
        assert int(info[0]) == 0
        return out

    a, b = grads(poison), grads(0x33)
    for k in ("d_xyz", "d_ls", "d_rot", "d_op", "d_sh"):
        assert bool(torch.isfinite(a[k]).all()), f"{k}: an entry was left unwritten (poison {poison:#x}) or is not finite"
        # sums of float atomics: equal to rounding, never to a poison-sized difference
        scale = float(b[k].abs().max()) + 1e-20
        assert float((a[k] - b[k]).abs().max()) <= 1e-4 * scale, f"{k} depends on stale memory"


def test_two_host_threads_render_concurrently(scene):
    """entry points are stream-ordered and thread-safe per stream.  Two host threads (ctypes drops the GIL
    around every foreign call), each with its own renderer and stream, enqueue frames of different views at the same
    time; every frame must equal the one rendered alone.  The error string is per thread."""
    import threading

    from gsr_b200 import _cabi, render as R

    cloud, cams, model, want, max_dup = scene
    model.prepare()
    torch.cuda.synchronize()
    errors, barrier = [], threading.Barrier(2)

    def worker(k):
        try:
            torch.cuda.set_device(0)
            stream = torch.cuda.Stream()
            r = R.Renderer()
            barrier.wait()
            with torch.cuda.stream(stream):
                for it in range(12):
                    img = r.render(model, cams[k], check=(it % 4 == 3), prepared=(it % 2 == 0))
                    if it % 4 == 3 and not torch.equal(img, want[k]):
                        errors.append(f"thread {k}, frame {it}: image differs from the frame rendered alone")
            # a failing call in each thread, with different messages: after BOTH have failed, each thread still reads its own
            expect = ("bad arguments", "depth must be 1..4")[k]
            barrier.wait()
            try:
                if k == 0:
                    _cabi.call("gsr_render_workspace_bytes", N, -1, H, max_dup, ctypes.byref(ctypes.c_size_t()))
                else:
                    _cabi.call("gsr_host_renderer_create", 0, 8, 8, 8, 9, ctypes.byref(ctypes.c_void_p()))
                errors.append(f"thread {k}: invalid arguments accepted")
            except _cabi.GsrError as exc:
                if expect not in str(exc):
                    errors.append(f"thread {k}: unexpected error string {exc}")
            barrier.wait()
            again = _cabi.load().gsr_last_error().decode()
            if expect not in again:
                errors.append(f"thread {k}: gsr_last_error() is not per thread: {again!r}")
            stream.synchronize()
        except Exception as exc:  # noqa: BLE001 - reported by the main thread
            errors.append(f"thread {k}: {exc!r}")
            try:
                barrier.abort()
            except Exception:  # noqa: BLE001
                pass

    threads = [threading.Thread(target=worker, args=(k,)) for k in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(120)
    assert not errors, errors
