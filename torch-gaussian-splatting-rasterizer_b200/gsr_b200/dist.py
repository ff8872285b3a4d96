"""Multi-GPU rendering: one process per GPU, Gaussians replicated, work sharded by view or by tile-row band.

The reference is single-process; the path shards naturally, so the only collective is the
gather of finished pixels (NCCL all_gather over NVLink on GPUs, gloo in the CPU tests):

  * by VIEW  (video / orbit, BASELINE.json config 4): rank r renders views r, r+P, r+2P, ...; one all_gather per
    round of P frames, issued on a side stream so that it overlaps the next render.
  * by BAND  (one large frame, config 5): the tile rows are cut into P contiguous bands balanced on the number of
    (Gaussian, tile) pairs per row; every rank preprocesses all Gaussians (cheaper than exchanging them) but bins
    and blends only its band; one all_gather of the (padded) bands.

Everything that decides WHO renders WHAT is a pure function below, so it is testable without a GPU.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

TILE = 16


# ----------------------------------------------------------------------------------------------
# pure partitioning logic
# ----------------------------------------------------------------------------------------------
def views_of_rank(n_views: int, rank: int, world: int) -> List[int]:
    """Round-robin: rank r owns views r, r+world, ... (keeps every round of `world` frames in orbit order)."""
    return list(range(rank, n_views, world))


def rounds(n_views: int, world: int) -> int:
    return (n_views + world - 1) // world


def partition_bands(row_loads: Sequence[float], world: int) -> List[Tuple[int, int]]:
    """Cut tile rows 0..R into `world` contiguous [begin, end) bands with balanced total load.

    Band k ends at the first row boundary where the running load reaches k/world of the total; every band gets at
    least one row while rows remain (with fewer rows than ranks the last bands are empty).  Deterministic, so
    every rank computes the same cut.
    """
    if world <= 0:
        raise ValueError("world must be positive")
    rows = len(row_loads)
    prefix = [0.0]
    for v in row_loads:
        prefix.append(prefix[-1] + float(v))
    total = prefix[-1]
    bounds = [0]
    for k in range(1, world):
        lo = bounds[-1]
        if total > 0:
            target = total * k / world
            e = lo
            while e < rows and prefix[e] < target - 1e-9:
                e += 1
        else:
            e = (rows * k) // world
        e = max(e, min(lo + 1, rows))            # at least one row for band k-1 while rows remain
        e = min(e, max(lo, rows - (world - k)))  # leave a row for each band still to come
        bounds.append(max(e, lo))
    bounds.append(rows)
    return [(bounds[k], bounds[k + 1]) for k in range(world)]


def band_pixel_rows(band: Tuple[int, int], height: int) -> Tuple[int, int]:
    return min(band[0] * TILE, height), min(band[1] * TILE, height)


def assemble_bands(gathered: torch.Tensor, bands: Sequence[Tuple[int, int]], height: int) -> torch.Tensor:
    """`gathered` is [world, max_band_px, W, 3] (each rank's band at the top, padded); returns the [H,W,3] frame."""
    parts = []
    for r, band in enumerate(bands):
        y0, y1 = band_pixel_rows(band, height)
        parts.append(gathered[r, : y1 - y0])
    return torch.cat(parts, dim=0)


# ----------------------------------------------------------------------------------------------
# collectives
# ----------------------------------------------------------------------------------------------
def _world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def _all_gather_stack(dst: torch.Tensor, src: torch.Tensor) -> None:
    """dst[r] = rank r's src.  Flattened so that NCCL and gloo agree on the shapes."""
    assert dst.is_contiguous() and src.is_contiguous()
    dist.all_gather_into_tensor(dst.view(-1), src.view(-1))


class PeerGather:
    """All-gather of finished frames by peer-to-peer DMA instead of an NCCL kernel.

    Every rank owns one symmetric buffer ``[slots, world, *frame_shape]`` (``torch.distributed._symmetric_memory``:
    cuMem allocations mapped into every peer over NVLink / NVSwitch).  ``push(frame, slot)`` copies the rank's frame
    into ``[slot, rank]`` of EVERY rank's buffer - plain device-to-device copies that the copy engines execute, so the
    render kernels of the next frame keep all 148 SMs (an NCCL all_gather kernel of 8 x 25 MB frames took ~10 % of
    the 8-GPU frame rate) - and then joins a device-side barrier; once it has passed on a rank, ``view(slot)`` holds
    the frames of all ranks there.  Everything is enqueued on the caller's current stream.
    """

    def __init__(self, frame_shape: Sequence[int], device, slots: int = 2, group=None):
        import torch.distributed._symmetric_memory as symm

        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.slots = slots
        shape = (slots, self.world, *frame_shape)
        self.buf = symm.empty(shape, dtype=torch.float32, device=device)
        self.handle = symm.rendezvous(self.buf, group if group is not None else dist.group.WORLD)
        self.peers = [self.handle.get_buffer(p, shape, torch.float32) for p in range(self.world)]
        # one DMA stream per few destinations: a single copy stream serialises the world-size copies of a frame on one
        # copy engine (8 x 25 MB took about as long as a frame renders); several streams keep more engines / links busy
        self.copy_streams = [torch.cuda.Stream(device=device) for _ in range(min(4, max(1, self.world - 1)))]

    @staticmethod
    def create(frame_shape: Sequence[int], device, slots: int = 2, group=None) -> Optional["PeerGather"]:
        """None when the process group / platform cannot do it (CPU tests, no peer access): callers fall back to NCCL."""
        try:
            if not (dist.is_initialized() and dist.get_backend(group) == "nccl" and torch.device(device).type == "cuda"):
                return None
            return PeerGather(frame_shape, device, slots, group)
        except Exception as exc:  # noqa: BLE001 - any failure here only means "use the NCCL gather"
            import logging

            logging.getLogger(__name__).warning("peer-to-peer gather unavailable (%r): using NCCL all_gather", exc)
            return None

    def push(self, frame: torch.Tensor, slot: int) -> torch.Tensor:
        cur = torch.cuda.current_stream()
        ready = torch.cuda.Event()
        ready.record(cur)  # the frame is complete (and the slot free) at this point of the caller's stream
        for st in self.copy_streams:
            st.wait_event(ready)
        for k in range(self.world):  # staggered so that the ranks do not all write to the same peer at once
            p = (self.rank + k) % self.world
            with torch.cuda.stream(self.copy_streams[k % len(self.copy_streams)]):
                self.peers[p][slot, self.rank].copy_(frame, non_blocking=True)
        for st in self.copy_streams:
            cur.wait_stream(st)
        self.handle.barrier(channel=slot)
        return self.buf[slot]

    def push_to(self, frame: torch.Tensor, slot: int, consumer: int) -> torch.Tensor:
        """Gather-to-one: copy this rank's frame into ``[slot, rank]`` of rank `consumer`'s buffer only (1 / world of the
        all-gather's traffic - for a pipeline whose frames are consumed by one rank each, e.g. round-robin encoders),
        then join the barrier; on `consumer`, ``view(slot)`` then holds the frames of all ranks."""
        cur = torch.cuda.current_stream()
        ready = torch.cuda.Event()
        ready.record(cur)
        st = self.copy_streams[0]
        st.wait_event(ready)
        with torch.cuda.stream(st):
            self.peers[consumer][slot, self.rank].copy_(frame, non_blocking=True)
        cur.wait_stream(st)
        self.handle.barrier(channel=slot)
        return self.buf[slot]

    def view(self, slot: int) -> torch.Tensor:
        return self.buf[slot]


def render_views_sharded(render_view: Callable[[int, torch.Tensor], None], n_views: int, frame_shape: Sequence[int],
                         device, dtype=torch.float32, gather_to_all: bool = True,
                         side_stream: Optional["torch.cuda.Stream"] = None) -> Optional[torch.Tensor]:
    """Render `n_views` views sharded by view and gather them in view order: returns [n_views, *frame_shape].

    `render_view(view_index, out)` renders one view into `out` (enqueued on the current stream).  When the number
    of views is not a multiple of the world size the last round is padded with the rank's last frame and cropped.
    """
    rank, world = _world()
    n_rounds = rounds(n_views, world)
    out = torch.empty((n_rounds * world, *frame_shape), dtype=dtype, device=device)
    mine = [torch.empty(tuple(frame_shape), dtype=dtype, device=device) for _ in range(2)]
    use_side = side_stream is not None and world > 1
    peer = PeerGather.create(frame_shape, device, slots=2) if (use_side and dtype == torch.float32) else None
    pending: List[Optional["torch.cuda.Event"]] = [None, None]
    for k in range(n_rounds):
        v = k * world + rank
        buf = mine[k % 2]
        if use_side and pending[k % 2] is not None:
            torch.cuda.current_stream().wait_event(pending[k % 2])  # the gather that read this buffer is done
        render_view(min(v, n_views - 1), buf)
        dst = out[k * world:(k + 1) * world]
        if world == 1:
            dst[0].copy_(buf)
        elif use_side:
            done = torch.cuda.Event(); done.record()
            side_stream.wait_event(done)
            with torch.cuda.stream(side_stream):
                if peer is not None:  # copy-engine gather into the symmetric ring, then into the result
                    dst.copy_(peer.push(buf, k % 2), non_blocking=True)
                    peer.handle.barrier(channel=2 + k % 2)  # nobody overwrites the slot before everyone has read it
                else:
                    _all_gather_stack(dst, buf)
                pending[k % 2] = torch.cuda.Event(); pending[k % 2].record()
        else:
            _all_gather_stack(dst, buf)
    if use_side:
        torch.cuda.current_stream().wait_stream(side_stream)
    return out[:n_views]


def render_frame_banded(render_band: Callable[[Tuple[int, int], torch.Tensor], None], row_loads: Sequence[float],
                        width: int, height: int, device, dtype=torch.float32) -> torch.Tensor:
    """Render one frame split into tile-row bands across the ranks and gather it: returns [H,W,3] on every rank.

    `render_band((row_begin, row_end), frame)` must fill the pixel rows of that band in the full-size `frame`.
    """
    rank, world = _world()
    tiles_y = (height + TILE - 1) // TILE
    assert len(row_loads) == tiles_y
    bands = partition_bands(row_loads, world)
    frame = torch.zeros((height, width, 3), dtype=dtype, device=device)
    render_band(bands[rank], frame)
    if world == 1:
        return frame
    max_px = max(band_pixel_rows(b, height)[1] - band_pixel_rows(b, height)[0] for b in bands)
    y0, y1 = band_pixel_rows(bands[rank], height)
    send = torch.zeros((max_px, width, 3), dtype=dtype, device=device)
    send[: y1 - y0].copy_(frame[y0:y1])
    gathered = torch.empty((world, max_px, width, 3), dtype=dtype, device=device)
    _all_gather_stack(gathered, send)
    return assemble_bands(gathered, bands, height)


# ----------------------------------------------------------------------------------------------
# GPU front-ends
# ----------------------------------------------------------------------------------------------
def _agree_on_overflow(renderer) -> int:
    """Sticky device status of this rank's frames, agreed across the ranks: 0 when every frame of every rank fit its
    pair capacity, else the capacity (max over ranks of the worst frame's need, +10 %) everybody must redraw with."""
    info = renderer.frame_info()
    need = info["n_dup_peak"] if info["status"] == 4 else 0     # 4 = GSR_ERR_DUP_OVERFLOW
    failed = info["status"] if info["status"] not in (0, 4) else 0
    if _world()[1] > 1:  # one MAX all-reduce of two words; every rank takes the same branch below (no rank left waiting)
        t = torch.tensor([need, failed], dtype=torch.int64,
                         device=renderer.device if dist.get_backend() == "nccl" else "cpu")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        need, failed = int(t[0].item()), int(t[1].item())
    if failed:
        renderer.clear_status()
        raise RuntimeError(f"libgsr_b200 frame status {failed} on this or another rank")
    if need:
        renderer.clear_status()
        return int(need * 1.1) + 1024
    return 0


def render_orbit(model, intr, views, degree: int = 3, renderer=None) -> torch.Tensor:
    """[V,H,W,3]: config 4 - all views of an orbit, sharded by view across the initialised process group.

    Frames are enqueued without per-frame status reads; the device-side status is sticky, so ONE read after the orbit
    (agreed across ranks with a MAX all-reduce of a single word) tells whether any frame of any rank overflowed its pair
    capacity - in that case every rank raises its capacity to the reported need and the orbit is redrawn once."""
    from .camera import camera_for
    from .render import default_renderer

    r = renderer or default_renderer()
    cams = [camera_for(intr, v, degree) for v in views]
    side = torch.cuda.Stream() if _world()[1] > 1 else None
    for attempt in range(2):
        out = render_views_sharded(lambda i, o: r.render(model, cams[i], out=o, check=False), len(views),
                                   (intr.height, intr.width, 3), r.device, side_stream=side)
        cap = _agree_on_overflow(r)
        if not cap:
            return out
        if attempt == 1:
            raise RuntimeError(f"orbit frames still overflow a pair capacity of {r.max_dup}")
        r.reserve(model.n, intr.width, intr.height, max_dup=cap)
    return out


def tile_row_loads(model, cam, renderer=None) -> List[float]:
    """(Gaussian, tile) pairs per tile row, from one preprocess + binning pass (identical on every rank)."""
    from . import render as R

    pre = R.preprocess(model, cam)
    n_pairs = int((pre["tiles_touched"] & 0x7FFFFFFF).sum().item())
    binned = R.bin_tiles(pre, cam.width, cam.height, max(1024, n_pairs))
    tiles_x, tiles_y = (cam.width + TILE - 1) // TILE, (cam.height + TILE - 1) // TILE
    rng = binned["tile_ranges"].to(torch.int64)
    per_tile = (rng[:, 1] - rng[:, 0]).reshape(tiles_y, tiles_x)
    return [float(v) for v in per_tile.sum(dim=1).tolist()]


def render_banded(model, cam, renderer=None, row_loads: Optional[Sequence[float]] = None) -> torch.Tensor:
    """[H,W,3]: config 5 - one frame split into tile-row bands across the initialised process group."""
    from .render import default_renderer

    r = renderer or default_renderer()
    if row_loads is None:
        row_loads = tile_row_loads(model, cam, r)
    for attempt in range(2):
        out = render_frame_banded(lambda band, frame: r.render(model, cam, out=frame, tile_rows=band, check=False),
                                  row_loads, cam.width, cam.height, r.device)
        cap = _agree_on_overflow(r)  # sticky status, agreed across ranks (every rank must take the same branch)
        if not cap:
            return out
        if attempt == 1:
            raise RuntimeError(f"band still overflows a pair capacity of {r.max_dup}")
        r.reserve(model.n, cam.width, cam.height, max_dup=cap)
    return out
