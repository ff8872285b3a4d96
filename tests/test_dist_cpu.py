"""Multi-rank host logic on CPU: partitioning functions and the gather paths over a 2-rank gloo group."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gsr_b200 import dist as gd


def test_views_round_robin_covers_everything_once():
    for n, w in ((256, 8), (7, 2), (3, 4), (1, 1)):
        seen = sorted(v for r in range(w) for v in gd.views_of_rank(n, r, w))
        assert seen == list(range(n))
        assert gd.rounds(n, w) == max(len(gd.views_of_rank(n, r, w)) for r in range(w))


@pytest.mark.parametrize("loads,world", [([1] * 68, 8), ([0, 0, 5, 100, 3, 0, 0, 1], 4), ([10, 1, 1, 1, 1, 1, 1, 10], 2),
                                         ([0] * 9, 4), ([3, 4], 4), ([7], 1), (list(range(1, 136)), 8)])
def test_partition_bands_is_a_balanced_contiguous_cover(loads, world):
    bands = gd.partition_bands(loads, world)
    assert len(bands) == world and bands[0][0] == 0 and bands[-1][1] == len(loads)
    assert all(bands[k][1] == bands[k + 1][0] for k in range(world - 1))
    assert all(b <= e for b, e in bands)
    if len(loads) >= world:
        assert all(e > b for b, e in bands)
    total = sum(loads)
    if total and len(loads) >= 4 * world:  # balance: no band exceeds its fair share by more than the heaviest row
        assert max(sum(loads[b:e]) for b, e in bands) <= total / world + max(loads) + 1e-9


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_views, H, W):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # --- view sharding: a stub renderer paints each frame with its view index ---
        rendered = []

        def render_view(i, out):
            rendered.append(i)
            out.fill_(float(i))

        frames = gd.render_views_sharded(render_view, n_views, (H, W, 3), "cpu")
        assert tuple(frames.shape) == (n_views, H, W, 3)
        for v in range(n_views):
            assert torch.all(frames[v] == float(v)), (rank, v)
        expect = gd.views_of_rank(n_views, rank, world)
        assert rendered[: len(expect)] == expect  # this rank rendered exactly its share (plus padding of the last round)

        # --- band sharding: every rank paints its band's pixel rows with row indices ---
        tiles_y = (H + 15) // 16
        loads = [float((r * 7) % 5 + 1) for r in range(tiles_y)]

        def render_band(band, frame):
            y0, y1 = gd.band_pixel_rows(band, H)
            frame[y0:y1] = torch.arange(y0, y1, dtype=torch.float32)[:, None, None]

        img = gd.render_frame_banded(render_band, loads, W, H, "cpu")
        assert tuple(img.shape) == (H, W, 3)
        assert torch.equal(img[:, 0, 0], torch.arange(H, dtype=torch.float32))
    finally:
        dist.destroy_process_group()


class _FakeRenderer:
    """Stands in for gsr_b200.render.Renderer in the overflow agreement: sticky status words held on the host."""

    def __init__(self, status, peak):
        self.device, self.info, self.max_dup, self.cleared = "cpu", dict(status=status, n_dup_peak=peak), 1000, 0

    def frame_info(self):
        return dict(self.info)

    def clear_status(self):
        self.info, self.cleared = dict(status=0, n_dup_peak=0), self.cleared + 1


def _overflow_worker(rank, world, port):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # nobody overflowed: nothing to do, status untouched
        r = _FakeRenderer(0, 700)
        assert gd._agree_on_overflow(r) == 0 and r.cleared == 0
        # only rank 1 overflowed (a frame in the middle of its sequence needed 5000 pairs): EVERY rank must learn the
        # same capacity, or the ranks would take different branches around the collective that follows
        r = _FakeRenderer(4 if rank == 1 else 0, 5000 if rank == 1 else 800)
        cap = gd._agree_on_overflow(r)
        assert cap == int(5000 * 1.1) + 1024 and r.cleared == 1 and r.frame_info()["status"] == 0
        # both overflowed: the larger need wins
        r = _FakeRenderer(4, 3000 + 1000 * rank)
        assert gd._agree_on_overflow(r) == int(4000 * 1.1) + 1024
        # a failure other than an overflow on ONE rank is raised on EVERY rank (nobody is left waiting in a collective)
        r = _FakeRenderer(3 if rank == 0 else 0, 0)
        try:
            gd._agree_on_overflow(r)
            raise AssertionError("a failed frame on rank 0 went unnoticed on rank %d" % rank)
        except RuntimeError as exc:
            assert "status 3" in str(exc) and r.cleared == 1
    finally:
        dist.destroy_process_group()


def test_overflow_status_is_agreed_across_ranks():
    mp.spawn(_overflow_worker, args=(2, _free_port()), nprocs=2, join=True)


@pytest.mark.parametrize("n_views,H,W", [(5, 40, 24), (4, 33, 16)])
def test_two_rank_gloo_gathers(n_views, H, W):
    mp.spawn(_worker, args=(2, _free_port(), n_views, H, W), nprocs=2, join=True)


def test_peer_gather_declines_without_nccl_cuda():
    """PeerGather is the CUDA / NCCL fast path; on CPU (or without a process group) `create` returns None and the
    callers use the collective - the path the gloo tests above exercise."""
    from gsr_b200 import dist as gd

    assert gd.PeerGather.create((4, 4, 3), "cpu") is None
