#!/usr/bin/env python
"""Benchmark of the render path (BASELINE.json: MPix/s & frames/s, 3M Gaussians SH3 @1080p).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one frame per GPU.  N=1: config C3 (3M synth_v1 Gaussians, SH degree 3, 1920x1080, front view), the
Gaussians resident in HBM as the PREPARED model (gsr_prepare runs once at load, like the PLY unpack; its cost is
reported in `stages_ms.prepare_once` and is paid inside every step of the host-buffer `e2e` path instead).  N>1: the same
scene replicated on every GPU, rank r renders orbit view (step*N + r) of the 256-view orbit (config C4: view sharding),
finished frames gathered on a side stream - weak scaling (one frame per GPU per step).

Prints ONE JSON line (rank 0).  Keys beyond the driver's contract:
  fps, stages_ms, roofline (blend = dominant kernel, FP32-FMA bound), roofline_stages (preprocess / binning, HBM bound,
  against the SURVEY's algorithmic bytes AND against the bytes this design moves), cpu_baseline, e2e (host tensors in,
  host image out, through the C-level host pipeline), clocks; at N>1 also gather_modes (none / nccl / p2p / p2p-rr
  frames/s of the same loop) and banded_c5 (BASELINE config 5: 6M Gaussians @3840x2160 split into tile-row bands over the
  N ranks, forward + backward).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "torch-gaussian-splatting-rasterizer_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

# The contract is ONE JSON line on stdout.  Native libraries print there too (NCCL's version banner comes from C code,
# whatever NCCL_DEBUG says on this image), so file descriptor 1 points at stderr while the benchmark runs and is restored
# for the final line only.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    os.dup2(_REAL_STDOUT, 1)
    print(json.dumps(line), flush=True)
    os.dup2(2, 1)


import numpy as np  # noqa: E402
import torch  # noqa: E402

FLOP_PER_EVAL = 24.0           # one Gaussian-pixel evaluation
PREPROCESS_BYTES = {0: 104.0, 1: 140.0, 2: 200.0, 3: 284.0}  # SURVEY: (11 + 3(d+1)^2)*4 in + 48 out, per Gaussian
KERNELS_PER_FRAME = 10         # preprocess, 4 depth-sort passes (identity passes return at once), scan+emit, 2 tile-sort
#                                passes, tile ranges, blend (+ 2 memsets, not counted)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            d = json.load(fh)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def load_profile(name):
    path = os.path.join(ROOT, "profiles", name)
    if os.path.exists(path):
        with open(path) as fh:
            return json.load(fh)
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for nme, val in zip(names, r[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nme)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def build_scene(n_override=None, config="C3"):
    from gsr_b200 import scenes

    c = scenes.CONFIGS[config]
    n = c["n"] if n_override is None else n_override
    t0 = time.time()
    cloud = scenes.synth_v1(n, c["seed"])
    intr = scenes.Intrinsics.synthetic(c["width"], c["height"])
    log(f"[bench] generated {n} Gaussians ({config}) in {time.time() - t0:.1f}s")
    return cloud, intr, c["degree"]


# ----------------------------------------------------------------------------------------------
# CPU arms (reported baselines, not targets).  Both keep the reference's execution model: vectorised per-Gaussian set-up,
# then ONE eager-torch blend step per Gaussian in global depth order (/root/reference/rasterize.py:436-446); the loop is
# Python-dispatch bound and serial, so the sample is a prefix of the loop, extrapolated linearly over the blended count.
# ----------------------------------------------------------------------------------------------
def cpu_port_baseline(cloud, intr, degree, budget_s=20.0, quiet=False):
    """oracle/torch_port.py (kind "port"): test infrastructure, timed here as the checker's cost - never the product."""
    from gsr_b200 import camera, scenes
    from oracle import torch_port

    cam = camera.camera_for(intr, scenes.front_view(0), degree)
    cores = os.cpu_count() or 1
    tens = torch_port.cloud_tensors(cloud)
    best = None
    for nt in sorted({1, min(4, cores), max(1, cores - 1)}):  # eager ops: pick the thread count that is fastest here
        torch.set_num_threads(nt)
        with torch.no_grad():
            t0 = time.time()
            pre = torch_port.preprocess(*tens, cam)
            t_pre = time.time() - t0
            W, H = cam.width, cam.height
            screen, T = torch.zeros((W, H, 3)), torch.ones((W, H))
            todo = [g for g in pre["order"].tolist()[: 200_000] if not bool(pre["skip"][g])][:150]
            t0 = time.time()
            for g in todo:
                torch_port.blend_step(g, pre, screen, T, False)
            per = (time.time() - t0) / max(1, len(todo))
        if best is None or per < best[1]:
            best = (nt, per, t_pre)
    nt, per, t_pre = best
    torch.set_num_threads(nt)
    sample = max(200, int(budget_s / max(per, 1e-6)))
    t0 = time.time()
    out = torch_port.render(*tens, cam, max_blended=sample)
    t_total = time.time() - t0
    blended, blendable = out["n_blended"], out["n_blendable"]
    t_loop = max(t_total - t_pre, 1e-9)
    est_frame_s = t_pre + t_loop * (blendable / max(1, blended))
    mpix = intr.width * intr.height / 1e6 / est_frame_s
    if not quiet:
        log(f"[bench] cpu port: {nt} threads, preprocess {t_pre:.2f}s, {blended} of {blendable} blendable Gaussians in "
            f"{t_loop:.1f}s -> est. {est_frame_s:.0f}s/frame")
    return {"value": mpix, "unit": "MPix/s", "cores": nt, "kind": "port",
            "sample": f"oracle/torch_port.py on config C3: full vectorised preprocess + first {blended} of {blendable} "
                      f"blendable Gaussians in depth order, loop time extrapolated linearly (extrapolated)",
            "est_seconds_per_frame": est_frame_s, "host_cpus": cores}


_REF_CHILD = r"""
import json, logging, os, sys, tempfile, time
root, n, budget = sys.argv[1], int(sys.argv[2]), float(sys.argv[3])
sys.path[:0] = [os.path.join(root, "torch-gaussian-splatting-rasterizer_b200"), root]
os.environ.setdefault("TQDM_DISABLE", "1")
logging.disable(logging.CRITICAL)
from gsr_b200 import scenes
from oracle import ref_harness
c = scenes.CONFIGS["C3"]
cloud = scenes.synth_v1(n, c["seed"])
intr = scenes.Intrinsics.synthetic(c["width"], c["height"])
with tempfile.TemporaryDirectory() as d:
    t0 = time.time()
    scenes.write_scene_dir(d, cloud, intr, [scenes.front_view(0)])
    t_write = time.time() - t0
    probe = ref_harness.time_reference(d, 0, max_blended=2000)         # cost per blended Gaussian on this host
    per = probe["loop_s"] / max(1, probe["blended"])
    sample = max(200, int(budget / max(per, 1e-6)))
    out = ref_harness.time_reference(d, 0, max_blended=sample)
out["scene_write_s"] = t_write
print("REF_RESULT " + json.dumps(out), flush=True)
"""


def staged_reference_dir():
    for cand in (os.environ.get("GSR_REFERENCE_DIR"), os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if cand and os.path.isfile(os.path.join(cand, "rasterize.py")):
            return cand
    return None


def cpu_reference_baseline(n, intr, budget_s=20.0):
    """The UNMODIFIED reference (staged copy under the git-ignored baseline/_ref) through its own click command, in a
    child process with the GPUs hidden (rasterize.py:324 would otherwise pick CUDA) and the reference's own thread
    setting (os.cpu_count()-1, :323); left after a bounded number of loop iterations."""
    ref_dir = staged_reference_dir()
    if ref_dir is None:
        return None
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", GSR_REFERENCE_DIR=ref_dir, TQDM_DISABLE="1")
    try:
        res = subprocess.run([sys.executable, "-c", _REF_CHILD, ROOT, str(n), str(budget_s)], env=env, capture_output=True,
                             text=True, timeout=900)
    except subprocess.TimeoutExpired:
        return None
    line = [l for l in res.stdout.splitlines() if l.startswith("REF_RESULT ")]
    if res.returncode != 0 or not line:
        log("[bench] reference child failed:", res.stderr[-600:])
        return None
    o = json.loads(line[-1][len("REF_RESULT "):])
    est = o["setup_s"] + o["loop_s"] * (o["blendable"] / max(1, o["blended"]))
    log(f"[bench] unmodified reference: set-up {o['setup_s']:.1f}s, {o['blended']} of {o['blendable']} blended Gaussians in "
        f"{o['loop_s']:.1f}s -> est. {est:.0f}s/frame ({o['threads']} threads)")
    return {"value": intr.width * intr.height / 1e6 / est, "unit": "MPix/s", "cores": o["threads"], "kind": "reference",
            "sample": f"unmodified reference (baseline/_ref, run_rasterization.callback, GPUs hidden) on config C3: PLY load + "
                      f"its whole vectorised set-up + first {o['blended']} of {o['blendable']} iterations of its blend loop "
                      f"that are not skipped, loop time extrapolated linearly (extrapolated)",
            "est_seconds_per_frame": est, "host_cpus": os.cpu_count()}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cloud, intr, degree = build_scene(args.n)
    b = cpu_reference_baseline(cloud.n, intr, budget_s=args.cpu_budget) or cpu_port_baseline(cloud, intr, degree, budget_s=args.cpu_budget)
    line = {"impl": "reference", "metric": "MPix/s, 3M-Gaussian SH3 @1080p render",
            "what": ("the unmodified reference's CPU render (staged copy in baseline/_ref)" if b["kind"] == "reference" else
                     "the reference's CPU render timed through the oracle's torch port (no staged reference copy on this box)"),
            "value": b["value"], "unit": "MPix/s", "fps": b["value"] * 1e6 / (intr.width * intr.height),
            "n_gpus": args.gpus, "steps": 1, "warmup": 0, "ms_per_step": b["est_seconds_per_frame"] * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": ("C3: 3M synth_v1 Gaussians (seed 3), SH degree 3, 1920x1080, front view"
                                    if args.gpus <= 1 else
                                    "C4: same 3M-Gaussian scene, 256-view orbit (the CPU render of ONE view is timed, on "
                                    "rank 0; views are independent and cost the same)") if args.n is None
                       else f"C3 scene with n={args.n}"},
            "cpu_baseline": b, "e2e": {"value": b["value"], "unit": "MPix/s", "h2d_bytes_per_step": 0,
                                       "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    emit(line)


# ----------------------------------------------------------------------------------------------
# config 5 inside the multi-GPU run: one 6M-Gaussian 3840x2160 frame split into tile-row bands, forward + backward
# ----------------------------------------------------------------------------------------------
def banded_c5(world, rank, n_override=None, reps=5):
    import torch.distributed as dist
    from gsr_b200 import camera, dist as gd, render as R, scenes

    cloud, intr, degree = build_scene(n_override, "C5")
    W, H = intr.width, intr.height
    cam = camera.camera_for(intr, scenes.front_view(0), degree)
    model = R.GaussianModel.from_cloud(cloud)
    model.prepare()
    dev = model.device
    G = torch.from_numpy(np.random.default_rng(55).uniform(-1, 1, (H, W, 3)).astype(np.float32)).to(dev)
    loads = gd.tile_row_loads(model, cam)
    bands = gd.partition_bands(loads, world)
    band = bands[rank]
    y0, y1 = gd.band_pixel_rows(band, H)
    max_px = max(gd.band_pixel_rows(b, H)[1] - gd.band_pixel_rows(b, H)[0] for b in bands)
    r = R.Renderer()
    final_T = torch.empty((H, W), device=dev)
    n_contrib = torch.zeros((H, W), dtype=torch.int32, device=dev)
    frame = torch.zeros((H, W, 3), device=dev)
    send = torch.zeros((max_px, W, 3), device=dev)
    gathered = torch.empty((world, max_px, W, 3), device=dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]

    def sync_all():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    def timed(fn):  # max over ranks of the mean device time of `reps` calls
        for _ in range(2):
            fn()
        sync_all()
        ev[0].record()
        for _ in range(reps):
            fn()
        ev[1].record()
        torch.cuda.synchronize()
        t = torch.tensor([ev[0].elapsed_time(ev[1]) / reps], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    out = {"config": "C5: 6M synth_v1 Gaussians (seed 5), SH degree 3, 3840x2160, one frame", "n_gaussians": cloud.n,
           "bands_tile_rows": [list(b) for b in bands], "n_ranks": world}
    # -- single GPU reference numbers (every rank does the same full frame: per-GPU time, no communication) --
    r.render(model, cam, out=frame, final_T=final_T, n_contrib=n_contrib)  # checked: establishes the capacity
    out["n_dup_full"] = r.frame_info()["n_dup"]

    def full_fwd():
        r.render(model, cam, out=frame, final_T=final_T, n_contrib=n_contrib, check=False)

    def full_fwd_bwd():
        full_fwd()
        return r.backward(model, cam, final_T, n_contrib, G)

    t1_fwd = timed(full_fwd)
    t1_fb = timed(full_fwd_bwd)
    full_image = frame.clone()
    full_grads = [g.clone() for g in full_fwd_bwd()]
    # -- banded: this rank's band only, bands all-gathered (NCCL, the path's only collective) --
    rb = R.Renderer()
    rb.reserve(model.n, W, H, max_dup=r.max_dup)
    rb.adaptive_capacity = False
    bframe = torch.zeros((H, W, 3), device=dev)

    def band_fwd():
        rb.render(model, cam, out=bframe, final_T=final_T, n_contrib=n_contrib, tile_rows=band, check=False)
        send[: y1 - y0].copy_(bframe[y0:y1])
        dist.all_gather_into_tensor(gathered.view(-1), send.view(-1))

    def band_fwd_bwd():
        band_fwd()
        return rb.backward(model, cam, final_T, n_contrib, G, tile_rows=band)

    tb_fwd = timed(band_fwd)
    tb_fb = timed(band_fwd_bwd)
    part = band_fwd_bwd()
    torch.cuda.synchronize()
    status = rb.frame_info()["status"]
    image = gd.assemble_bands(gathered, bands, H)
    bit_identical = bool(torch.equal(image, full_image))
    # sum of the per-band partial gradients vs the single-GPU gradient (validation only; the product ships partials).
    # Reduced over the ranks with all_gather + a local sum - NCCL stays a gather on this path.
    rel = {}
    for name, p, f in zip(("d_xyz", "d_log_scales", "d_rots", "d_opacity", "d_sh"), part, full_grads):
        if name == "d_sh":  # 1.15 GB per rank: validate the DC plane and one degree-3 plane of every channel
            p, f = p[:, [0, 12], :].contiguous(), f[:, [0, 12], :].contiguous()
        p = p.contiguous()
        allp = torch.empty((world, *p.shape), device=dev)
        dist.all_gather_into_tensor(allp.view(-1), p.view(-1))
        tot = allp.sum(dim=0)
        scale = max(float(f.abs().max()), 1e-12)
        rel[name] = float((tot - f).abs().max()) / scale
        del allp, tot
    out.update({"full_frame_ms": {"forward": t1_fwd, "forward_backward": t1_fb},
                "banded_ms": {"forward_incl_allgather": tb_fwd, "forward_backward": tb_fb},
                "speedup": {"forward": t1_fwd / tb_fwd, "forward_backward": t1_fb / tb_fb},
                "forward_bit_identical": bit_identical, "band_status": status,
                "grad_sum_of_partials_max_rel_err": rel,
                "note": "every rank preprocesses all Gaussians (replicated on purpose: cheaper than exchanging them), bins "
                        "and blends only its band; times are max over ranks of CUDA-event means; backward yields per-band "
                        "partial gradients (their sum is the full gradient - float atomics, so equal to rounding)"})
    del model, cloud
    torch.cuda.empty_cache()
    return out


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=None, help="override the Gaussian count (debugging only)")
    ap.add_argument("--orbit", action="store_true", help="N=1 only: render the 256-view orbit of the N>1 workload (to separate view cost from scaling loss)")
    ap.add_argument("--gather", choices=["none", "nccl", "p2p", "p2p-rr"], default="p2p-rr",
                    help="N>1: how finished frames are gathered in the timed loop (every mode is also measured and reported in "
                         "gather_modes).  Default: each step's N frames go to ONE consumer rank, round-robin, by peer DMA - "
                         "what a consumer of frames (an encoder per rank) needs, 1/N of an all-gather's traffic; measured at "
                         "N=8 (profiles/r02_bench_8gpu_allgather.json): no gather 9976 fps, this 9886, peer-DMA all-gather "
                         "9233, NCCL all_gather 8322 - the scaling loss of round 1 was the all-gather's HBM/NVLink traffic")
    ap.add_argument("--graph", action="store_true", help="replay each frame from a captured CUDA graph")
    ap.add_argument("--ring", type=int, default=8, help="frame buffers per rank (N>1): how far a rank may run ahead of the gather")
    ap.add_argument("--cpu-budget", type=float, default=20.0)
    ap.add_argument("--e2e-steps", type=int, default=12)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-banded", action="store_true", help="N>1: skip the config-5 sub-benchmark")
    ap.add_argument("--no-gather-modes", action="store_true", help="N>1: skip the secondary gather-mode loops")
    ap.add_argument("--banded-n", type=int, default=None, help="override the Gaussian count of the config-5 sub-benchmark")
    ap.add_argument("--streams", type=int, default=3,
                    help="frames in flight per GPU: successive steps alternate between this many renderers/streams")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    if args.warmup < 3:
        args.warmup = 3

    import torch.distributed as dist
    from gsr_b200 import _cabi, camera, render as R, scenes

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    _cabi.require_cuda()
    hbm_peak, peak_src = load_peaks()

    cloud, intr, degree = build_scene(args.n)
    W, H = intr.width, intr.height
    model = R.GaussianModel.from_cloud(cloud)
    pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    model.prepare()
    pe0.record(); model.prepare(force=True); pe1.record(); torch.cuda.synchronize()
    t_prepare = pe0.elapsed_time(pe1)
    renderer = R.Renderer()
    n_streams = max(1, args.streams)
    # every lane has its own stream (graph capture needs non-default streams); the main stream only forks and joins
    lanes = [(renderer, torch.cuda.Stream())] + [(R.Renderer(), torch.cuda.Stream()) for _ in range(n_streams - 1)]
    if world == 1 and not args.orbit:
        views = [scenes.front_view(0)]
    else:
        views = scenes.orbit_views(256)
    cams = [camera.camera_for(intr, v, degree) for v in views]

    def cam_index(step):
        return (step * world + rank) % len(cams)

    frame = torch.empty((H, W, 3), dtype=torch.float32, device="cuda")
    gathered = torch.empty((world, H, W, 3), dtype=torch.float32, device="cuda") if world > 1 else None
    comm = torch.cuda.Stream() if world > 1 else None
    peer = None
    if world > 1:
        from gsr_b200.dist import PeerGather
        peer = PeerGather.create((H, W, 3), "cuda", slots=4)
    n_buf = max(args.ring, n_streams)  # frames a rank may run ahead of the slowest rank's gather
    frames = [torch.empty_like(frame) for _ in range(n_buf)]
    gathered_ev = [None] * n_buf
    graphs = {}

    def enqueue_frame(rend, cam_i, buf_i):
        rend.render(model, cams[cam_i], out=frames[buf_i], check=False)

    def render_step(lane_i, cam_i, buf_i):
        rend, _ = lanes[lane_i]
        if not args.graph:
            return enqueue_frame(rend, cam_i, buf_i)
        key = (lane_i, cam_i, buf_i)
        g = graphs.get(key)
        if g is None:  # capture the frame's 2 memsets + 10 kernels once; the camera is baked into the kernel arguments
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=torch.cuda.current_stream()):
                enqueue_frame(rend, cam_i, buf_i)
            graphs[key] = g
        g.replay()

    def make_step(mode):
        def step(i):
            # Successive frames are independent: frame i runs on lane i % n_streams (own renderer, workspace and stream),
            # so the latency-bound binning kernels of one frame overlap the issue-bound blend of another.
            _, strm = lanes[i % n_streams]
            bi = i % n_buf
            with torch.cuda.stream(strm):
                if gathered_ev[bi] is not None:  # the gather that last read this buffer must have finished
                    strm.wait_event(gathered_ev[bi])
                render_step(i % n_streams, cam_index(i), bi)
                if world > 1 and mode != "none":  # the only collective: gather the finished frames (overlaps the next render)
                    done = torch.cuda.Event(); done.record(strm)
                    comm.wait_event(done)
                    with torch.cuda.stream(comm):
                        if mode == "p2p" and peer is not None:      # copy-engine all-gather over NVLink: no SM time
                            peer.push(frames[bi], i % peer.slots)
                        elif mode == "p2p-rr" and peer is not None:  # each step's frames go to ONE consumer rank, round-robin
                            peer.push_to(frames[bi], i % peer.slots, i % world)
                        else:
                            dist.all_gather_into_tensor(gathered.view(-1), frames[bi].view(-1))
                        gathered_ev[bi] = torch.cuda.Event(); gathered_ev[bi].record(comm)
        return step

    def fork():   # the lanes start after everything already queued on the main stream
        ev = torch.cuda.Event(); ev.record()
        for _, strm in lanes:
            strm.wait_event(ev)

    def join():   # the main stream waits for every lane (and the gathers)
        for _, strm in lanes:
            torch.cuda.current_stream().wait_stream(strm)
        if world > 1:
            torch.cuda.current_stream().wait_stream(comm)

    # reserve workspaces (one checked frame per lane also validates max_dup); orbit views need up to ~5 % more pairs than
    # the front view, the capacity keeps 50 % head-room, and an overflow on any frame is caught by the sticky status
    for rend, strm in lanes:
        with torch.cuda.stream(strm):
            rend.render(model, cams[0], out=frame, check=True)
    torch.cuda.synchronize()
    info0 = renderer.frame_info()
    log(f"[bench] rank {rank}: n_visible {info0['n_visible']} n_dup {info0['n_dup']} max_dup {renderer.max_dup}")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_loop(mode, steps, warmup, first):
        step = make_step(mode)
        for k in range(n_buf):
            gathered_ev[k] = None
        fork()
        for i in range(warmup):
            step(first + i)
        join()
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        fork()
        for i in range(steps):
            step(first + warmup + i)
        join()
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    main_mode = args.gather if world > 1 else "none"
    if main_mode.startswith("p2p") and peer is None and world > 1:
        main_mode = "nccl"
    with ClockSampler(local) as clocks:
        ms = timed_loop(main_mode, args.steps, args.warmup, 0)
    ms_per_step = ms / args.steps
    frames_per_s = world * args.steps / (ms / 1e3)
    mpix_s = frames_per_s * W * H / 1e6
    status_after = max(r.frame_info()["status"] for r, _ in lanes)
    gather_kind = {"none": None, "nccl": "NCCL all_gather", "p2p": "peer-DMA all-gather (symmetric memory)",
                   "p2p-rr": "peer-DMA gather to one consumer rank per step, round-robin"}[main_mode]

    line = {"metric": "MPix/s, 3M-Gaussian SH3 @1080p render", "value": mpix_s, "unit": "MPix/s", "fps": frames_per_s,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": ("C3: 3M synth_v1 Gaussians (seed 3), SH degree 3, 1920x1080, front view" if world == 1
                                    else f"C4: same 3M-Gaussian scene replicated per GPU, 256-view orbit sharded by view "
                                         f"(rank r renders view step*{world}+r), frames gathered: {gather_kind}"),
                       "n_gaussians": cloud.n, "width": W, "height": H, "sh_degree": degree,
                       "l2_policy": "inputs (720 MB prepared model) exceed the 126 MB L2; no flush needed",
                       "parallelism": f"view-sharded x{world}" if world > 1 else "single GPU",
                       "frames_in_flight": n_streams, "cuda_graph": bool(args.graph),
                       "resident": "prepared model (gsr_prepare once at load, see stages_ms.prepare_once)"},
            "gpu_launches": KERNELS_PER_FRAME * args.steps, "frame_status": status_after}

    # ---- secondary loops (N>1): the same frames with every gather mode, so the cost of the collective is attributed ----
    if world > 1 and not args.no_gather_modes:
        modes = {}
        k = max(20, min(args.steps, 100))
        for mode in ("none", "nccl", "p2p", "p2p-rr"):
            if mode.startswith("p2p") and peer is None:
                continue
            t = timed_loop(mode, k, 3, 1000)
            modes[mode] = {"fps": world * k / (t / 1e3), "ms_per_step": t / k}
        modes["what"] = ("frames/s of the timed loop with: no gather (pure render + launch rate), NCCL all_gather (the collective "
                         "north_star names), peer-DMA all-gather, peer-DMA gather-to-one-consumer; value uses --gather " + main_mode)
        line["gather_modes"] = modes

    if rank == 0:
        # ---- per-stage timing with CUDA events on the launching stream: preallocated buffers, one C call per stage ----
        cam = cams[0]
        st = R.Stages(model, W, H, renderer.max_dup)
        count = torch.zeros(8, dtype=torch.int32, device="cuda")
        st.preprocess(cam); st.bin(); st.blend(count)
        torch.cuda.synchronize()
        cnt = count.cpu().numpy().view(np.uint32)
        warp_evals = int(cnt[4]) | (int(cnt[5]) << 32)
        pixel_evals = int(cnt[6]) | (int(cnt[7]) << 32)
        binfo = st.info.cpu().numpy().view(np.uint32)
        n_dup, n_vis = int(binfo[2]), int(binfo[1])
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        acc = np.zeros(3)
        reps = max(5, min(args.steps, 20))
        for _ in range(3):
            st.preprocess(cam); st.bin(); st.blend()
        torch.cuda.synchronize()
        for _ in range(reps):
            evs[0].record(); st.preprocess(cam)
            evs[1].record(); st.bin()
            evs[2].record(); st.blend()
            evs[3].record(); torch.cuda.synchronize()
            acc += [evs[k].elapsed_time(evs[k + 1]) for k in range(3)]
        t_pre, t_bin, t_blend = (acc / reps).tolist()
        # single frame, alone on the GPU (latency)
        evs[0].record()
        for _ in range(reps):
            renderer.render(model, cam, out=frame, check=False)
        evs[1].record(); torch.cuda.synchronize()
        t_single = evs[0].elapsed_time(evs[1]) / reps
        # forward with contributor tracking + backward (north_star item 4), same frame
        fT = torch.empty((H, W), device="cuda")
        nC = torch.zeros((H, W), dtype=torch.int32, device="cuda")
        Gimg = torch.rand((H, W, 3), device="cuda")
        rb_ = R.Renderer()
        rb_.render(model, cam, out=frame, final_T=fT, n_contrib=nC, prepared=False)
        for _ in range(2):
            rb_.render(model, cam, out=frame, final_T=fT, n_contrib=nC, check=False, prepared=False)
            rb_.backward(model, cam, fT, nC, Gimg)
        torch.cuda.synchronize()
        evs[0].record()
        for _ in range(5):
            rb_.render(model, cam, out=frame, final_T=fT, n_contrib=nC, check=False, prepared=False)
        evs[1].record()
        for _ in range(5):
            rb_.backward(model, cam, fT, nC, Gimg)
        evs[2].record(); torch.cuda.synchronize()
        t_fwd_tracked, t_bwd = evs[0].elapsed_time(evs[1]) / 5, evs[1].elapsed_time(evs[2]) / 5
        del rb_, fT, nC, Gimg
        peak_t, mhz = _cabi.ctypes.c_double(), _cabi.ctypes.c_double()
        _cabi.call("gsr_fp32_fma_peak", _cabi.ctypes.byref(peak_t), _cabi.ctypes.byref(mhz))
        fp32_peak = peak_t.value
        tiles = ((W + 15) // 16) * ((H + 15) // 16)
        passes = -(-(32 + int(np.ceil(np.log2(tiles)))) // 8)
        bin_bytes_survey = 8 * cloud.n + n_dup * (28 + 24 * passes) + 8 * tiles
        # what THIS design moves: first depth pass reads key+flag of all n and writes the visible pairs,
        # 3 more passes over the visible (key,id) pairs, scan+emit gathers a 16-byte span per visible Gaussian and writes
        # the pairs, 2 tile passes over the pairs (the last writes ids only... keys too), ranges read the tile ids
        bin_bytes_design = 8 * cloud.n + 8 * n_vis + 3 * 16 * n_vis + (4 + 16) * n_vis + 8 * n_dup + 2 * 16 * n_dup + 4 * n_dup + 8 * tiles
        pre_bytes = PREPROCESS_BYTES[degree] * cloud.n
        pre_bytes_design = (240.0 + 24.0) * cloud.n + 48.0 * n_vis  # prepared rows in; key, flags, span out; records of the visible
        blend_tflops = warp_evals * FLOP_PER_EVAL / (t_blend * 1e-3) / 1e12
        line["stages_ms"] = {"preprocess": t_pre, "binning": t_bin, "blend": t_blend, "single_frame": t_single,
                             "prepare_once": t_prepare, "forward_for_training": t_fwd_tracked, "backward": t_bwd,
                             "note": "one C-ABI call per stage on preallocated buffers, CUDA events on the launching stream; "
                                     "single_frame = one whole frame alone on the GPU (latency); prepare_once = gsr_prepare, "
                                     "paid once per model (at load), not per frame; forward_for_training = the raw-parameter "
                                     "frame with contributor tracking, backward = gsr_render_bwd of that frame (blend "
                                     "backward + preprocess / SH backward, gradients of all 59 parameters per Gaussian)"}
        prof = load_profile("r02_kernel_dram.json") or {}
        bl = prof.get("blend_kernel", {})
        line["roofline"] = {"kernel": "blend_kernel<4,0,0,0,2>", "bound": "fp32_fma", "achieved": blend_tflops, "peak": fp32_peak,
                            "unit": "TFLOP/s", "frac": blend_tflops / fp32_peak if fp32_peak else None,
                            "traffic": bl.get("dram_bytes"),
                            "traffic_static": {"static": True, "kernel": bl.get("kernel"), "source": bl.get("source"),
                                               "note": "DRAM bytes per launch from the committed ncu --set full capture of this "
                                                       "kernel (not measured in this run); algorithmic memory side is "
                                                       "n_dup*52 + 12*W*H - the kernel is issue-bound"},
                            "peak_source": f"FMA micro-benchmark in this run (gsr_fp32_fma_peak, implies {mhz.value:.0f} MHz)",
                            "evals_warp_granular": warp_evals, "evals_live_pixels": pixel_evals,
                            "gevals_per_s": warp_evals / (t_blend * 1e-3) / 1e9, "flop_per_eval": FLOP_PER_EVAL}

        def hbm(bytes_, t_ms):
            return bytes_ / (t_ms * 1e-3) / 1e9

        line["roofline_stages"] = {
            "preprocess": {"kernel": "preprocess_prepared_kernel<3>", "bound": "hbm", "achieved": hbm(pre_bytes, t_pre),
                           "peak": hbm_peak, "unit": "GB/s", "frac": hbm(pre_bytes, t_pre) / hbm_peak, "peak_source": peak_src,
                           "algorithmic_bytes": pre_bytes,
                           "design_bytes": pre_bytes_design, "frac_design_bytes": hbm(pre_bytes_design, t_pre) / hbm_peak,
                           "traffic": prof.get("preprocess_prepared_kernel", {}).get("dram_bytes"),
                           "note": "algorithmic = SURVEY 8(d): 284 B/Gaussian; design = 240 B prepared rows in + 24 B out per "
                                   "Gaussian + a 48 B record per visible one; traffic = ncu dram bytes per launch (static)"},
            "binning": {"bound": "hbm", "achieved": hbm(bin_bytes_survey, t_bin), "peak": hbm_peak, "unit": "GB/s",
                        "frac": hbm(bin_bytes_survey, t_bin) / hbm_peak, "peak_source": peak_src,
                        "algorithmic_bytes": bin_bytes_survey, "design_bytes": bin_bytes_design,
                        "frac_design_bytes": hbm(bin_bytes_design, t_bin) / hbm_peak, "n_dup": n_dup, "n_visible": n_vis,
                        "traffic": prof.get("binning_total", {}).get("dram_bytes"),
                        "note": "algorithmic = SURVEY 8(d) formula for a 6-pass 64-bit-key LSD sort (this design does not run it); "
                                "design = bytes the 9 binning kernels read + write; the stage is latency / issue bound (8 short "
                                "kernels), frac_design_bytes is the honest HBM fraction"}}
        del st

    # ---- end to end: host tensors in (pinned), host image out, through the C-level host pipeline ----
    # every rank uploads ITS copy of the scene, renders ITS view and reads ITS image back; the job's value is the
    # sum over ranks, timed as the max over ranks between two barriers
    host = [torch.from_numpy(a).pin_memory() for a in (cloud.xyz, cloud.scales, cloud.rots, cloud.opacity)]
    sh_host = torch.from_numpy(cloud.sh_planar()).pin_memory().permute(2, 1, 0)
    depth = 2
    img_hosts = [torch.empty((H, W, 3), dtype=torch.float32).pin_memory() for _ in range(depth)]
    cam_e2e = cams[rank % len(cams)]
    h2d = sum(t.numel() * 4 for t in host) + sh_host.numel() * 4
    d2h = img_hosts[0].numel() * 4
    for rend, _ in lanes[1:]:
        rend._ws = None
    torch.cuda.empty_cache()
    hr = R.HostRenderer(cloud.n, W, H, max_dup=renderer.max_dup, depth=depth)

    def wall(fn, n):
        fn(2)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        fn(n)
        t = torch.tensor([(time.perf_counter() - t0) / n], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def pipelined(n):   # `depth` frames in flight: upload of frame k+1 overlaps render + read-back of frame k
        tickets = []
        for i in range(n):
            if len(tickets) == depth:
                hr.wait(tickets.pop(0))
            tickets.append(hr.submit(*host, sh_host, cam_e2e, img_hosts[i % depth]))
        for t in tickets:
            hr.wait(t)

    def one_at_a_time(n):
        for _ in range(n):
            hr.render(*host, sh_host, cam_e2e, img_hosts[0])

    def resident(n):
        for _ in range(n):
            renderer.render(model, cam_e2e, out=frame, check=False)
            img_hosts[0].copy_(frame, non_blocking=True)
            torch.cuda.synchronize()

    copy_stream = torch.cuda.Stream()
    res_frames = [torch.empty_like(frame) for _ in range(2)]
    res_hosts = [torch.empty((H, W, 3), dtype=torch.float32).pin_memory() for _ in range(2)]

    def resident_pipelined(n):   # serving loop: the read-back of frame k (copy engine) runs under the render of frame k+1
        copied = [None, None]
        for i in range(n):
            b = i & 1
            if copied[b] is not None:
                torch.cuda.current_stream().wait_event(copied[b])   # the copy that read this buffer has finished
                copied[b].synchronize()                              # ... and the host owns image i-2 now
            renderer.render(model, cam_e2e, out=res_frames[b], check=False)
            done = torch.cuda.Event(); done.record()
            copy_stream.wait_event(done)
            with torch.cuda.stream(copy_stream):
                res_hosts[b].copy_(res_frames[b], non_blocking=True)
                copied[b] = torch.cuda.Event(); copied[b].record(copy_stream)
        torch.cuda.synchronize()

    lane_frames = [torch.empty_like(frame) for _ in lanes]
    lane_hosts = [torch.empty((H, W, 3), dtype=torch.float32).pin_memory() for _ in lanes]

    def resident_lanes(n):   # serving loop, frames in flight: lane i % L renders frame i on its stream and reads it back there
        done = [None] * len(lanes)
        for i in range(n):
            li = i % len(lanes)
            rend, strm = lanes[li]
            if done[li] is not None:
                done[li].synchronize()   # the host owns image i - L now; its buffers are free again
            with torch.cuda.stream(strm):
                rend.render(model, cam_e2e, out=lane_frames[li], check=False)
                lane_hosts[li].copy_(lane_frames[li], non_blocking=True)
                done[li] = torch.cuda.Event(); done[li].record(strm)
        torch.cuda.synchronize()

    # the PCIe floor of this box, measured: the same bytes as ONE monolithic pinned copy (no kernel, no read-back)
    floor_src = torch.empty(h2d // 4, dtype=torch.float32).pin_memory()
    floor_dst = torch.empty(h2d // 4, dtype=torch.float32, device="cuda")

    def h2d_floor(n):
        for _ in range(n):
            floor_dst.copy_(floor_src, non_blocking=True)
            torch.cuda.synchronize()

    dt_floor = wall(h2d_floor, 4)
    del floor_src, floor_dst
    dt = wall(pipelined, args.e2e_steps)
    dt_sync = wall(one_at_a_time, max(3, args.e2e_steps // 2))
    dt_res = wall(resident, 4 * args.e2e_steps)
    dt_res_pipe = wall(resident_pipelined, 8 * args.e2e_steps)
    hr.close()
    dt_res_lanes = wall(resident_lanes, 16 * args.e2e_steps)   # after the host renderer released its buffers: the lanes re-reserve
    lanes_status = max(r.frame_info()["status"] for r, _ in lanes)
    if rank == 0:
        line["e2e"] = {"value": world * W * H / 1e6 / dt, "unit": "MPix/s", "fps": world / dt, "ms_per_step": dt * 1e3,
                       "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
                       "what": "gsr_host_renderer (C ABI, gsr_b200.render.HostRenderer): pinned host tensors in, host image out, "
                               "every step uploads all Gaussian parameters (chunked, the raw fused preprocess kernel runs per "
                               "chunk as it lands - no prepared copy, so nothing is amortised), 2 frames in flight; PCIe-bound: "
                               "708 MB per frame is ~13 ms at ~54 GB/s, i.e. the 100 fps target is defined on the resident "
                               "scene; whole-job sum, max over ranks",
                       "h2d_floor_ms": dt_floor * 1e3, "h2d_floor_gbs": h2d / dt_floor / 1e9,
                       "h2d_floor_what": "one monolithic pinned copy of the same 708 MB with nothing else running: the PCIe "
                                         "floor of a host-buffer frame on this box",
                       "one_frame_at_a_time": {"fps": world / dt_sync, "ms_per_step": dt_sync * 1e3,
                                               "what": "same object, submit + wait per frame (latency of one host-buffer frame)"},
                       "resident_scene": {"fps": world / dt_res, "ms_per_step": dt_res * 1e3,
                                          "h2d_bytes_per_step": 176 * world, "d2h_bytes_per_step": d2h * world,
                                          "what": "scene kept in HBM between steps (video / serving use): camera up, image down, "
                                                  "synchronised every frame",
                                          "pipelined": {"fps": world / dt_res_pipe, "ms_per_step": dt_res_pipe * 1e3,
                                                        "what": "same, the image read-back of frame k on a copy stream under "
                                                                "the render of frame k+1 (two frame buffers)"},
                                          "frames_in_flight": {"fps": world / dt_res_lanes, "ms_per_step": dt_res_lanes * 1e3,
                                                               "lanes": len(lanes), "status": lanes_status,
                                                               "what": "same, one frame per lane in flight (the timed loop's "
                                                                       "renderers and streams), each lane reads its image back "
                                                                       "on its own stream; the host takes image i - lanes "
                                                                       "before it submits frame i"}}}
    del hr, host, sh_host
    if world > 1 and not args.no_banded:
        del model, lanes, renderer, frames, gathered
        graphs.clear()
        torch.cuda.empty_cache()
        try:
            c5 = banded_c5(world, rank, args.banded_n)
            if rank == 0:
                line["banded_c5"] = c5
        except Exception as exc:  # the sub-benchmark must never take the headline down with it
            if rank == 0:
                line["banded_c5"] = {"error": repr(exc)}
    if rank == 0:
        line["clocks"] = clocks.summary()
        if not args.no_cpu_baseline and world == 1:
            try:
                line["cpu_baseline"] = (cpu_reference_baseline(cloud.n, intr, budget_s=args.cpu_budget)
                                        or cpu_port_baseline(cloud, intr, degree, budget_s=args.cpu_budget))
            except Exception as exc:  # the baseline must never take the GPU number down with it
                line["cpu_baseline"] = {"error": repr(exc)}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
