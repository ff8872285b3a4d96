#!/usr/bin/env python
"""Benchmark of the render path (BASELINE.json: MPix/s & frames/s, 3M Gaussians SH3 @1080p).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one frame per GPU.  N=1: config C3 (3M synth_v1 Gaussians, SH degree 3, 1920x1080, front view),
the Gaussians resident in HBM.  N>1: the same scene replicated on every GPU, rank r renders orbit view
(step*N + r) of the 256-view orbit (config C4: view sharding), frames gathered with one NCCL all_gather per
step on a side stream - weak scaling (one frame per GPU per step).

Prints ONE JSON line (rank 0).  Keys beyond the driver's contract:
  fps, stages_ms, roofline (blend = dominant kernel, FP32-FMA bound), roofline_stages (preprocess / binning HBM),
  cpu_baseline (oracle torch port on the host cores, bounded sample), e2e (host tensors in, host image out), clocks.
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

import numpy as np  # noqa: E402
import torch  # noqa: E402

FLOP_PER_EVAL = 24.0           # one Gaussian-pixel evaluation
PREPROCESS_BYTES = {0: 104.0, 1: 140.0, 2: 200.0, 3: 284.0}  # (11 + 3(d+1)^2)*4 in + 48 out, per Gaussian


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            d = json.load(fh)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


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


def build_scene(n_override=None):
    from gsr_b200 import scenes

    c = scenes.CONFIGS["C3"]
    n = c["n"] if n_override is None else n_override
    t0 = time.time()
    cloud = scenes.synth_v1(n, c["seed"])
    intr = scenes.Intrinsics.synthetic(c["width"], c["height"])
    log(f"[bench] generated {n} Gaussians in {time.time() - t0:.1f}s")
    return cloud, intr, c["degree"]


# ----------------------------------------------------------------------------------------------
# CPU baseline: the oracle's torch port (same execution model as the reference) on a bounded sample
# ----------------------------------------------------------------------------------------------
def cpu_baseline(cloud, intr, degree, budget_s=20.0, quiet=False):
    from gsr_b200 import camera, scenes
    from oracle import torch_port

    cam = camera.camera_for(intr, scenes.front_view(0), degree)
    cores = os.cpu_count() or 1
    tens = torch_port.cloud_tensors(cloud)
    # the per-Gaussian loop is serial; pick the thread count at which eager torch ops are fastest on this host
    best = None
    for nt in sorted({1, min(4, cores), max(1, cores - 1)}):
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


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cloud, intr, degree = build_scene(args.n)
    vals = []
    for _ in range(max(1, min(args.steps, 2))):
        vals.append(cpu_baseline(cloud, intr, degree, budget_s=args.cpu_budget))
    b = vals[-1]
    # same metric / unit / workload strings as our arm: the driver divides one line by the other
    line = {"impl": "reference", "metric": "MPix/s, 3M-Gaussian SH3 @1080p render",
            "what": "the reference's CPU render timed through the oracle's torch port (the reference is Python and cannot travel to the GPU box)",
            "value": b["value"], "unit": "MPix/s", "fps": b["value"] * 1e6 / (intr.width * intr.height),
            "n_gpus": args.gpus, "steps": len(vals), "warmup": 0, "ms_per_step": b["est_seconds_per_frame"] * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": ("C3: 3M synth_v1 Gaussians (seed 3), SH degree 3, 1920x1080, front view"
                                    if args.gpus <= 1 else
                                    "C4: same 3M-Gaussian scene, 256-view orbit (the CPU render of ONE view is timed, on "
                                    "rank 0; views are independent and cost the same)") if args.n is None
                       else f"C3 scene with n={args.n}"},
            "cpu_baseline": b, "e2e": {"value": b["value"], "unit": "MPix/s", "h2d_bytes_per_step": 0,
                                       "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


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
    ap.add_argument("--gather", choices=["p2p", "nccl"], default="p2p", help="N>1: how finished frames are all-gathered")
    ap.add_argument("--ring", type=int, default=8, help="frame buffers per rank (N>1): how far a rank may run ahead of the gather")
    ap.add_argument("--cpu-budget", type=float, default=20.0)
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
    renderer = R.Renderer()
    n_streams = max(1, args.streams)
    lanes = [(renderer, torch.cuda.current_stream())] + [(R.Renderer(), torch.cuda.Stream()) for _ in range(n_streams - 1)]
    if world == 1 and not args.orbit:
        views = [scenes.front_view(0)]
    else:
        views = scenes.orbit_views(256)
    cams = [camera.camera_for(intr, v, degree) for v in views]

    def cam_for(step):
        return cams[(step * world + rank) % len(cams)]

    frame = torch.empty((H, W, 3), dtype=torch.float32, device="cuda")
    gathered = torch.empty((world, H, W, 3), dtype=torch.float32, device="cuda") if world > 1 else None
    comm = torch.cuda.Stream() if world > 1 else None
    peer = None
    if world > 1 and args.gather == "p2p":
        from gsr_b200.dist import PeerGather
        peer = PeerGather.create((H, W, 3), "cuda", slots=4)
    gather_kind = "p2p-dma" if peer is not None else ("nccl" if world > 1 else None)
    n_buf = max(args.ring, n_streams)  # frames a rank may run ahead of the slowest rank's gather
    frames = [torch.empty_like(frame) for _ in range(n_buf)]
    gathered_ev = [None] * n_buf

    def step(i):
        # Successive frames are independent: frame i runs on lane i % n_streams (own renderer, workspace and stream),
        # so the latency-bound binning kernels of one frame overlap the issue-bound blend of another.
        rend, strm = lanes[i % n_streams]
        buf = frames[i % n_buf]
        with torch.cuda.stream(strm):
            if gathered_ev[i % n_buf] is not None:  # the gather that last read this buffer must have finished
                strm.wait_event(gathered_ev[i % n_buf])
            rend.render(model, cam_for(i), out=buf, check=False)
            if world > 1:  # the only collective: gather the finished frames (overlaps with the next render)
                done = torch.cuda.Event(); done.record(strm)
                comm.wait_event(done)
                with torch.cuda.stream(comm):
                    if peer is not None:  # copy-engine gather over NVLink (gsr_b200.dist.PeerGather): no SM time
                        peer.push(buf, i % peer.slots)
                    else:
                        dist.all_gather_into_tensor(gathered.view(-1), buf.view(-1))
                    gathered_ev[i % n_buf] = torch.cuda.Event(); gathered_ev[i % n_buf].record(comm)

    def fork():   # side lanes start after everything already queued on the main stream
        ev = torch.cuda.Event(); ev.record()
        for _, strm in lanes[1:]:
            strm.wait_event(ev)

    def join():   # the main stream waits for every lane (and the gathers)
        for _, strm in lanes[1:]:
            torch.cuda.current_stream().wait_stream(strm)
        if world > 1:
            torch.cuda.current_stream().wait_stream(comm)

    # reserve workspaces (one checked frame per lane also validates max_dup)
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

    fork()
    for i in range(args.warmup):
        step(i)
    join()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        barrier()
        ev0.record()
        fork()
        for i in range(args.steps):
            step(args.warmup + i)
        join()
        ev1.record()
        barrier()
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    frames_per_s = world * args.steps / (ms / 1e3)
    mpix_s = frames_per_s * W * H / 1e6
    status_after = max(r.frame_info()["status"] for r, _ in lanes)

    line = {"metric": "MPix/s, 3M-Gaussian SH3 @1080p render", "value": mpix_s, "unit": "MPix/s", "fps": frames_per_s,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": ("C3: 3M synth_v1 Gaussians (seed 3), SH degree 3, 1920x1080, front view" if world == 1
                                    else f"C4: same 3M-Gaussian scene replicated per GPU, 256-view orbit sharded by view "
                                         f"(rank r renders view step*{world}+r), frames all-gathered ({gather_kind})"),
                       "n_gaussians": cloud.n, "width": W, "height": H, "sh_degree": degree,
                       "l2_policy": "inputs (708 MB of Gaussian parameters) exceed the 126 MB L2; no flush needed",
                       "parallelism": f"view-sharded x{world}" if world > 1 else "single GPU",
                       "frames_in_flight": n_streams},
            "gpu_launches": 10 * args.steps, "frame_status": status_after}

    if rank == 0:
        # ---- per-stage timing with CUDA events on the launching stream (same kernels, stage entry points) ----
        cam = cams[0]
        pre = R.preprocess(model, cam, prepared=True)
        binned = R.bin_tiles(pre, W, H, renderer.max_dup)
        out = R.blend(pre, binned, W, H, count=True, track=False)
        torch.cuda.synchronize()
        cnt = out["info"].cpu().numpy().view(np.uint32)
        warp_evals = int(cnt[4]) | (int(cnt[5]) << 32)
        pixel_evals = int(cnt[6]) | (int(cnt[7]) << 32)
        n_dup = info0["n_dup"]
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        acc = np.zeros(3)
        reps = max(5, min(args.steps, 20))
        for _ in range(3):
            pre = R.preprocess(model, cam, prepared=True); binned = R.bin_tiles(pre, W, H, renderer.max_dup); R.blend(pre, binned, W, H, track=False)
        torch.cuda.synchronize()
        for _ in range(reps):
            evs[0].record(); pre = R.preprocess(model, cam, prepared=True)
            evs[1].record(); binned = R.bin_tiles(pre, W, H, renderer.max_dup)
            evs[2].record(); R.blend(pre, binned, W, H, track=False)
            evs[3].record(); torch.cuda.synchronize()
            acc += [evs[k].elapsed_time(evs[k + 1]) for k in range(3)]
        t_pre, t_bin, t_blend = (acc / reps).tolist()
        peak_t, mhz = _cabi.ctypes.c_double(), _cabi.ctypes.c_double()
        _cabi.call("gsr_fp32_fma_peak", _cabi.ctypes.byref(peak_t), _cabi.ctypes.byref(mhz))
        fp32_peak = peak_t.value
        tiles = ((W + 15) // 16) * ((H + 15) // 16)
        passes = -(-(32 + int(np.ceil(np.log2(tiles)))) // 8)
        bin_bytes = 8 * cloud.n + n_dup * (28 + 24 * passes) + 8 * tiles
        pre_bytes = PREPROCESS_BYTES[degree] * cloud.n
        blend_tflops = warp_evals * FLOP_PER_EVAL / (t_blend * 1e-3) / 1e12
        line["stages_ms"] = {"preprocess": t_pre, "binning": t_bin, "blend": t_blend,
                             "note": "stage entry points timed back to back with CUDA events; includes workspace allocs in binning"}
        traffic = None  # DRAM bytes per launch of the blend kernel from the committed ncu --set full capture
        tpath = os.path.join(ROOT, "profiles", "r01_blend_traffic.json")
        if os.path.exists(tpath) and args.n is None:
            with open(tpath) as fh:
                tj = json.load(fh)
            traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
        line["roofline"] = {"kernel": "blend_kernel", "bound": "fp32_fma", "achieved": blend_tflops, "peak": fp32_peak,
                            "unit": "TFLOP/s", "frac": blend_tflops / fp32_peak if fp32_peak else None, "traffic": traffic,
                            "traffic_note": "bytes; the kernel is issue-bound - algorithmic memory side is n_dup*40 + 12*W*H",
                            "peak_source": f"FMA micro-benchmark in this run (gsr_fp32_fma_peak, implies {mhz.value:.0f} MHz)",
                            "evals_warp_granular": warp_evals, "evals_live_pixels": pixel_evals,
                            "gevals_per_s": warp_evals / (t_blend * 1e-3) / 1e9, "flop_per_eval": FLOP_PER_EVAL}
        line["roofline_stages"] = {
            "preprocess": {"bound": "hbm", "achieved": pre_bytes / (t_pre * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                           "frac": pre_bytes / (t_pre * 1e-3) / 1e9 / hbm_peak, "peak_source": peak_src,
                           "algorithmic_bytes": pre_bytes},
            "binning": {"bound": "hbm", "achieved": bin_bytes / (t_bin * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": bin_bytes / (t_bin * 1e-3) / 1e9 / hbm_peak, "peak_source": peak_src,
                        "algorithmic_bytes": bin_bytes, "n_dup": n_dup,
                        "note": "algorithmic bytes = SURVEY 8(d) formula for a 64-bit-key LSD sort; this design moves fewer"}}
        del pre, binned, out

    # ---- end to end: host tensors in (pinned), host image out, through the public render() ----
    # every rank uploads ITS copy of the scene, renders ITS view and reads ITS image back; the job's value is the
    # sum over ranks, timed as the max over ranks between two barriers
    host = [torch.from_numpy(a).pin_memory() for a in (cloud.xyz, cloud.scales, cloud.rots, cloud.opacity)]
    sh_host = torch.from_numpy(cloud.sh_planar()).pin_memory().permute(2, 1, 0)
    img_host = torch.empty((H, W, 3), dtype=torch.float32).pin_memory()
    v = views[rank % len(views)]
    h2d = sum(t.numel() * 4 for t in host) + sh_host.numel() * 4
    d2h = img_host.numel() * 4

    def e2e_step():
        img = R.render(host[0], host[1], host[2], host[3], sh_host, v.qvec, v.tvec, intr.fx, intr.fy,
                       intr.cam_width, intr.cam_height, W, H, degree, renderer=renderer)
        img_host.copy_(img, non_blocking=True)
        torch.cuda.synchronize()

    def resident_step():
        renderer.render(model, cams[rank % len(cams)], out=frame, check=False)
        img_host.copy_(frame, non_blocking=True)
        torch.cuda.synchronize()

    def timed(step, n):
        for _ in range(2):
            step()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(n):
            step()
        t = torch.tensor([(time.perf_counter() - t0) / n], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    dt = timed(e2e_step, args.e2e_steps)
    # for context: the same loop with the scene resident (only the 176-byte camera goes up, the image comes down)
    dt_res = timed(resident_step, 4 * args.e2e_steps)
    if rank == 0:
        line["e2e"] = {"value": world * W * H / 1e6 / dt, "unit": "MPix/s", "fps": world / dt, "ms_per_step": dt * 1e3,
                       "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
                       "what": "rasterize.render(host tensors) -> host image on every rank: uploads all Gaussian "
                               "parameters every step (PCIe-bound); whole-job sum, max over ranks",
                       "resident_scene": {"fps": world / dt_res, "ms_per_step": dt_res * 1e3,
                                          "h2d_bytes_per_step": 176 * world, "d2h_bytes_per_step": d2h * world,
                                          "what": "scene kept in HBM between steps (video / serving use): camera up, image down"}}
    if rank == 0:
        line["clocks"] = clocks.summary()
        if not args.no_cpu_baseline and world == 1:
            try:
                line["cpu_baseline"] = cpu_baseline(cloud, intr, degree, budget_s=args.cpu_budget)
            except Exception as exc:  # the baseline must never take the GPU number down with it
                line["cpu_baseline"] = {"error": repr(exc)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
