"""Band-sharded render of one large frame (BASELINE config C5: 6 M Gaussians, 3840x2160) on the ranks of a torchrun job.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 profiles/prof_banded.py [C5]
Every rank preprocesses all Gaussians, bins and blends only its band of tile rows (balanced on the pairs per row),
and the bands are all-gathered; rank 0 prints the time per frame (max over ranks) next to the single-GPU time.
"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "torch-gaussian-splatting-rasterizer_b200"), ROOT]
import torch
import torch.distributed as dist
from gsr_b200 import camera, dist as gd, render as R, scenes

def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "C5"
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cloud, intr, views, degree = scenes.make_config(name)
    cam = camera.camera_for(intr, views[0], degree)
    model = R.GaussianModel.from_cloud(cloud)
    r = R.Renderer()
    full = r.render(model, cam)                      # single-GPU frame (also reserves the workspace)
    loads = gd.tile_row_loads(model, cam, r)         # identical on every rank
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    def timed(fn, reps=20):
        for _ in range(3):
            fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev[0].record()
        for _ in range(reps):
            fn()
        ev[1].record(); torch.cuda.synchronize()
        t = torch.tensor([ev[0].elapsed_time(ev[1]) / reps], device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    t_full = timed(lambda: r.render(model, cam, out=full, check=False))
    t_band = timed(lambda: gd.render_banded(model, cam, renderer=r, row_loads=loads))
    img = gd.render_banded(model, cam, renderer=r, row_loads=loads)
    same = bool(torch.equal(img, full))
    if rank == 0:
        bands = gd.partition_bands(loads, world)
        out = dict(config=name, n_gpus=world, single_gpu_ms=t_full, banded_ms=t_band, speedup=t_full / t_band,
                   bit_identical_to_single_gpu=same, bands=bands)
        print(json.dumps(out), flush=True)
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"banded_{name}_{world}gpu.json"), "w"), indent=1)
    if world > 1:
        dist.barrier(); dist.destroy_process_group()

if __name__ == "__main__":
    main()
