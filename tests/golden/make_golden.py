"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on CPU.

Run here (the build container), never on the GPU box:   python tests/golden/make_golden.py
The fixtures pin oracle/gsr_oracle.c and oracle/torch_port.py (tests/test_oracle_cpu.py) and are the
ground truth the CUDA path is compared with (tests/test_gpu_*.py).  Scenes are re-generated from
(seed, sizes) by gsr_b200.scenes; a checksum of the generated parameters is stored to detect drift.
"""
from __future__ import annotations

import logging
import os
import sys
import tempfile
import time

import numpy as np

os.environ.setdefault("TQDM_DISABLE", "1")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "torch-gaussian-splatting-rasterizer_b200"))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from gsr_b200 import scenes  # noqa: E402
from oracle import ref_harness  # noqa: E402

from cases import CASES, GRAD_CASE, build_case, cloud_digest  # noqa: E402


def main():
    logging.disable(logging.CRITICAL)
    if not ref_harness.reference_available():
        raise SystemExit("reference checkout not found; goldens can only be generated in the build container")
    only = [a for a in sys.argv[1:] if a in CASES]       # `make_golden.py needles`: (re)generate the named scenes only
    for name, c in CASES.items():
        if only and name not in only:
            continue
        cloud, intr, view = build_case(c)
        with tempfile.TemporaryDirectory() as d:
            scenes.write_scene_dir(d, cloud, intr, [view])
            t0 = time.time()
            out = ref_harness.run_reference(d, 0, stable=True)
            dt = time.time() - t0
            unstable = ref_harness.run_reference(d, 0, stable=False, capture=False)["image"]
        keep = {k: out[k] for k in ("image", "final_T", "cov3d", "cam_means", "mean2d", "cov2d", "conic", "tile_rect",
                                    "pixel_bbox", "rgb", "opacity", "depth", "order")}
        keep["tile_rect"] = keep["tile_rect"].astype(np.int32)
        keep["pixel_bbox"] = keep["pixel_bbox"].astype(np.int32)
        keep["order"] = keep["order"].astype(np.int32)
        keep["n_blended"] = np.int64(out["n_blended"])
        keep["digest"] = np.array(cloud_digest(cloud))
        keep["unstable_sort_maxabs"] = np.float64(np.abs(unstable - out["image"]).max())
        keep["ref_seconds"] = np.float64(dt)
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **keep)
        print(f"{name}: N={c['n']} {c['width']}x{c['height']} blended={out['n_blended']} ref={dt:.1f}s "
              f"unstable-vs-stable maxabs={float(keep['unstable_sort_maxabs']):.2e}")

    if only:
        return
    # gradients through the reference's own functions (small N)
    cloud, intr, view = build_case(GRAD_CASE)
    g = ref_harness.reference_autograd(cloud, intr, view, GRAD_CASE["degree"], grad_seed=3)
    g["digest"] = np.array(cloud_digest(cloud))
    np.savez_compressed(os.path.join(HERE, "grad_small.npz"), **g)
    print("grad_small: |d_xyz|max", np.abs(g["d_xyz"]).max(), "|d_sh|max", np.abs(g["d_sh"]).max())

    # BASELINE.json config 1 at full size: store per-16x16-tile sums and 4096 probe pixels (not 3 MB of image)
    cloud, intr, views, degree = scenes.make_config("C1")
    with tempfile.TemporaryDirectory() as d:
        scenes.write_scene_dir(d, cloud, intr, views)
        t0 = time.time()
        out = ref_harness.run_reference(d, 0, stable=True, capture=False)
        dt = time.time() - t0
    img = out["image"].astype(np.float64)
    H, W, _ = img.shape
    th, tw = (H + 15) // 16, (W + 15) // 16
    pad = np.zeros((th * 16, tw * 16, 3)); pad[:H, :W] = img
    tile_sums = pad.reshape(th, 16, tw, 16, 3).sum(axis=(1, 3))
    rng = np.random.default_rng(101)
    py, px = rng.integers(0, H, 4096), rng.integers(0, W, 4096)
    np.savez_compressed(os.path.join(HERE, "c1_probe.npz"), tile_sums=tile_sums, py=py, px=px,
                        probe=out["image"][py, px], digest=np.array(cloud_digest(cloud)),
                        ref_seconds=np.float64(dt), threads=np.int64(os.cpu_count() - 1))
    print(f"c1_probe: reference end-to-end {dt:.1f}s")


if __name__ == "__main__":
    main()
