"""Golden vectors for the drop-in helper functions, from the UNMODIFIED reference (run in the build container).

    python tests/golden/make_helper_golden.py   ->  tests/golden/helpers.npz
quaternion_to_rotation_matrix (/root/reference/rasterize.py:41-56) on 512 un-normalised quaternions (+ edge cases) and
get_covariance_matrix_from_mesh (:89-120) on a seeded 700-Gaussian PLY read through the reference's own call path.
"""
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(ROOT, "torch-gaussian-splatting-rasterizer_b200"), ROOT, HERE]

from gsr_b200 import scenes  # noqa: E402
from gsr_b200.ply import PlyData, write_gaussian_ply  # noqa: E402
from oracle import ref_harness  # noqa: E402


def main():
    ref = ref_harness.load_reference()["rasterize"]
    rng = np.random.default_rng(2024)
    q = rng.normal(0, 1, (4, 512)).astype(np.float32)
    q[:, :4] = np.eye(4, dtype=np.float32)          # axis-aligned unit quaternions
    q[:, 4] = 0.0                                   # the zero quaternion
    q[:, 5] = [1e-20, 0, 0, 0]                      # denormal-scale
    q[:, 6] = [3e4, -2e4, 1e4, 5e3]                 # large, far from unit norm
    R = ref.quaternion_to_rotation_matrix(torch.from_numpy(q)).numpy()
    cloud = scenes.synth_v1(700, 31, scale_factor=0.7)
    cloud.rots[:3] = 0.0                            # normalise() of a zero quaternion (eps path)
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "p.ply")
        write_gaussian_ply(path, cloud.ply_table())
        cov = ref.get_covariance_matrix_from_mesh(PlyData.read(path)).numpy()
    np.savez_compressed(os.path.join(HERE, "helpers.npz"), quat=q, rotmat=R, cov3d=cov)
    print("helpers.npz:", R.shape, cov.shape)


if __name__ == "__main__":
    main()
