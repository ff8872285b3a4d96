"""CPU-side checks of the C-ABI boundary: the library builds/loads and exports every symbol the header declares;
argument validation that needs no GPU; host-side camera math against the reference's captured values."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from cases import CASES, build_case
from gsr_b200 import _cabi, build, camera

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "gsr_b200.h")


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _cabi.load()


def _declared_functions():
    text = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    return sorted(set(re.findall(r"\b(gsr_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    names = _declared_functions()
    assert len(names) >= 18
    for name in names:
        assert hasattr(lib, name), f"{name} declared in gsr_b200.h but not exported"
    bound = set(_cabi.PROTOTYPES) | {"gsr_last_error"}
    assert set(names) == bound, set(names) ^ bound


def test_abi_version_and_struct_sizes(lib):
    assert lib.gsr_abi_version() == 2
    assert ctypes.sizeof(camera.CameraStruct) == 4 * (16 + 16 + 3 + 2 + 2) + 16
    assert ctypes.sizeof(_cabi.Gaussians) == 8 * 16
    assert ctypes.sizeof(_cabi.FrameInfo) == 32


def test_invalid_arguments_are_reported_not_crashed(lib):
    need = ctypes.c_size_t()
    assert lib.gsr_render_workspace_bytes(-1, 64, 64, 100, ctypes.byref(need)) == _cabi.GSR_ERR_INVALID_ARGUMENT
    assert b"invalid argument" in lib.gsr_last_error()
    assert lib.gsr_render_workspace_bytes(1000, 64, 64, 8000, ctypes.byref(need)) == _cabi.GSR_OK
    assert need.value > 1000 * 48
    with pytest.raises(_cabi.GsrError):
        _cabi.call("gsr_sort_pairs", None, None, None, None, None, None, 10, None, 0, 32, None, 0, None)


def test_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import rasterize
    import spherical_harmonics

    with pytest.raises(_cabi.GsrError):
        spherical_harmonics.sh_to_rgb(torch.zeros(4, 3), torch.zeros(4, 16, 3), torch.eye(4), degree=3)
    with pytest.raises(_cabi.GsrError):
        rasterize.render(torch.zeros(4, 3), torch.zeros(4, 3), torch.ones(4, 4), torch.zeros(4), torch.zeros(4, 16, 3),
                         [1, 0, 0, 0], [0, 0, 4], 100.0, 100.0, 64, 64, 32, 32)


@pytest.mark.parametrize("name", ["synth_small", "orbit_b", "tilted"])
def test_host_camera_reproduces_reference_projection(name, golden):
    """cam_means / mean2d captured from the reference are exact functions of the camera struct: recompute them in
    numpy with the documented fma recipe and require bit equality."""
    c = CASES[name]
    cloud, intr, view = build_case(c)
    g = golden(name)
    cam = camera.camera_for(intr, view, 3)
    m = np.array(list(cam.w2c_t), dtype=np.float32).reshape(4, 4)
    x, y, z = (cloud.xyz[:, i].astype(np.float64) for i in range(3))
    # fma in float64 of f32 inputs is exact up to the final rounding: products of two f32 fit in f64
    def f32(a):
        return a.astype(np.float32)
    acc = f32(x * m[0, 2].astype(np.float64))
    acc = f32(y * np.float64(m[1, 2]) + acc.astype(np.float64))
    acc = f32(z * np.float64(m[2, 2]) + acc.astype(np.float64))
    depth = acc + m[3, 2]
    assert np.array_equal(depth.view(np.int32), g["depth"].view(np.int32))


def test_drop_in_modules_expose_reference_names():
    import data_reader
    import rasterize
    import spherical_harmonics
    import utils

    for name in ("quaternion_to_rotation_matrix", "get_world_to_camera_matrix", "project_to_camera_space",
                 "get_covariance_matrix_from_mesh", "get_projection_matrix", "compute_covering_bbox",
                 "compute_2d_covariance", "rasterize_gaussian", "run_rasterization", "render", "Z_FAR", "Z_NEAR",
                 "GAUSSIAN_SPREAD", "BLOCK_SIZE", "MAX_GAUSSIAN_DENSITY", "MIN_ALPHA"):
        assert hasattr(rasterize, name), name
    assert (rasterize.Z_FAR, rasterize.Z_NEAR, rasterize.GAUSSIAN_SPREAD, rasterize.BLOCK_SIZE) == (100.0, 0.01, 3, 16)
    assert rasterize.MAX_GAUSSIAN_DENSITY == 0.99 and rasterize.MIN_ALPHA == 1 / 255
    for name in ("sh_to_rgb", "SH_0", "SH_C1", "SH_C2", "SH_C3"):
        assert hasattr(spherical_harmonics, name)
    assert len(spherical_harmonics.SH_C2) == 5 and len(spherical_harmonics.SH_C3) == 7
    for name in ("read_color_components", "read_scene"):
        assert hasattr(utils, name)
    for name in ("Camera", "BaseImage", "read_extrinsics_binary", "read_intrinsics_binary", "CAMERA_MODEL_IDS"):
        assert hasattr(data_reader, name)
    opts = {p.name for p in rasterize.run_rasterization.params}
    assert opts == {"input_dir", "trained_model_path", "output_path", "scene_index", "scale_factor", "generate_video"}


def test_scene_io_round_trip(tmp_path):
    import utils
    from gsr_b200 import scenes
    from gsr_b200.ply import PlyData

    cloud = scenes.synth_v1(257, 5)
    intr = scenes.Intrinsics.synthetic(70, 50)
    views = scenes.orbit_views(3, total=8)
    root = scenes.write_scene_dir(str(tmp_path), cloud, intr, views)
    imgs, cams = utils.read_scene(root)
    assert sorted(imgs) == [0, 1, 2] and list(cams) == [1]
    assert cams[1].model == "PINHOLE" and (cams[1].width, cams[1].height) == (140, 100)
    assert np.allclose(cams[1].params, [168.0, 168.0, 70.0, 50.0])
    for v in views:
        assert np.array_equal(imgs[v.image_id].qvec, v.qvec) and np.array_equal(imgs[v.image_id].tvec, v.tvec)
        assert imgs[v.image_id].name == "view.png" and imgs[v.image_id].camera_id == 1
    ply = PlyData.read(os.path.join(root, "point_cloud/iteration_30000/point_cloud.ply"))
    assert np.array_equal(ply.elements[0]["opacity"], cloud.opacity)
    sh = utils.read_color_components(ply)
    assert tuple(sh.shape) == (257, 16, 3) and sh.stride() == (1, 257, 16 * 257)
    assert np.array_equal(sh.numpy(), cloud.sh_planar().transpose(2, 1, 0))


def test_ply_reader_takes_every_format_plyfile_does(tmp_path):
    """The reference reads its checkpoint through `plyfile` (rasterize.py:353), which accepts binary little / big endian
    and ascii files.  The in-tree reader yields the same little-endian float32 columns for all three, plus mixed
    property types, comments and a second element after the vertices; malformed files raise ValueError."""
    import utils
    from gsr_b200 import ply, scenes

    cloud = scenes.synth_v1(37, 5)
    table = cloud.ply_table()
    base = str(tmp_path / "le.ply")
    ply.write_gaussian_ply(base, table)
    head = open(base, "rb").read().split(b"end_header\n")[0] + b"end_header\n"
    head = head.replace(b"element vertex", b"comment written by a test\nelement vertex")
    be, txt = str(tmp_path / "be.ply"), str(tmp_path / "ascii.ply")
    open(be, "wb").write(head.replace(b"binary_little_endian", b"binary_big_endian") + table.astype(">f4").tobytes())
    body = "\n".join(" ".join(repr(float(v)) for v in row) for row in table) + "\n"
    open(txt, "wb").write(head.replace(b"binary_little_endian", b"ascii") + body.encode())
    for path in (base, be, txt):
        block, names = ply.read_vertex_block(path)
        assert names == ply.GS_PROPERTIES and block.dtype == np.float32 and np.array_equal(block, table), path
        data = ply.PlyData.read(path)
        assert np.array_equal(data.elements[0]["opacity"], cloud.opacity)
        assert np.array_equal(utils.read_color_components(data).numpy(), cloud.sh_planar().transpose(2, 1, 0))
    # mixed property types and a trailing element (a mesh export): columns by name, faces ignored by the render path
    mixed = str(tmp_path / "mixed.ply")
    hdr = ("ply\nformat binary_little_endian 1.0\nelement vertex 3\nproperty double x\nproperty float y\n"
           "property uchar red\nelement other 2\nproperty int a\nend_header\n").encode()
    rec = np.zeros(3, dtype=[("x", "<f8"), ("y", "<f4"), ("red", "u1")])
    rec["x"], rec["y"], rec["red"] = [1.5, -2.0, 3.25], [0.5, 0.25, 0.125], [7, 8, 9]
    open(mixed, "wb").write(hdr + rec.tobytes() + np.array([4, 5], "<i4").tobytes())
    data = ply.PlyData.read(mixed)
    assert np.array_equal(data["vertex"]["x"], rec["x"]) and np.array_equal(data["vertex"]["red"], rec["red"])
    assert np.array_equal(data["other"]["a"], [4, 5]) and len(data.elements[0]) == 3
    with pytest.raises(ValueError):
        ply.read_vertex_block(mixed)              # the device loader needs all-float32 vertices
    for bad in (b"plx\n", head.replace(b"binary_little_endian", b"binary_middle_endian"), head + table.tobytes()[:-5],
                head.replace(b"binary_little_endian", b"ascii") + body.encode()[: len(body) // 2].rsplit(b"\n", 1)[0] + b"\n"):
        broken = str(tmp_path / "broken.ply")
        open(broken, "wb").write(bad)
        with pytest.raises(ValueError):
            ply.PlyData.read(broken)


def test_ground_truth_comparison_and_frame_writer_run_on_the_host(tmp_path):
    """The helpers around the path are plain host code: PSNR / side-by-side PNG against a photo and
    the frame-sequence writer need neither the library nor a GPU."""
    import numpy as np
    import torch
    from PIL import Image

    import rasterize

    rng = np.random.default_rng(0)
    a = torch.from_numpy(rng.random((20, 30, 3), dtype=np.float32))
    b = (a + 0.01).clamp(0, 1)
    assert rasterize.psnr(a, a) == float("inf")
    assert 35.0 < rasterize.psnr(a, b) < 45.0  # an offset of 0.01 is 40 dB
    gt = tmp_path / "gt.png"
    Image.fromarray((b.numpy() * 255).astype(np.uint8)).save(gt)
    out = rasterize.compare_with_ground_truth(a, str(gt), str(tmp_path / "cmp.png"))
    assert out["psnr_db"] > 30 and out["mae"] < 0.02
    assert Image.open(tmp_path / "cmp.png").size == (30, 40)  # render above the photo
    n = rasterize.write_frame_sequence(torch.stack([a, b]), str(tmp_path / "seq"))
    assert n == 2 and sorted(p.name for p in (tmp_path / "seq").iterdir()) == ["frame_0000000.png", "frame_0000001.png"]


def test_host_struct_layouts():
    """gsr_b200.render._host_struct: what the host pipeline is handed for each layout of the host tensors - dense rows
    and planar / dense SH are passed as they are (no copy: the data pointer is the caller's), anything else is made
    dense first."""
    import torch

    from gsr_b200 import render as R, scenes

    cloud = scenes.synth_v1(37, 3)
    xyz, ls, rot, op = (torch.from_numpy(a) for a in (cloud.xyz, cloud.scales, cloud.rots, cloud.opacity))
    planar = torch.from_numpy(cloud.sh_planar())            # [3,16,N] storage
    sh_view = planar.permute(2, 1, 0)                        # the reference's [N,16,3] view (utils.py:31)
    g, keep = R._host_struct(xyz, ls, rot, op, sh_view)
    assert g.n == 37 and (g.sh_sn, g.sh_sk, g.sh_sc) == (1, 37, 16 * 37) and g.sh == planar.data_ptr()
    assert g.xyz == xyz.data_ptr() and (g.xyz_sn, g.xyz_sc) == (3, 1) and g.opacity_sn == 1
    dense = sh_view.contiguous()
    g, keep = R._host_struct(xyz, ls, rot, op, dense)
    assert (g.sh_sn, g.sh_sk, g.sh_sc) == (48, 3, 1) and g.sh == dense.data_ptr()
    odd = torch.zeros((37, 16, 6))[:, :, ::2]                # a strided view that is neither: copied to dense rows
    g, keep = R._host_struct(xyz.double(), ls, rot[:, :].t().contiguous().t(), op, odd)
    assert (g.sh_sn, g.sh_sk, g.sh_sc) == (48, 3, 1) and g.sh != odd.data_ptr()
    assert g.xyz != xyz.data_ptr() and (g.rot_sn, g.rot_sc) == (4, 1)   # float64 / column-major inputs were converted
