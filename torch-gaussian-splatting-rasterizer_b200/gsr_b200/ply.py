"""Minimal PLY reader (binary little / big endian, ascii) and binary-little-endian writer for 3DGS point clouds.

The reference reads its Gaussians through the third-party ``plyfile`` package
(/root/reference/rasterize.py:353, /root/reference/utils.py:21-27), which is not
vendored and not installed here.  Only one thing about that package matters to
the render path: ``PlyData.read(path).elements[0][name]`` yields one numpy
column per vertex property.  This module provides exactly that for the
3DGS layout (62 little-endian float32 properties per vertex):

    x y z  nx ny nz  f_dc_0..2  f_rest_0..44  opacity  scale_0..2  rot_0..3

``PlyData`` / ``PlyElement`` mirror the two names the reference imports.
"""
from __future__ import annotations

import os
from typing import Dict, List, Sequence

import numpy as np

_NP_OF_PLY = {
    "float": "<f4", "float32": "<f4", "double": "<f8", "float64": "<f8",
    "uchar": "u1", "uint8": "u1", "char": "i1", "int8": "i1",
    "short": "<i2", "int16": "<i2", "ushort": "<u2", "uint16": "<u2",
    "int": "<i4", "int32": "<i4", "uint": "<u4", "uint32": "<u4",
}

GS_PROPERTIES: List[str] = (
    ["x", "y", "z", "nx", "ny", "nz"]
    + [f"f_dc_{i}" for i in range(3)]
    + [f"f_rest_{i}" for i in range(45)]
    + ["opacity"]
    + [f"scale_{i}" for i in range(3)]
    + [f"rot_{i}" for i in range(4)]
)
assert len(GS_PROPERTIES) == 62


class PlyElement:
    """One PLY element: column access by property name (``el["x"]`` -> ndarray)."""

    def __init__(self, name: str, data: np.ndarray):
        self.name = name
        self.data = data  # numpy structured array, one field per property

    def __getitem__(self, key: str) -> np.ndarray:
        return self.data[key]

    def __len__(self) -> int:
        return int(self.data.shape[0])

    @property
    def properties(self) -> Sequence[str]:
        return self.data.dtype.names


class PlyData:
    """``PlyData.read(path).elements[0]["opacity"]`` as the reference expects."""

    def __init__(self, elements: List[PlyElement]):
        self.elements = elements

    def __getitem__(self, name: str) -> PlyElement:
        for el in self.elements:
            if el.name == name:
                return el
        raise KeyError(name)

    @staticmethod
    def read(path: str) -> "PlyData":
        return PlyData(_read_elements(path))


def _read_elements(path: str) -> List[PlyElement]:
    with open(path, "rb") as fh:
        magic = fh.readline().strip()
        if magic != b"ply":
            raise ValueError(f"{path}: not a PLY file")
        fmt = None
        layout: List[Dict] = []
        while True:
            line = fh.readline()
            if not line:
                raise ValueError(f"{path}: unterminated PLY header")
            tok = line.decode("ascii").split()
            if not tok or tok[0] == "comment" or tok[0] == "obj_info":
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                layout.append({"name": tok[1], "count": int(tok[2]), "props": []})
            elif tok[0] == "property":
                if tok[1] == "list":
                    raise ValueError(f"{path}: list properties are not supported")
                layout[-1]["props"].append((tok[2], _NP_OF_PLY[tok[1]]))
            elif tok[0] == "end_header":
                break
        if fmt not in ("binary_little_endian", "binary_big_endian", "ascii"):
            raise ValueError(f"{path}: unknown PLY format {fmt}")
        out = []
        for el in layout:
            native = np.dtype(el["props"])  # little-endian record: what every consumer (and the GPU unpack) expects
            if fmt == "ascii":
                # one vertex per line, properties separated by blanks (what plyfile writes with text=True)
                rows = []
                while len(rows) < el["count"]:
                    line = fh.readline()
                    if not line:
                        raise ValueError(f"{path}: truncated element {el['name']}")
                    if line.strip():
                        rows.append(line.split())
                arr = np.empty(el["count"], dtype=native)
                if el["count"]:
                    cols = list(zip(*rows))
                    if len(cols) != len(native.names):
                        raise ValueError(f"{path}: element {el['name']} has {len(cols)} columns, header says {len(native.names)}")
                    for name, col in zip(native.names, cols):
                        kind = native[name].kind
                        arr[name] = np.array(col, dtype=np.float64 if kind == "f" else np.int64).astype(native[name])
                out.append(PlyElement(el["name"], arr))
                continue
            file_dt = native if fmt == "binary_little_endian" else native.newbyteorder(">")
            buf = fh.read(file_dt.itemsize * el["count"])
            if len(buf) != file_dt.itemsize * el["count"]:
                raise ValueError(f"{path}: truncated element {el['name']}")
            arr = np.frombuffer(buf, dtype=file_dt, count=el["count"])
            out.append(PlyElement(el["name"], arr if file_dt == native else arr.astype(native)))
        return out


def write_gaussian_ply(path: str, table: np.ndarray) -> None:
    """Write an ``[N, 62]`` float32 table in the 3DGS vertex layout."""
    table = np.ascontiguousarray(table, dtype="<f4")
    if table.ndim != 2 or table.shape[1] != len(GS_PROPERTIES):
        raise ValueError("expected an [N, 62] table")
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    head = ["ply", "format binary_little_endian 1.0", f"element vertex {table.shape[0]}"]
    head += [f"property float {p}" for p in GS_PROPERTIES]
    head += ["end_header", ""]
    with open(path, "wb") as fh:
        fh.write("\n".join(head).encode("ascii"))
        fh.write(table.tobytes())


def read_gaussian_table(path: str) -> np.ndarray:
    """Read the vertex element as a dense ``[N, P]`` float32 table plus check the 3DGS columns exist."""
    el = PlyData.read(path).elements[0]
    missing = [p for p in GS_PROPERTIES if p not in el.properties and p not in ("nx", "ny", "nz")]
    if missing:
        raise ValueError(f"{path}: missing 3DGS properties {missing[:4]}...")
    return el


def read_vertex_block(path: str):
    """The vertex element as ONE ``[n, P]`` float32 array (a view of the file's records) plus the property names.

    Requires every vertex property to be a 4-byte float (true for 3DGS checkpoints); used by the device loader,
    which de-interleaves the block on the GPU instead of touching 59 columns on the host."""
    el = PlyData.read(path).elements[0]
    dt = el.data.dtype
    if any(dt[name] != np.dtype("<f4") for name in dt.names):
        raise ValueError(f"{path}: vertex properties are not all float32")
    block = np.ascontiguousarray(el.data).view("<f4").reshape(len(el), len(dt.names))  # a view: no per-column work
    return block, list(dt.names)
