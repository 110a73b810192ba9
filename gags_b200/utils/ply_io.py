"""plyfile-free PLY reader / writer for the Gaussian point cloud, and the per-view feature files.

The reference stores a trained scene as `point_cloud.ply` through the `plyfile` package
(/root/reference/scene/gaussian_model.py:222-259 save, :266-319 load): ONE `vertex` element whose
properties are all float32 — x y z nx ny nz f_dc_* f_rest_* opacity scale_* rot_* semantic_* — in
binary little-endian.  `plyfile` is not a dependency here: the header is written / parsed by hand and
the body is one numpy structured array, byte-compatible with what plyfile produces and reads.

Per-view distillation inputs (/root/reference/preprocess.py:332-336, read back at
/root/reference/scene/dataset_readers.py:183-187): `<image>_f.npy` = float embeddings
[n_segments, C], `<image>_s.npy` = segment maps [4, h, w].
"""
from __future__ import annotations

import os
from typing import Dict, List, Tuple

import numpy as np

_PLY_TYPES = {"char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1", "short": "i2",
              "int16": "i2", "ushort": "u2", "uint16": "u2", "int": "i4", "int32": "i4",
              "uint": "u4", "uint32": "u4", "float": "f4", "float32": "f4", "double": "f8",
              "float64": "f8"}


def write_vertex_ply(path: str, names: List[str], columns: np.ndarray) -> None:
    """`columns` [N, len(names)] -> binary little-endian PLY with one float32 property per name."""
    columns = np.ascontiguousarray(columns, dtype="<f4")
    if columns.ndim != 2 or columns.shape[1] != len(names):
        raise ValueError("columns must be [N, len(names)]")
    d = os.path.dirname(path)
    if d:
        os.makedirs(d, exist_ok=True)
    header = ["ply", "format binary_little_endian 1.0", f"element vertex {columns.shape[0]}"]
    header += [f"property float {n}" for n in names]
    header.append("end_header")
    with open(path, "wb") as f:
        f.write(("\n".join(header) + "\n").encode("ascii"))
        f.write(columns.tobytes())


def read_vertex_ply(path: str) -> Tuple[List[str], Dict[str, np.ndarray]]:
    """Reads the `vertex` element of an ascii or binary PLY (scalar properties only, as 3DGS / GAGS
    write them).  Returns (property names in file order, name -> 1-D array)."""
    with open(path, "rb") as f:
        if f.readline().strip() != b"ply":
            raise ValueError(f"{path}: not a PLY file")
        fmt, elements, cur = None, [], None
        while True:
            line = f.readline()
            if not line:
                raise ValueError(f"{path}: header without end_header")
            tok = line.decode("ascii").split()
            if not tok or tok[0] == "comment" or tok[0] == "obj_info":
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                cur = {"name": tok[1], "count": int(tok[2]), "props": []}
                elements.append(cur)
            elif tok[0] == "property":
                if tok[1] == "list":
                    raise ValueError(f"{path}: list properties are not supported")
                cur["props"].append((tok[2], _PLY_TYPES[tok[1]]))
            elif tok[0] == "end_header":
                break
        if fmt not in ("ascii", "binary_little_endian", "binary_big_endian"):
            raise ValueError(f"{path}: unknown PLY format {fmt}")
        out = None
        for el in elements:
            endian = ">" if fmt == "binary_big_endian" else "<"
            dt = np.dtype([(n, endian + t) for n, t in el["props"]])
            if fmt == "ascii":
                rows = [f.readline().split() for _ in range(el["count"])]
                arr = np.array([tuple(float(v) for v in r) for r in rows], dtype=dt) \
                    if rows else np.empty(0, dtype=dt)
            else:
                arr = np.frombuffer(f.read(dt.itemsize * el["count"]), dtype=dt, count=el["count"])
            if el["name"] == "vertex":
                out = ([n for n, _ in el["props"]], {n: np.asarray(arr[n]) for n, _ in el["props"]})
                break
        if out is None:
            raise ValueError(f"{path}: no vertex element")
        return out


def save_feature_files(prefix: str, img_embed: np.ndarray, seg_map: np.ndarray) -> None:
    """`<prefix>_f.npy` / `<prefix>_s.npy` exactly as preprocess.py:332-336 writes them."""
    np.save(prefix + "_s.npy", np.asarray(seg_map))
    np.save(prefix + "_f.npy", np.asarray(img_embed))


def load_feature_files(prefix: str):
    """(img_embed [n_seg, C], seg_map [4, h, w]) as dataset_readers.py:183-187 reads them."""
    if not os.path.exists(prefix + "_f.npy"):
        raise FileNotFoundError("Semantic feature file not found: " + prefix + "_f.npy")
    return np.load(prefix + "_f.npy"), np.load(prefix + "_s.npy")
