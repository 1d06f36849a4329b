"""Point-cloud file I/O with the reference's function names (src/io.py:11-224): `load_file`, `save_file`,
`read_ply` / `write_ply`, `read_pcd` / `write_pcd`.  SURVEY.md §8(f)-4.

Host-side code, nothing here touches the GPU.  Differences in construction, not in results:
* the PLY header is parsed from BYTES (the reference decodes the whole file as ISO-8859-1 text to find the
  header length) and both byte orders are accepted;
* the PLY writer streams the rows in chunks of `CHUNK_ROWS` through one structured buffer, so a 100 M-point
  plot with 3 + k float64 columns (5.6 GB) never exists twice in host memory; the bytes
  written are identical to the reference's `to_records().tobytes()` (tests/test_io.py pins this against a
  file written by the reference itself);
* LAS / LAZ need `laspy`, which this image does not ship: a clear error instead of an ImportError deep inside.
"""
from __future__ import annotations

import os
from typing import Iterable, List, Sequence

import numpy as np
import pandas as pd

__all__ = ["read_ply", "write_ply", "read_pcd", "write_pcd", "load_file", "save_file", "CHUNK_ROWS"]

CHUNK_ROWS = 1 << 20

_PLY_TYPES = {"char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1", "short": "i2", "int16": "i2",
              "ushort": "u2", "uint16": "u2", "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4",
              "float": "f4", "float32": "f4", "double": "f8", "float64": "f8"}


def _ply_header(raw: bytes):
    """(header length in bytes, format, vertex count, [(name, numpy code)]) of a PLY file's first bytes."""
    end = raw.find(b"end_header")
    if not raw.startswith(b"ply") or end < 0:
        raise Exception("not a PLY file (no 'ply' magic / 'end_header')")
    stop = raw.index(b"\n", end) + 1
    fmt, count, props, in_vertex = None, None, [], False
    for line in raw[:stop].decode("latin-1").splitlines():
        tok = line.split()
        if not tok:
            continue
        if tok[0] == "format":
            fmt = tok[1]
        elif tok[0] == "element":
            if tok[1] == "face" and int(tok[2]) > 0:
                raise Exception(".ply appears to be a mesh")                      # src/io.py:35-36
            in_vertex = tok[1] == "vertex"
            if in_vertex:
                count = int(tok[2])
        elif tok[0] == "property" and in_vertex:
            if tok[1] == "list":
                raise Exception(".ply vertex elements with list properties are not supported")
            props.append((tok[2], _PLY_TYPES[tok[1]]))
    if fmt is None or count is None or not props:
        raise Exception("incomplete PLY header")
    return stop, fmt, count, props


def read_ply(fp, newline=None) -> pd.DataFrame:
    """src/io.py:11-47: one DataFrame column per vertex property, native dtypes kept."""
    with open(fp, "rb") as f:
        head = f.read(1 << 16)
        while b"end_header" not in head:
            more = f.read(1 << 16)
            if not more:
                break
            head += more
        stop, fmt, count, props = _ply_header(head)
        f.seek(stop)
        if fmt == "ascii":
            arr = np.loadtxt(f, ndmin=2)
            if arr.shape[0] != count or arr.shape[1] != len(props):
                raise Exception(f"ascii PLY body is {arr.shape}, header says {count} x {len(props)}")
            return pd.DataFrame({name: arr[:, i] for i, (name, _) in enumerate(props)})
        order = "<" if fmt == "binary_little_endian" else ">"
        rec = np.dtype([(name, order + code) for name, code in props])
        arr = np.fromfile(f, dtype=rec, count=count)
    if arr.shape[0] != count:
        raise Exception(f"PLY body holds {arr.shape[0]} vertices, header says {count}")
    return pd.DataFrame({name: arr[name].astype(arr[name].dtype.newbyteorder("=")) for name, _ in props})


def _ply_columns(pc: pd.DataFrame):
    """Column order and on-disk type of write_ply (src/io.py:49-86): x y z float64, then red green blue as
    int when present, then every other column that converts to float64, in frame order."""
    cols = [("x", "<f8"), ("y", "<f8"), ("z", "<f8")]
    if "red" in pc.columns:
        cols += [("red", "<i4"), ("green", "<i4"), ("blue", "<i4")]
    taken = {c for c, _ in cols}
    for col in pc.columns:
        if col in taken:
            continue
        if pc[col].dtype.kind not in "biuf":          # numeric columns always convert; anything else is tried whole
            try:
                pc[col].astype("float64")
            except (TypeError, ValueError):
                continue                                                            # src/io.py:82-83: silently skipped
        cols.append((col, "<f8"))
        taken.add(col)
    return cols


def write_ply(output_name, pc: pd.DataFrame, comments: Iterable[str] = ()) -> None:
    """src/io.py:49-86, byte-identical output; the body is streamed in CHUNK_ROWS-row pieces."""
    cols = _ply_columns(pc)
    names = {"<f8": "float64", "<i4": "int"}
    header = ["ply", "format binary_little_endian 1.0", "comment Author: Phil Wilkes"]
    header += [f"comment {c}" for c in comments]
    header += ["obj_info generated with pcd2ply.py", f"element vertex {len(pc)}"]
    header += [f"property {names[code]} {col}" for col, code in cols]
    header.append("end_header")
    rec = np.dtype(cols)
    with open(output_name, "wb") as f:
        f.write(("\n".join(header) + "\n").encode("ascii"))
        n = len(pc)
        buf = np.empty(min(n, CHUNK_ROWS), dtype=rec)
        source = {col: (pc[col].to_numpy() if pc[col].dtype.kind in "biuf" else pc[col].astype("float64").to_numpy())
                  for col, _ in cols}
        for lo in range(0, n, CHUNK_ROWS):
            hi = min(lo + CHUNK_ROWS, n)
            piece = buf[: hi - lo]
            for col, _ in cols:
                piece[col] = source[col][lo:hi]
            piece.tofile(f)


_PCD_TYPES = {("F", 4): "f4", ("F", 8): "f8", ("I", 1): "i1", ("I", 2): "i2", ("I", 4): "i4", ("U", 1): "u1",
              ("U", 2): "u2", ("U", 4): "u4"}


def read_pcd(fp) -> pd.DataFrame:
    """src/io.py:90-119 (binary float fields, or ascii)."""
    with open(fp, "rb") as f:
        raw = f.read()
    pos, fields, sizes, types, width, fmt = 0, None, None, None, None, None
    while fmt is None:
        nl = raw.index(b"\n", pos)
        tok = raw[pos:nl].decode("latin-1").split()
        pos = nl + 1
        if not tok or tok[0].startswith("#"):
            continue
        key = tok[0].upper()
        if key == "FIELDS":
            fields = tok[1:]
        elif key == "SIZE":
            sizes = [int(t) for t in tok[1:]]
        elif key == "TYPE":
            types = tok[1:]
        elif key in ("WIDTH", "POINTS"):
            width = int(tok[1]) if width is None or key == "POINTS" else width
        elif key == "DATA":
            fmt = tok[1]
    if fields is None or width is None:
        raise Exception("incomplete PCD header")
    if fmt == "ascii":
        arr = np.loadtxt(raw[pos:].decode("latin-1").splitlines(), ndmin=2)
        return pd.DataFrame(arr[:width, : len(fields)], columns=fields)
    if fmt != "binary":
        raise Exception(f"PCD DATA {fmt} is not supported")
    if sizes is None or types is None:
        sizes, types = [4] * len(fields), ["F"] * len(fields)                        # what the reference assumes
    rec = np.dtype([(name, "<" + _PCD_TYPES[(t, s)]) for name, t, s in zip(fields, types, sizes)])
    arr = np.frombuffer(raw, dtype=rec, count=width, offset=pos)
    return pd.DataFrame({name: arr[name] for name in fields})


def write_pcd(df: pd.DataFrame, path, binary: bool = True) -> None:
    """src/io.py:121-145: x y z (+ intensity) as binary float32."""
    df = df.rename(columns={"scalar_intensity": "intensity"})
    columns = ["x", "y", "z"] + (["intensity"] if "intensity" in df.columns else [])
    k = len(columns)
    header = ["# .PCD v0.7 - Point Cloud Data file format", "VERSION 0.7", "FIELDS " + " ".join(columns) + " ",
              "SIZE " + "4 " * k, "TYPE " + "F " * k, "COUNT " + "1 " * k, f"WIDTH {len(df)}", "HEIGHT 1",
              "VIEWPOINT 0 0 0 1 0 0 0", f"POINTS {len(df)}", "DATA binary"]
    with open(path, "wb") as f:
        f.write(("\n".join(header) + "\n").encode("ascii"))
        for lo in range(0, len(df), CHUNK_ROWS):
            df[columns].iloc[lo: lo + CHUNK_ROWS].to_numpy(dtype="<f4").tofile(f)


def load_file(filename, additional_headers: bool = False, verbose: bool = False):
    """src/io.py:152-181."""
    ext = os.path.splitext(filename)[1].lower()
    if ext in (".las", ".laz"):
        try:
            import laspy
        except ImportError as e:
            raise Exception("reading .las / .laz needs the laspy package") from e
        las = laspy.read(filename)
        pc = pd.DataFrame({"x": np.asarray(las.x), "y": np.asarray(las.y), "z": np.asarray(las.z)})
    elif ext == ".ply":
        pc = read_ply(filename)
    elif ext == ".pcd":
        pc = read_pcd(filename)
    else:
        raise Exception("point cloud format not recognised" + filename)
    if verbose:
        print(f"read in {filename} with {len(pc)} points")
    if additional_headers:
        return pc, [c for c in pc.columns if c not in ("x", "y", "z")]
    return pc


def save_file(filename, pointcloud, additional_fields: Sequence[str] = (), verbose: bool = False) -> None:
    """src/io.py:184-224: .ply (DataFrame or [N, 3+k] array), .csv; .las needs laspy."""
    fields: List[str] = ["x", "y", "z"] + [f for f in additional_fields if f not in ("x", "y", "z")]
    if verbose:
        print("Saving file:", filename)
    if filename.endswith(".ply"):
        if not isinstance(pointcloud, pd.DataFrame):
            pointcloud = pd.DataFrame(np.asarray(pointcloud), columns=fields)
        write_ply(filename, pointcloud[[f for f in fields if f in pointcloud.columns]])
    elif filename.endswith(".csv"):
        pd.DataFrame(pointcloud).to_csv(filename, header=None, index=None, sep=" ")
    elif filename.endswith(".pcd"):
        if not isinstance(pointcloud, pd.DataFrame):
            pointcloud = pd.DataFrame(np.asarray(pointcloud), columns=fields)
        write_pcd(pointcloud, filename)
    elif filename.endswith(".las") or filename.endswith(".laz"):
        try:
            import laspy
        except ImportError as e:
            raise Exception("writing .las / .laz needs the laspy package") from e
        arr = pointcloud[fields].to_numpy() if isinstance(pointcloud, pd.DataFrame) else np.asarray(pointcloud)
        las = laspy.create(file_version="1.4", point_format=7)
        las.header.offsets = np.min(arr[:, :3], axis=0)
        las.header.scales = [0.001, 0.001, 0.001]
        las.x, las.y, las.z = arr[:, 0], arr[:, 1], arr[:, 2]
        for i, name in enumerate(fields[3:], start=3):
            if name not in ("red", "green", "blue"):
                las.add_extra_dim(laspy.ExtraBytesParams(name=name, type="f8"))
            setattr(las, name, arr[:, i])
        las.write(filename)
    else:
        raise Exception("output format not recognised" + filename)
