"""ctypes/numpy front end of the CPU oracle (oracle/p2w_oracle.c).

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's CPU legs may
import this.  PARITY UNPINNED for the third-party primitives (see the header of
p2w_oracle.c): the reference has no tests or golden vectors, torch_cluster /
torch_scatter / torch_geometric are absent, so the semantics follow SURVEY.md Appendix A
and are anchored on the reference's call sites (cited per function).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libp2w_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "p2w_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libp2w_oracle.so"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
    return _lib


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def batch_to_ptr(batch, batch_size=None):
    """ptr = bucketize(arange(B+1), batch) for a sorted batch vector."""
    batch = _i64(batch)
    if batch_size is None:
        batch_size = int(batch.max()) + 1 if batch.size else 0
    return np.searchsorted(batch, np.arange(batch_size + 1), side="left").astype(np.int64)


# "brute": the C restatement of torch_cluster's CUDA scan (the parity checker).  "kdtree": one KD-tree per
# tile, single-threaded per call, leaf size 10 -- the cost profile of torch_cluster's CPU back end (nanoflann,
# SURVEY.md Appendix A.2/A.3), used ONLY by bench.py's timed CPU arm; same neighbour sets up to ties / last ulp.
SEARCH = "brute"


def _kdtree_search(x, y, k, ptr_x, ptr_y, r=None):
    from scipy.spatial import cKDTree
    nbr = np.full((y.shape[0], k), -1, dtype=np.int64)
    d2 = np.full((y.shape[0], k), 1e10, dtype=np.float32)
    for b in range(len(ptr_x) - 1):
        x0, x1, y0, y1 = int(ptr_x[b]), int(ptr_x[b + 1]), int(ptr_y[b]), int(ptr_y[b + 1])
        if x1 == x0 or y1 == y0:
            continue
        tree = cKDTree(x[x0:x1], leafsize=10)
        kk = min(k, x1 - x0)
        d, i = tree.query(y[y0:y1], k=kk, workers=1, **({} if r is None else {"distance_upper_bound": r}))
        d, i = d.reshape(y1 - y0, kk), i.reshape(y1 - y0, kk)
        ok = i < (x1 - x0)
        if r is not None:                       # radius: the hits, ascending index (CUDA order)
            i = np.sort(np.where(ok, i, x1 - x0), axis=1)
            ok = i < (x1 - x0)
        nbr[y0:y1, :kk] = np.where(ok, i + x0, -1)
        d2[y0:y1, :kk] = np.where(ok, (d * d).astype(np.float32), np.float32(1e10))
    return nbr, d2


def knn(x, y, k, ptr_x=None, ptr_y=None, return_d2=False):
    """[Ny,k] int64 neighbour table, -1 padded (model.py:120,149; Appendix A.2)."""
    x, y = _f32(x), _f32(y)
    ptr_x = _i64([0, x.shape[0]]) if ptr_x is None else _i64(ptr_x)
    ptr_y = _i64([0, y.shape[0]]) if ptr_y is None else _i64(ptr_y)
    if SEARCH == "kdtree":
        nbr, d2 = _kdtree_search(x, y, k, ptr_x, ptr_y)
        return (nbr, d2) if return_d2 else nbr
    nbr = np.empty((y.shape[0], k), dtype=np.int64)
    d2 = np.empty((y.shape[0], k), dtype=np.float32)
    rc = lib().orc_knn(_p(x, ctypes.c_float), _p(y, ctypes.c_float), _p(ptr_x, ctypes.c_int64),
                       _p(ptr_y, ctypes.c_int64), len(ptr_x) - 1, x.shape[1], k,
                       _p(nbr, ctypes.c_int64), _p(d2, ctypes.c_float))
    if rc != 0:
        raise RuntimeError("knn: k must be in [1, 100]")
    return (nbr, d2) if return_d2 else nbr


def radius(x, y, r, ptr_x=None, ptr_y=None, max_num_neighbors=32):
    """([Ny,max] int64 -1 padded, cnt [Ny]) (model.py:118; Appendix A.3, CUDA semantics)."""
    x, y = _f32(x), _f32(y)
    ptr_x = _i64([0, x.shape[0]]) if ptr_x is None else _i64(ptr_x)
    ptr_y = _i64([0, y.shape[0]]) if ptr_y is None else _i64(ptr_y)
    if SEARCH == "kdtree":
        nbr, _ = _kdtree_search(x, y, max_num_neighbors, ptr_x, ptr_y, r=float(r))
        return nbr, (nbr >= 0).sum(1).astype(np.int32)
    nbr = np.empty((y.shape[0], max_num_neighbors), dtype=np.int64)
    cnt = np.empty(y.shape[0], dtype=np.int32)
    lib().orc_radius(_p(x, ctypes.c_float), _p(y, ctypes.c_float), _p(ptr_x, ctypes.c_int64),
                     _p(ptr_y, ctypes.c_int64), len(ptr_x) - 1, x.shape[1], ctypes.c_double(r),
                     max_num_neighbors, _p(nbr, ctypes.c_int64), _p(cnt, ctypes.c_int32))
    return nbr, cnt


def table_to_edges(nbr):
    """[Ny,K] -1-padded table -> upstream's [2,E] layout (row0 = y index, row1 = x index)."""
    mask = nbr >= 0
    row = np.broadcast_to(np.arange(nbr.shape[0], dtype=np.int64)[:, None], nbr.shape)[mask]
    return np.stack([row, nbr[mask]])


def fps(src, ptr=None, ratio=0.5):
    """torch_cluster.fps(random_start=False) with lowest-index ties (Appendix A.11)."""
    src = _f32(src)
    ptr = _i64([0, src.shape[0]]) if ptr is None else _i64(ptr)
    B = len(ptr) - 1
    n = np.diff(ptr)
    m = np.ceil(n.astype(np.float32) * np.float32(ratio)).astype(np.int64)
    out = np.empty(int(m.sum()), dtype=np.int64)
    out_ptr = np.empty(B + 1, dtype=np.int64)
    lib().orc_fps(_p(src, ctypes.c_float), _p(ptr, ctypes.c_int64), B, src.shape[1],
                  ctypes.c_double(ratio), _p(out, ctypes.c_int64), _p(out_ptr, ctypes.c_int64))
    return out


def grid(pos, size, start=None, end=None):
    """torch_cluster.grid_cluster (Appendix A.4): start/end default to the column min/max."""
    pos = _f32(pos)
    size = _f32(size)
    start = pos.min(0) if start is None else _f32(start)
    end = pos.max(0) if end is None else _f32(end)
    out = np.empty(pos.shape[0], dtype=np.int64)
    lib().orc_grid(_p(pos, ctypes.c_float), pos.shape[1], _p(size, ctypes.c_float),
                   _p(_f32(start), ctypes.c_float), _p(_f32(end), ctypes.c_float),
                   pos.shape[0], _p(out, ctypes.c_int64))
    return out


def voxel_grid(pos, size, batch=None):
    """torch_geometric.nn.voxel_grid (Appendix A.4): batch appended as an FP32 column of size 1."""
    pos = _f32(pos)
    if pos.ndim == 1:
        pos = pos[:, None]
    b = np.zeros(pos.shape[0], np.float32) if batch is None else np.asarray(batch).astype(np.float32)
    p = np.concatenate([pos, b[:, None]], 1)
    sz = np.concatenate([np.full(pos.shape[1], size, np.float32), np.ones(1, np.float32)])
    return grid(p, sz)


def consecutive_cluster(src):
    """Appendix A.5: (inverse, perm); perm[u] = HIGHEST member index of cluster u (CPU
    scatter_ order), clusters ascending by id."""
    src = _i64(src)
    uniq, inv = np.unique(src, return_inverse=True)
    perm = np.empty(len(uniq), dtype=np.int64)
    perm[inv] = np.arange(len(src), dtype=np.int64)      # later writes win -> highest index
    return inv.astype(np.int64), perm


def scatter_max(src, index, dim_size=None):
    """torch_scatter.scatter_max along dim 0 (Appendix A.7): empty slots -> 0, arg = len(src)."""
    src = np.asarray(src)
    index = _i64(index)
    n = int(index.max()) + 1 if dim_size is None else dim_size
    out = np.full((n,) + src.shape[1:], -np.inf, dtype=src.dtype)
    np.maximum.at(out, index, src)
    out[np.isneginf(out)] = 0
    return out
