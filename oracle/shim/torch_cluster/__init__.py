"""Shim of torch_cluster's Python API on the CPU oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from _common import O


def _ptr(batch, n, batch_size=None):
    if batch is None:
        return np.array([0, n], dtype=np.int64)
    return O.batch_to_ptr(batch.cpu().numpy(), batch_size)


def knn(x, y, k, batch_x=None, batch_y=None, cosine=False, num_workers=1, batch_size=None):
    assert not cosine
    if batch_size is None and batch_x is not None:
        batch_size = int(max(batch_x.max(), batch_y.max())) + 1
    px, py = _ptr(batch_x, x.size(0), batch_size), _ptr(batch_y, y.size(0), batch_size)
    nbr = O.knn(x.detach().cpu().numpy(), y.detach().cpu().numpy(), k, px, py)
    return torch.from_numpy(O.table_to_edges(nbr))


def radius(x, y, r, batch_x=None, batch_y=None, max_num_neighbors=32, num_workers=1, batch_size=None):
    if batch_size is None and batch_x is not None:
        batch_size = int(max(batch_x.max(), batch_y.max())) + 1
    px, py = _ptr(batch_x, x.size(0), batch_size), _ptr(batch_y, y.size(0), batch_size)
    nbr, _ = O.radius(x.detach().cpu().numpy(), y.detach().cpu().numpy(), r, px, py, max_num_neighbors)
    return torch.from_numpy(O.table_to_edges(nbr))


def grid_cluster(pos, size, start=None, end=None):
    return torch.from_numpy(O.grid(pos.cpu().numpy(), size.cpu().numpy(),
                                   None if start is None else start.cpu().numpy(),
                                   None if end is None else end.cpu().numpy()))


def fps(src, batch=None, ratio=0.5, random_start=True, batch_size=None, ptr=None):
    assert not random_start, "oracle pins random_start=False"
    p = _ptr(batch, src.size(0), batch_size) if ptr is None else ptr.cpu().numpy()
    return torch.from_numpy(O.fps(src.cpu().numpy(), p, float(ratio)))
