"""torch.ops-compatible entry points: the dispatcher-level schemas of torch_cluster / torch_scatter
 registered under the `p2w::` namespace on the CUDA key, so code that calls
`torch.ops.torch_cluster.knn(...)` can be pointed at `torch.ops.p2w.knn(...)` unchanged.

    p2w::knn(Tensor x, Tensor y, Tensor? ptr_x, Tensor? ptr_y, int k, bool cosine, int num_workers) -> Tensor
    p2w::radius(Tensor x, Tensor y, Tensor? ptr_x, Tensor? ptr_y, float r, int max_num_neighbors,
                int num_workers, bool ignore_same_index) -> Tensor
    p2w::fps(Tensor src, Tensor ptr, Tensor ratio, bool random_start) -> Tensor
    p2w::grid(Tensor pos, Tensor size, Tensor? start, Tensor? end) -> Tensor
    p2w::scatter_max(Tensor src, Tensor index, int dim, Tensor? out, int? dim_size) -> (Tensor, Tensor)
    p2w::scatter_min(...)                                                           -> (Tensor, Tensor)

There is no CPU registration on purpose: a CPU tensor reaches the op and raises.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import Tensor

from . import _lib, ops

_lib_def = torch.library.Library("p2w", "DEF")
_lib_def.define("knn(Tensor x, Tensor y, Tensor? ptr_x, Tensor? ptr_y, int k, bool cosine, int num_workers) -> Tensor")
_lib_def.define("radius(Tensor x, Tensor y, Tensor? ptr_x, Tensor? ptr_y, float r, int max_num_neighbors, "
                "int num_workers, bool ignore_same_index) -> Tensor")
_lib_def.define("fps(Tensor src, Tensor ptr, Tensor ratio, bool random_start) -> Tensor")
_lib_def.define("grid(Tensor pos, Tensor size, Tensor? start, Tensor? end) -> Tensor")
_lib_def.define("scatter_max(Tensor src, Tensor index, int dim, Tensor? out, int? dim_size) -> (Tensor, Tensor)")
_lib_def.define("scatter_min(Tensor src, Tensor index, int dim, Tensor? out, int? dim_size) -> (Tensor, Tensor)")


def _default_ptr(ptr: Optional[Tensor], n: int, dev) -> Tensor:
    return torch.tensor([0, n], device=dev, dtype=torch.int64) if ptr is None else ptr


def _knn(x, y, ptr_x, ptr_y, k, cosine, num_workers):
    if cosine:
        raise _lib.P2WError("knn: cosine distance is not on the PointsToWood path")
    if k > 100:
        raise _lib.P2WError("knn: k must be <= 100")
    return ops.table_to_edge_index(ops.knn_table(x, y, k, _default_ptr(ptr_x, x.size(0), x.device),
                                                 _default_ptr(ptr_y, y.size(0), x.device)))


def _radius(x, y, ptr_x, ptr_y, r, max_num_neighbors, num_workers, ignore_same_index):
    if ignore_same_index:
        raise _lib.P2WError("radius: ignore_same_index is not on the PointsToWood path")
    nbr, _ = ops.radius_table(x, y, r, _default_ptr(ptr_x, x.size(0), x.device),
                              _default_ptr(ptr_y, y.size(0), x.device), max_num_neighbors)
    return ops.table_to_edge_index(nbr)


def _fps(src, ptr, ratio, random_start):
    return ops.fps(src, ratio=float(ratio.reshape(-1)[0].item()), random_start=random_start, ptr=ptr)


def _grid(pos, size, start, end):
    return ops.grid_cluster(pos, size, start, end)


def _scatter_max(src, index, dim, out, dim_size) -> Tuple[Tensor, Tensor]:
    return ops.scatter_max(src, index, dim, out, dim_size)


def _scatter_min(src, index, dim, out, dim_size) -> Tuple[Tensor, Tensor]:
    return ops.scatter_min(src, index, dim, out, dim_size)


_lib_impl = torch.library.Library("p2w", "IMPL", "CUDA")
for _name, _fn in (("knn", _knn), ("radius", _radius), ("fps", _fps), ("grid", _grid),
                   ("scatter_max", _scatter_max), ("scatter_min", _scatter_min)):
    _lib_impl.impl(_name, _fn)
