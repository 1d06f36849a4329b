"""Shim of the torch_scatter calls on the reference's path."""
import torch


def _scatter_arg(src, index, dim, dim_size, reduce):
    assert dim in (0, -src.dim()), "reference only scatters along dim 0"
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() else 0
    idx = index
    if idx.dim() < src.dim():
        idx = idx.view(-1, *([1] * (src.dim() - 1)))
    idx = idx.expand_as(src)
    out = torch.zeros((dim_size,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    out = out.scatter_reduce(0, idx, src, reduce=reduce, include_self=False)
    # arg: first position attaining the extremum; slots without input -> src.size(0)
    arg = torch.full_like(out, src.size(0), dtype=torch.long)
    hit = src == out.gather(0, idx)
    pos = torch.arange(src.size(0), device=src.device).view(-1, *([1] * (src.dim() - 1))).expand_as(src)
    pos = torch.where(hit, pos, torch.full_like(pos, src.size(0)))
    arg = arg.scatter_reduce(0, idx, pos, reduce="amin", include_self=True)
    return out, arg


def scatter_max(src, index, dim=-1, out=None, dim_size=None):
    if src.dim() == 1 and dim == -1:
        dim = 0
    return _scatter_arg(src, index, dim, dim_size, "amax")


def scatter_min(src, index, dim=-1, out=None, dim_size=None):
    if src.dim() == 1 and dim == -1:
        dim = 0
    return _scatter_arg(src, index, dim, dim_size, "amin")


def scatter_mean(*a, **k):
    raise NotImplementedError("not on the reference's inference path")


scatter_std = scatter_mean
