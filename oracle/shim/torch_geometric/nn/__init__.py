import torch
from torch import Tensor
from torch_cluster import knn, radius, grid_cluster, fps  # noqa: F401
from .conv import MessagePassing


class PointNetConv(MessagePassing):   # the reference imports it, then shadows it with src.pointnet
    pass


def voxel_grid(pos, size, batch=None, start=None, end=None):
    """Appendix A.4."""
    pos = pos.unsqueeze(-1) if pos.dim() == 1 else pos
    dim = pos.size(1)
    if batch is None:
        batch = pos.new_zeros(pos.size(0), dtype=torch.long)
    pos = torch.cat([pos, batch.view(-1, 1).to(pos.dtype)], dim=-1)
    if not isinstance(size, Tensor):
        size = torch.tensor(size, dtype=pos.dtype, device=pos.device)
    size = size.repeat(dim) if size.numel() == 1 else size
    size = torch.cat([size, size.new_ones(1)])
    if start is not None:
        start = torch.cat([torch.as_tensor(start, dtype=pos.dtype).view(-1), pos.new_zeros(1)])
    if end is not None:
        end = torch.cat([torch.as_tensor(end, dtype=pos.dtype).view(-1), batch.max().to(pos.dtype).view(1)])
    return grid_cluster(pos, size, start, end)


def global_max_pool(x, batch, size=None):
    """Appendix A.9."""
    size = int(batch.max()) + 1 if size is None else size
    out = x.new_zeros((size, x.size(1)))
    return out.scatter_reduce(0, batch.view(-1, 1).expand_as(x), x, reduce="amax", include_self=False)


def knn_interpolate(x, pos_x, pos_y, batch_x=None, batch_y=None, k=3, num_workers=1):
    """Appendix A.8."""
    with torch.no_grad():
        assign = knn(pos_x, pos_y, k, batch_x=batch_x, batch_y=batch_y)
        y_idx, x_idx = assign[0], assign[1]
        diff = pos_x[x_idx] - pos_y[y_idx]
        sq = (diff * diff).sum(dim=-1, keepdim=True)
        w = 1.0 / torch.clamp(sq, min=1e-16)
    num = x.new_zeros((pos_y.size(0), x.size(1))).index_add_(0, y_idx, x[x_idx] * w)
    den = w.new_zeros((pos_y.size(0), 1)).index_add_(0, y_idx, w)
    return num / den
