"""One plot on several GPUs (BASELINE.json configs[3]).

The inference path shards with no collective inside it: tiles are independent samples (eval-mode
BatchNorm uses running statistics) and every reference batch keeps its own voxel-grid origin.  One process
per GPU (`torch.distributed`, NCCL):

1. every rank holds the cloud and runs the (cheap, deterministic) tiling, so all ranks agree on the tiles
   and on the reference batches without any exchange;
2. rank r classifies a CONTIGUOUS range of batches balanced by point count (contiguous, so that the
   super-batches of `predicter.classify_tiles` still merge consecutive batches);
3. ONE all-gather of the classified rows (xyz fp32, prob fp32, pred uint8: 17 bytes per tile point) --
   the only exchange of the path, the device-side counterpart of the reference's `np.vstack(output_list)`
   (src/predicter.py:217);
4. rank r runs the spatial vote (src/predicter.py:107-142) for its slice of the original points against
   ALL classified rows, so the result equals the single-GPU one; the per-point (label, pwood) slices are
   all-gathered (9 bytes per point).

Works on NCCL (device tensors) and, for the exchange helpers, on gloo (host tensors, CPU tests).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import ops
from .predicter import classify_tiles, plan_batches
from .preprocessing import Voxelise

__all__ = ["shard_contiguous", "all_gather_rows", "classify_plot"]


def shard_contiguous(batches: Sequence[Tuple[int, int]], ptr: np.ndarray, world_size: int, rank: int) -> List[int]:
    """Batch indices [b0, b1) of `rank`: contiguous ranges whose point counts are as even as a prefix
    split allows (boundary i goes where the cumulative point count crosses i/world of the total).
    Deterministic on every rank; every batch belongs to exactly one rank."""
    pts = np.array([ptr[b] - ptr[a] for a, b in batches], dtype=np.int64)
    cum = np.concatenate([[0], np.cumsum(pts)])
    total = int(cum[-1])
    bounds = [int(np.searchsorted(cum, total * r / world_size, side="left")) for r in range(world_size)] + [len(batches)]
    bounds[0] = 0
    for r in range(1, world_size + 1):                       # monotone, in range
        bounds[r] = min(max(bounds[r], bounds[r - 1]), len(batches))
    return list(range(bounds[rank], bounds[rank + 1]))


def all_gather_rows(rows: torch.Tensor) -> torch.Tensor:
    """Concatenation over ranks (rank order) of tensors that differ in their first dimension."""
    import torch.distributed as dist
    world = dist.get_world_size()
    if world == 1:
        return rows
    count = torch.tensor([rows.size(0)], device=rows.device, dtype=torch.int64)
    counts = [torch.empty_like(count) for _ in range(world)]
    dist.all_gather(counts, count)
    counts = [int(c.item()) for c in counts]
    longest = max(counts)
    padded = rows.new_zeros((longest,) + tuple(rows.shape[1:]))
    padded[: rows.size(0)] = rows
    bucket = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(bucket, padded)
    return torch.cat([b[:c] for b, c in zip(bucket, counts)])


@torch.no_grad()
def classify_plot(net: torch.nn.Module, cloud: torch.Tensor, min_pts: int = 128, max_pts: int = 16384,
                  grid_size=(2.0, 4.0), batch_size: int = 8, is_wood: float = 0.5, any_wood: float = 1,
                  max_points_per_launch: int = 1 << 21, rank: Optional[int] = None, world_size: Optional[int] = None):
    """cloud [N, >=4] (x, y, z, reflectance) on this rank's device -> (label uint8 [N], pwood float64 [N])
    for the WHOLE plot on every rank.  Single process: rank 0 of 1."""
    import torch.distributed as dist
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        rank = dist.get_rank() if world_size > 1 else 0
    store = Voxelise(cloud, minpoints=min_pts, maxpoints=max_pts, gridsize=grid_size).write_voxels()
    batches = plan_batches(store.num_tiles, batch_size)
    mine = shard_contiguous(batches, store.ptr, world_size, rank)
    prob, pred, xyz, _ = classify_tiles(net, store, batch_size, is_wood, batch_ids=mine,
                                        max_points_per_launch=max_points_per_launch, want_xyz=True)
    if xyz is None or xyz.numel() == 0:
        xyz = torch.empty((0, 3), device=cloud.device, dtype=torch.float32)
        prob = torch.empty(0, device=cloud.device, dtype=torch.float32)
        pred = torch.empty(0, device=cloud.device, dtype=torch.uint8)
    if world_size > 1:
        packed = torch.cat([xyz, prob[:, None], pred[:, None].to(torch.float32)], dim=1)      # one exchange
        packed = all_gather_rows(packed)
        xyz, prob, pred = packed[:, :3].contiguous(), packed[:, 3].contiguous(), packed[:, 4].to(torch.uint8)
    n = cloud.size(0)
    lo, hi = rank * n // world_size, (rank + 1) * n // world_size
    k = 32 if any_wood != 1 else 64
    label, pwood = ops.spatial_vote(xyz, prob, pred, cloud[lo:hi, :3].contiguous(), k, float(any_wood))
    if world_size > 1:
        label = all_gather_rows(label)
        pwood = all_gather_rows(pwood)
    return label, pwood
