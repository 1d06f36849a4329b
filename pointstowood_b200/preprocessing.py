"""Tiling front end: the B200-native counterpart of src/preprocessing.py (Voxelise / preprocess).

Same class, method and argument names as the reference, but everything stays on the device:
instead of one `voxels/voxel_k.pt` file per tile (src/preprocessing.py:125) the result is a
`TileStore` -- the 5-column feature array plus a CSR of member indices -- which
`predicter.SemanticSegmentation` consumes directly.

Pinned choices where the reference is random or order-unstable (the CPU
oracle oracle/ref_pipeline.py pins the same ones):
* reflectance ranks come from a STABLE sort (C.9);
* tiles with more than `maxpoints` members are thinned with the reference's sampling LAWS on a
  counter-based hash of (seed, point index) instead of torch's global generator (C.5): with
  reflectance, weighted sampling without replacement (torch.multinomial, :118) through
  Efraimidis-Spirakis keys -log(u)/w, rows in draw order; without reflectance, `maxpoints` uniform
  draws WITH replacement (torch.randint, :120);
* tiles are listed 2 m voxels first, then 4 m voxels, each by ascending voxel id (:57-63).
Files with further scalar columns (rgb, deviation, ...) are voxelised over those columns too, as the reference does.
Non-finite input: NaN reflectance raises ValueError as the reference does (:20-21); rows with a
non-finite coordinate join no tile (the reference's per-tile NaN row filter, :123) and get
n_z = NaN; `Voxelise.finite_rows` lists the rows that were tiled (None: all of them).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np
import torch
from torch import Tensor

from . import _lib, ops

__all__ = ["TileStore", "Voxelise", "preprocess", "SUBSAMPLE_SEED"]

SUBSAMPLE_SEED = 141190          # the reference's global seed (src/trainer.py:25-26)


@dataclass
class TileStore:
    feat: Tensor            # [N,5] float32 device: x, y, z, reflectance (normalised), n_z
    members: Tensor         # [M] int64 device: point ids, tile-major, reference row order inside a tile
    ptr: np.ndarray         # [T+1] int64 host: tile t owns members[ptr[t]:ptr[t+1]]
    grid_of_tile: np.ndarray  # [T] float32 host: the grid size that produced the tile
    finite_rows: Optional[Tensor] = None   # rows of the input cloud that `feat` holds (None: all; see Voxelise)

    @property
    def num_tiles(self) -> int:
        return len(self.ptr) - 1

    @property
    def sizes(self) -> np.ndarray:
        return np.diff(self.ptr)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def thin_tiles(feat: Tensor, refl_col: int, members: Tensor, sizes: np.ndarray, global_index: Optional[Tensor],
               refl_min: float, weighted: bool, voxel_ids: Tensor, maxpoints: int, seed: int, grid_ordinal: int) -> Tensor:
    """src/preprocessing.py:116-120 for a group of oversized tiles whose member rows (int32 rows of `feat`,
    ascending point index inside a tile) are concatenated in `members`, tile t holding sizes[t] of them.
    Returns int32 [len(sizes) * maxpoints]: the sampled rows of every tile, in draw order.
    weighted: Efraimidis-Spirakis order of the weights feat[:, refl_col] - refl_min + 1e-8 (torch.multinomial
    without replacement); else uniform draws with replacement (torch.randint).  `global_index` (int32 per row of
    feat, or None when rows ARE point indices) and `voxel_ids` (int64 per tile) feed the counter-based hash, so
    a plot gives the same sample whether it is tiled on one GPU or sharded over several."""
    dev = feat.device
    L = _lib.lib()
    members = members.to(torch.int32).contiguous()
    nbig = len(sizes)
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    picks = torch.empty(nbig * maxpoints, device=dev, dtype=torch.int32)
    if not weighted:
        seg = _lib.to_device(offs, dev, np.int64)
        _lib.check(L.p2w_replacement_picks(members.data_ptr(), seg.data_ptr(), voxel_ids.data_ptr(), nbig, maxpoints,
                                           (seed & 0xFFFFFFFF) | (grid_ordinal << 32), picks.data_ptr(), _stream()))
        return picks
    keys = torch.empty(members.numel(), device=dev, dtype=torch.int64)
    _lib.check(L.p2w_sampling_keys(feat.data_ptr(), feat.stride(0), refl_col, members.data_ptr(),
                                   None if global_index is None else global_index.data_ptr(), members.numel(),
                                   float(refl_min), seed & 0xFFFFFFFF, keys.data_ptr(), _stream()))
    _, by_key = ops.sort_pairs(keys, 64)                                  # all members by ascending key ...
    tile_of = torch.repeat_interleave(torch.arange(nbig, device=dev, dtype=torch.int64), _lib.to_device(sizes, dev, np.int64),
                                      output_size=int(offs[-1]))
    _, by_tile = ops.sort_pairs(tile_of[by_key.long()].contiguous(), max(1, int(nbig).bit_length()), values=by_key)
    take = (_lib.to_device(offs[:-1], dev, np.int64)[:, None] + torch.arange(maxpoints, device=dev)[None, :]).reshape(-1)
    return members[by_tile[take].long()]                                  # ... then stably by tile: (tile, key) order


class Voxelise:
    def __init__(self, pos, vxpath=None, minpoints=512, maxpoints=16384, gridsize=(2.0, 4.0), pointspacing=0.01,
                 seed: int = SUBSAMPLE_SEED):
        self.pos = pos
        self.vxpath = vxpath
        self.minpoints = minpoints
        self.maxpoints = maxpoints
        self.gridsize = list(gridsize)
        self.pointspacing = min(self.gridsize) / 100.0
        self.seed = seed
        self.cloud: Optional[Tensor] = None
        self.n_z: Optional[Tensor] = None
        self.finite_rows: Optional[Tensor] = None
        self.refl: Optional[Tensor] = None
        self._stats = None

    # ---- src/preprocessing.py:37-53
    def _cloud_stats(self):
        """Column min / max of (x, y, z, reflectance) and the number of non-finite entries per column: ONE host
        round trip serves the ground grid, the `reflectance != 0` test of :94 (any non-zero <=> min or max
        non-zero) and the NaN checks (:20-21, :123)."""
        if self._stats is None:
            mn, mx = ops._colminmax(self.cloud[:, :4])
            bad = (~torch.isfinite(self.cloud[:, :4])).sum(0).to(torch.float32)
            nan_refl = torch.isnan(self.cloud[:, 3]).sum().to(torch.float32).view(1)
            host = torch.cat([mn, mx, bad, nan_refl]).cpu().numpy()
            self._stats = (mn, host[:8].reshape(2, 4), host[8:12], host[12])
        return self._stats

    def gpu_ground(self) -> Tensor:
        cloud = self.cloud
        n = cloud.size(0)
        L = _lib.lib()
        mn, ext = self._cloud_stats()[:2]
        lo = ext[0, :2]
        hi = ext[1, :2] + np.float32(5.0)      # x_max + grid_resolution, rounded in fp32 like the tensor op
        nb = [max(1, int(math.ceil((float(hi[d]) - float(lo[d])) / 5.0))) for d in range(2)]
        cell_min = torch.empty((nb[0] + 1) * (nb[1] + 1), device=cloud.device, dtype=torch.float32)
        n_z = torch.empty(n, device=cloud.device, dtype=torch.float32)
        _lib.check(L.p2w_ground_normalize(cloud.data_ptr(), cloud.stride(0), n, mn.data_ptr(), 5.0, nb[0], nb[1],
                                          cell_min.data_ptr(), n_z.data_ptr(), _stream()))
        self.n_z = n_z
        return n_z

    # ---- src/preprocessing.py:18-30
    def quantile_normalize_reflectance(self) -> Tensor:
        cloud = self.cloud
        n = cloud.size(0)
        L = _lib.lib()
        keys = torch.empty(n, device=cloud.device, dtype=torch.int64)
        _lib.check(L.p2w_reflectance_keys(cloud.data_ptr(), cloud.stride(0), 3, n, keys.data_ptr(), _stream()))
        _, order = ops.sort_pairs(keys, 32)
        v = torch.empty(n, device=cloud.device, dtype=torch.float32)
        out = torch.empty(n, device=cloud.device, dtype=torch.float32)
        mnmx = torch.empty(2, device=cloud.device, dtype=torch.float32)
        _lib.check(L.p2w_reflectance_normalize(order.data_ptr(), n, v.data_ptr(), mnmx.data_ptr(), out.data_ptr(),
                                               _stream()))
        return out

    # ---- src/preprocessing.py:55-64: per grid size, the member lists of voxels with >= minpoints
    def grid(self, feat: Tensor):
        out = []
        n = feat.size(0)
        if feat.size(1) <= 8:
            mn, mx = ops._colminmax(feat)
        else:                                                           # more scalar fields than the kernel's eight columns
            mn, mx = feat.min(0).values.contiguous(), feat.max(0).values.contiguous()
        ext = torch.stack([mn, mx]).cpu().numpy()                       # one round trip for every grid size
        self._feat_ext = ext
        L = _lib.lib()
        queued = []
        for size in self.gridsize:                                       # all grid sizes are enqueued first ...
            cells = 1
            for d in range(feat.size(1)):
                cells *= int(np.float32(ext[1, d] - ext[0, d]) / np.float32(size)) + 1
            bits = max(1, int(cells).bit_length())
            sz = torch.full((feat.size(1),), float(size), device=feat.device, dtype=torch.float32)
            ids = ops.grid_cluster(feat, sz, mn, mx)
            keys, order = ops.sort_pairs(ids, bits)
            # voxel count and segment starts side by side: ONE device-to-host copy of [count | starts[:bound + 1]]
            buf = torch.empty(n + 2, device=feat.device, dtype=torch.int64)
            ws = torch.empty(max(int(L.p2w_unique_ws_bytes(n)), 8), device=feat.device, dtype=torch.uint8)
            _lib.check(L.p2w_unique_last(keys.data_ptr(), order.data_ptr(), n, None, None, buf[1:].data_ptr(),
                                         buf.data_ptr(), ws.data_ptr(), _stream()))
            queued.append((size, cells, keys, order, buf))
        for size, cells, keys, order, buf in queued:                     # ... then read back: one wait for all of them
            bound = min(int(cells), n)                                   # occupied voxels <= cells of the box
            if bound <= (1 << 22):
                host = buf[: bound + 2].cpu().numpy()
                nvox = int(host[0])
                seg = host[1: nvox + 2]
            else:
                nvox = int(buf[0].item())
                seg = buf[1: nvox + 2].cpu().numpy()
            out.append((float(size), order, seg, keys))
        return out

    def _thin(self, feat: Tensor, order: Tensor, seg: np.ndarray, big: np.ndarray, refl_min: float, weighted: bool,
              voxel_ids: Tensor, grid_ordinal: int) -> List[Tensor]:
        """`maxpoints` members of every oversized voxel (:116-120)."""
        members = torch.cat([order[seg[v]: seg[v + 1]] for v in big])
        sizes = (seg[big + 1] - seg[big]).astype(np.int64)
        picks = thin_tiles(feat, 3, members, sizes, None, refl_min, weighted, voxel_ids, self.maxpoints,
                           self.seed, grid_ordinal)
        return list(picks.view(len(big), self.maxpoints))

    # ---- src/preprocessing.py:79-127
    def write_voxels(self) -> TileStore:
        pos = self.pos
        if hasattr(pos, "columns"):                                  # pandas DataFrame, as in the reference
            has_nz = "n_z" in pos.columns
            nz_col = list(pos.columns).index("n_z") if has_nz else -1          # by name, not by position
            arr = np.ascontiguousarray(pos.values, dtype=np.float32)
        else:
            has_nz = False
            arr = pos
        cloud = torch.as_tensor(arr, dtype=torch.float32)
        if not cloud.is_cuda:
            cloud = cloud.cuda(non_blocking=True)
        self.cloud = cloud = cloud.contiguous()
        if cloud.dim() != 2 or cloud.size(1) < 4:
            raise _lib.P2WError("Voxelise: the cloud needs x, y, z, reflectance columns")
        self._stats = None
        self.finite_rows = None
        _, _, bad, nan_refl = self._cloud_stats()
        if nan_refl > 0:
            raise ValueError("Input reflectance tensor contains NaN values.")             # :20-21
        n_all = cloud.size(0)
        if bad[:3].sum() > 0:          # rows with a non-finite coordinate reach no tile (:123)
            self.finite_rows = torch.nonzero(torch.isfinite(cloud[:, :3]).all(dim=1)).view(-1)
            self.cloud = cloud = cloud[self.finite_rows].contiguous()
            self._stats = None
        n_z = cloud[:, nz_col].contiguous() if has_nz else self.gpu_ground()
        self.n_z = n_z
        if self.finite_rows is not None:
            self.n_z = torch.full((n_all,), float("nan"), device=cloud.device).index_copy_(0, self.finite_rows, n_z)
        ext = self._cloud_stats()[1]
        reflectance_not_zero = bool(ext[0, 3] != 0 or ext[1, 3] != 0)
        refl = self.quantile_normalize_reflectance() if reflectance_not_zero else None
        n = cloud.size(0)
        dev = cloud.device
        feat = torch.empty((n, 5), device=dev, dtype=torch.float32)
        _lib.check(_lib.lib().p2w_assemble5(cloud.data_ptr(), cloud.stride(0), None if refl is None else refl.data_ptr(),
                                            n_z.data_ptr(), n, feat.data_ptr(), _stream()))
        # The reference voxelises EVERY column it holds (src/preprocessing.py:58 on self.pos): x, y, z, the normalised
        # reflectance, whatever further scalar fields the file carried (predict.py:36-53 keeps them) and n_z -- in the
        # frame's column order, n_z appended last when gpu_ground computed it (:52).  The tiles' rows are packed from
        # `feat` (x, y, z, reflectance, n_z) either way.
        n_extra = cloud.size(1) - 4 - (1 if has_nz else 0)
        if n_extra > 0 or (has_nz and nz_col != cloud.size(1) - 1):
            cols = [feat[:, :4]] + [cloud[:, c: c + 1] for c in range(4, cloud.size(1))]
            if not has_nz:
                cols.append(feat[:, 4:5])
            grid_feat = torch.cat(cols, dim=1).contiguous()
        else:
            grid_feat = feat
        pieces: List[Tensor] = []
        sizes: List[np.ndarray] = []
        grids: List[np.ndarray] = []
        for gi, (size, order, seg, keys) in enumerate(self.grid(grid_feat)):
            counts = np.diff(seg)
            keep = np.nonzero(counts >= self.minpoints)[0]
            if not len(keep):
                continue
            big = keep[counts[keep] > self.maxpoints]
            # all kept voxels of this grid in one gather: member m of tile t sits at order[seg[v_t] + m]
            tile_sizes = np.minimum(counts[keep], self.maxpoints).astype(np.int64)
            off = np.concatenate([[0], np.cumsum(tile_sizes)]).astype(np.int64)
            plan = _lib.to_device(np.stack([seg[keep].astype(np.int64) - off[:-1], tile_sizes]), dev)
            src = torch.arange(int(off[-1]), device=dev) + torch.repeat_interleave(plan[0], plan[1], output_size=int(off[-1]))
            members = order[src].to(torch.int64)
            if len(big):                                                # thinned tiles overwrite their placeholder rows
                refl_min = float(self._feat_ext[0, 3])                    # min of the normalised reflectance (:99)
                where = {int(v): i for i, v in enumerate(keep.tolist())}
                vox = keys[torch.as_tensor(seg[big], device=dev)].contiguous()
                for v, picked in zip(big.tolist(), self._thin(feat, order, seg, big, refl_min, reflectance_not_zero,
                                                             vox, gi)):
                    members[off[where[v]]: off[where[v] + 1]] = picked.to(torch.int64)
            pieces.append(members)
            sizes.append(tile_sizes)
            grids.append(np.full(len(keep), size, dtype=np.float32))
        members = torch.cat(pieces) if pieces else torch.empty(0, dtype=torch.int64, device=dev)
        ptr = np.concatenate([[0], np.cumsum(np.concatenate(sizes))]).astype(np.int64) if sizes else np.zeros(1, np.int64)
        grids = np.concatenate(grids) if grids else np.zeros(0, np.float32)
        return TileStore(feat=feat, members=members, ptr=ptr, grid_of_tile=grids, finite_rows=self.finite_rows)


def preprocess(args) -> None:
    """src/preprocessing.py:129-131: tiles into args.tiles (instead of files under args.vxfile),
    n_z appended to args.pc when it is a DataFrame."""
    vox = Voxelise(args.pc, vxpath=getattr(args, "vxfile", None), minpoints=args.min_pts, maxpoints=args.max_pts,
                   pointspacing=getattr(args, "resolution", 0.01), gridsize=args.grid_size)
    args.tiles = vox.write_voxels()
    if hasattr(args.pc, "columns"):
        args.pc["n_z"] = vox.n_z.cpu().numpy()
