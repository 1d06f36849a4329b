"""One plot sharded over several GPUs (BASELINE.json configs[3]).

One process per GPU (`torch.distributed`, NCCL).  Rank r holds a CONTIGUOUS CHUNK of the plot's rows
(rows o_r .. o_r + n_r of the file, rank order = row order) and nothing is replicated except
metadata-sized tables:

1. statistics: column min / max, 5 m ground cells (src/preprocessing.py:37-53) -- all-reduce MIN over a few
   KB; reflectance ranks (:18-30) need the global order: a distributed sort by key ranges (8 B / point out,
   4 B / point back, all-to-all), stable in the point index like one sort of the whole column;
2. tiling (:55-64): every rank sorts ITS rows by 5-D voxel id and lists its occupied voxels with their
   counts; the lists (16 B / occupied voxel / rank) are all-gathered and merged into the plot's tile table --
   voxels with >= min_pts members, 2 m list then 4 m list by ascending id, batches of `batch_size`
   consecutive tiles (the composition one GPU uses, src/predicter.py:177-180), whole batches dealt round
   to the ranks (batch b to rank b mod world: the same mix of small and large tiles everywhere);
3. ONE all-to-all moves every tile member to the owner of its tile (24 B / tile point: x, y, z,
   reflectance, point index, tile).  Stable on both sides, so a tile's rows arrive in ascending point
   index -- the order one GPU sees;
4. each rank classifies its batches (predicter.classify_tiles; no collective, tiles are independent);
5. spatial vote (src/predicter.py:107-142), sharded by x-slabs with equal query counts: queries (12 B) and
   classified rows (20 B: x, y, z, prob, row id) go to their slab's rank by all-to-all, rows within `halo` of a
   slab edge also to the neighbour.  Rows carry their global row id and are sorted by it on arrival, so
   distance ties break as on one GPU.  Every query's k-th neighbour distance is checked against its distance to the edge of the halo;
   if a single query fails the bound the halo grows and the vote is redone, so the result is EXACT;
6. (label, pwood) return to the rank that holds the row (9 B / point, all-to-all).

The result equals the single-GPU one bit for bit (tests/test_gpu_distributed.py).  Bytes per collective
are recorded in `ShardedPlot.traffic`.  The host-side plan (tables, routing) is plain torch and runs on
gloo/CPU tensors too (tests/test_dist_gloo.py); the per-point work goes through `_Kernels` = libp2w.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
from torch import Tensor

from . import _lib, ops
from .predicter import classify_tiles, plan_batches
from .preprocessing import SUBSAMPLE_SEED, TileStore, thin_tiles

__all__ = ["shard_bounds", "shard_contiguous", "all_gather_rows", "merge_voxel_tables", "slab_bounds", "halo_entries",
           "Comm", "ShardedPlot", "classify_plot", "GRID_FLAG_SHIFT"]

GRID_FLAG_SHIFT = 58          # voxel ids of the g-th grid size travel as id | g << 58
DEFAULT_HALO = 1.0            # metres; the k-th neighbour of a TLS point is centimetres away (0.5 m needed a second
                              # round for a handful of isolated points on the 16 M and 100 M-point plots)
HIST_BINS = 4096
REFL_SAMPLE = 4096            # keys per rank in the splitter sample of the distributed reflectance ranking


# ------------------------------------------------------------------------------------------ host-side plans
def shard_bounds(batches: Sequence[Tuple[int, int]], ptr: np.ndarray, world_size: int) -> List[int]:
    """[world+1] batch indices: rank r owns batches bounds[r] .. bounds[r+1], contiguous ranges whose point
    counts are as even as a prefix split allows (boundary i sits where the cumulative point count crosses
    i/world of the total).  Deterministic, identical on every rank."""
    pts = np.array([ptr[b] - ptr[a] for a, b in batches], dtype=np.int64)
    cum = np.concatenate([[0], np.cumsum(pts)])
    total = int(cum[-1])
    bounds = [int(np.searchsorted(cum, total * r / world_size, side="left")) for r in range(world_size)] + [len(batches)]
    bounds[0] = 0
    for r in range(1, world_size + 1):                       # monotone, in range
        bounds[r] = min(max(bounds[r], bounds[r - 1]), len(batches))
    return bounds


def shard_contiguous(batches: Sequence[Tuple[int, int]], ptr: np.ndarray, world_size: int, rank: int) -> List[int]:
    """Batch indices of `rank` (see shard_bounds); every batch belongs to exactly one rank."""
    b = shard_bounds(batches, ptr, world_size)
    return list(range(b[rank], b[rank + 1]))


def merge_voxel_tables(table: Tensor, min_pts: int):
    """table int64 [U, 2] = (flagged voxel id, member count), the concatenation of every rank's occupied
    voxels.  Returns (gid [G] ascending distinct ids, total [G] members over all ranks, kept [G] bool,
    ordinal [G] = index of the voxel in the plot's tile list (valid where kept)).  Ascending flagged id IS
    the reference's tile order: first grid size first, then ascending voxel id (src/preprocessing.py:57-63)."""
    gid, inv = torch.unique(table[:, 0], sorted=True, return_inverse=True)
    total = torch.zeros(gid.numel(), dtype=torch.int64, device=table.device).index_add_(0, inv, table[:, 1])
    kept = total >= min_pts
    ordinal = torch.cumsum(kept.to(torch.int64), 0) - 1
    return gid, total, kept, ordinal


def slab_bounds(hist: np.ndarray, lo: float, hi: float, world_size: int) -> np.ndarray:
    """[world-1] float32 x-coordinates that cut the plot into slabs of (nearly) equal point counts; `hist` is
    the plot-wide histogram of x over HIST_BINS equal bins of [lo, hi]."""
    cum = np.cumsum(hist.astype(np.float64))
    width = (float(hi) - float(lo)) / len(hist)
    cuts = []
    for r in range(1, world_size):
        b = int(np.searchsorted(cum, cum[-1] * r / world_size, side="left"))
        cuts.append(np.float32(float(lo) + (b + 1) * width))
    return np.maximum.accumulate(np.asarray(cuts, dtype=np.float32)) if cuts else np.zeros(0, np.float32)


def halo_entries(x: Tensor, bounds: Tensor, halo: float, world_size: int, slots: int):
    """Destinations of classified rows: row i goes to every slab that [x_i - halo, x_i + halo] touches, slab s
    covering bounds[s-1] <= x < bounds[s].  Returns (keys int64 [M * slots] row-major: the destination of
    (row, slot) or `world_size` for an unused slot, span [M] = number of slabs the row touches)."""
    s0 = torch.searchsorted(bounds, (x - halo).contiguous(), right=True)
    s1 = torch.searchsorted(bounds, (x + halo).contiguous(), right=True)
    d = s0[:, None] + torch.arange(slots, device=x.device)[None, :]
    keys = torch.where(d <= s1[:, None], d, torch.full_like(d, world_size))
    return keys.reshape(-1).contiguous(), (s1 - s0 + 1)


# ------------------------------------------------------------------------------------------ collectives
class Comm:
    """The collectives of the sharded path with their byte counts; world size 1 short-circuits."""

    def __init__(self, group=None, rank: Optional[int] = None, world_size: Optional[int] = None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        if world_size is None:
            on = dist.is_available() and dist.is_initialized()
            world_size = dist.get_world_size(group) if on else 1
            rank = dist.get_rank(group) if on else 0
        self.world, self.rank = world_size, rank
        self.traffic: Dict[str, int] = {}
        # gloo moves host memory: device tensors are staged through the host (two ranks sharing one GPU in the tests)
        self.stage = world_size > 1 and dist.is_initialized() and dist.get_backend(group) == "gloo"

    def _in(self, t: Tensor) -> Tensor:
        return t.cpu() if (self.stage and t.is_cuda) else t

    def _note(self, what: str, nbytes: int):
        self.traffic[what] = self.traffic.get(what, 0) + int(nbytes)

    def all_reduce(self, t: Tensor, op: str, what: str) -> Tensor:
        if self.world > 1:
            h = self._in(t)
            self.dist.all_reduce(h, op=getattr(self.dist.ReduceOp, op), group=self.group)
            if h is not t:
                t.copy_(h)
            self._note(what, t.numel() * t.element_size())
        return t

    def all_gather_equal(self, t: Tensor, what: str) -> Tensor:
        """[world, *t.shape] of same-shaped tensors."""
        if self.world == 1:
            return t.unsqueeze(0)
        h = self._in(t.contiguous())
        out = h.new_empty((self.world,) + tuple(t.shape))
        self.dist.all_gather_into_tensor(out.view(-1), h.view(-1), group=self.group)
        self._note(what, out.numel() * out.element_size())
        return out.to(t.device)

    def all_gather_v(self, t: Tensor, counts: Sequence[int], what: str) -> Tensor:
        """Concatenation over ranks (rank order) of tensors whose first dimensions are `counts` (known on the host)."""
        if self.world == 1:
            return t
        longest = max(counts)
        if min(counts) == longest:
            return self.all_gather_equal(t, what).view((-1,) + tuple(t.shape[1:]))
        padded = t.new_zeros((longest,) + tuple(t.shape[1:]))
        padded[: t.size(0)] = t
        out = self.all_gather_equal(padded, what)
        return torch.cat([out[r, : counts[r]] for r in range(self.world)])

    def all_to_all(self, rows: Tensor, send: Sequence[int], recv: Sequence[int], what: str) -> Tensor:
        """rows (first dimension split by `send`, destination-major) -> what the other ranks sent here, source-major."""
        if self.world == 1:
            return rows
        h = self._in(rows.contiguous())
        out = h.new_empty((int(sum(recv)),) + tuple(rows.shape[1:]))
        self.dist.all_to_all_single(out, h, [int(c) for c in recv], [int(c) for c in send], group=self.group)
        self._note(what, rows.numel() * rows.element_size())
        return out.to(rows.device)


def all_gather_rows(rows: Tensor) -> Tensor:
    """Concatenation over ranks (rank order) of tensors that differ in their first dimension."""
    comm = Comm()
    if comm.world == 1:
        return rows
    count = torch.tensor([rows.size(0)], device=rows.device, dtype=torch.int64)
    counts = comm.all_gather_equal(count, "counts").view(-1).tolist()
    return comm.all_gather_v(rows, counts, "rows")


# ------------------------------------------------------------------------------------------ per-point work
def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


class _Kernels:
    """The per-point device work of the sharded path: thin wrappers over libp2w (no CPU path; the gloo tests
    substitute an oracle-backed stand-in to exercise the exchange plan on CPU tensors)."""

    def colminmax(self, a: Tensor):
        return ops._colminmax(a)

    def ground_min(self, cloud: Tensor, mn_xy: Tensor, nbx: int, nby: int) -> Tensor:
        cell_min = torch.empty((nbx + 1) * (nby + 1), device=cloud.device, dtype=torch.float32)
        _lib.check(_lib.lib().p2w_ground_min(cloud.data_ptr(), cloud.stride(0), cloud.size(0), mn_xy.data_ptr(), 5.0, nbx,
                                             nby, cell_min.data_ptr(), _stream()))
        return cell_min

    def ground_apply(self, cloud: Tensor, mn_xy: Tensor, nbx: int, nby: int, cell_min: Tensor) -> Tensor:
        n_z = torch.empty(cloud.size(0), device=cloud.device, dtype=torch.float32)
        _lib.check(_lib.lib().p2w_ground_apply(cloud.data_ptr(), cloud.stride(0), cloud.size(0), mn_xy.data_ptr(), 5.0, nbx,
                                               nby, cell_min.data_ptr(), n_z.data_ptr(), _stream()))
        return n_z

    def reflectance_keys(self, cloud: Tensor) -> Tensor:
        """int64 [n]: order-preserving 32-bit keys of the reflectance column (src/preprocessing.py:22's sort key)."""
        keys = torch.empty(cloud.size(0), device=cloud.device, dtype=torch.int64)
        _lib.check(_lib.lib().p2w_reflectance_keys(cloud.data_ptr(), cloud.stride(0), 3, cloud.size(0), keys.data_ptr(), _stream()))
        return keys

    def reflectance_values(self, order: Tensor, rank0: int, n_total: int):
        """(v [n] with v[order[p]] = the normal score of global rank rank0 + p, mnmx [2] = min / max of v) (:24-26)."""
        n = order.numel()
        v = torch.empty(n, device=order.device, dtype=torch.float32)
        mnmx = torch.empty(2, device=order.device, dtype=torch.float32)
        _lib.check(_lib.lib().p2w_reflectance_values(order.data_ptr(), n, rank0, n_total, v.data_ptr(), mnmx.data_ptr(), _stream()))
        return v, mnmx

    def reflectance_scale(self, v: Tensor, mnmx: Tensor) -> Tensor:
        """2 (v - min) / (max - min) - 1 (:27-29)."""
        out = torch.empty_like(v)
        _lib.check(_lib.lib().p2w_reflectance_scale(v.data_ptr(), v.numel(), mnmx.data_ptr(), out.data_ptr(), _stream()))
        return out

    def assemble5(self, cloud: Tensor, refl: Optional[Tensor], n_z: Tensor) -> Tensor:
        feat = torch.empty((cloud.size(0), 5), device=cloud.device, dtype=torch.float32)
        _lib.check(_lib.lib().p2w_assemble5(cloud.data_ptr(), cloud.stride(0), None if refl is None else refl.data_ptr(),
                                            n_z.data_ptr(), cloud.size(0), feat.data_ptr(), _stream()))
        return feat

    def grid_ids(self, feat: Tensor, size: float, start: Tensor, end: Tensor) -> Tensor:
        sz = torch.full((feat.size(1),), float(size), device=feat.device, dtype=torch.float32)
        return ops.grid_cluster(feat, sz, start, end)

    def stable_order(self, keys: Tensor, bits: int):
        """(sorted keys, int32 positions) of non-negative int64 keys, stable."""
        return ops.sort_pairs(keys, bits)

    def segments(self, sorted_keys: Tensor, order: Tensor):
        """(starts int64 [n+1] device (first n_unique+1 entries valid), n_unique int64 [1] device)."""
        n = sorted_keys.numel()
        L = _lib.lib()
        buf = torch.empty(n + 2, device=sorted_keys.device, dtype=torch.int64)
        ws = torch.empty(max(int(L.p2w_unique_ws_bytes(n)), 8), device=sorted_keys.device, dtype=torch.uint8)
        _lib.check(L.p2w_unique_last(sorted_keys.data_ptr(), order.data_ptr(), n, None, None, buf[1:].data_ptr(),
                                     buf.data_ptr(), ws.data_ptr(), _stream()))
        return buf[1:], buf[:1]

    def thin(self, *args, **kw) -> Tensor:
        return thin_tiles(*args, **kw)

    def vote(self, rows_xyz: Tensor, prob: Tensor, pred: Tensor, queries: Tensor, k: int, any_wood: float):
        """(label, pwood, nbr [nq, k] int32) -- ops.spatial_vote with the neighbour table kept for the halo check."""
        return ops.spatial_vote(rows_xyz, prob, pred, queries, k, any_wood, return_table=True)


def _to_dev(a, dev, dtype=np.int64) -> Tensor:
    """Small host array to `dev` without blocking the host on CUDA (pinned staging); plain tensor on CPU."""
    if dev.type == "cuda":
        return _lib.to_device(a, dev, dtype)
    return torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=dtype)))


def _sorted_counts(sorted_keys: Tensor, nbins: int) -> Tensor:
    """int64 [nbins]: how many of the SORTED non-negative keys equal 0 .. nbins-1 (nbins is the world size: a
    histogram kernel would serialise a million atomics on two to eight counters; the stable sort that groups the
    rows by destination has the boundaries for free)."""
    edges = torch.searchsorted(sorted_keys, torch.arange(nbins + 1, device=sorted_keys.device, dtype=sorted_keys.dtype))
    return edges[1:] - edges[:-1]


def _bits(n: int) -> int:
    return max(1, int(max(n, 1) - 1).bit_length())


class ShardedPlot:
    """Tiling, member exchange and spatial vote of one plot whose rows are chunked over the ranks."""

    def __init__(self, chunk: Tensor, comm: Optional[Comm] = None, min_pts: int = 128, max_pts: int = 16384,
                 grid_size=(2.0, 4.0), batch_size: int = 8, seed: int = SUBSAMPLE_SEED, kernels: Optional[_Kernels] = None):
        self.chunk = chunk.contiguous()
        if self.chunk.dim() != 2 or self.chunk.size(1) != 4 or self.chunk.dtype != torch.float32:
            raise _lib.P2WError("ShardedPlot: the chunk must be float32 [n, 4] (x, y, z, reflectance); further scalar "
                                "columns would take part in the 5-D voxel grid (Voxelise handles them on one GPU)")
        self.comm = comm or Comm()
        self.min_pts, self.max_pts, self.grid_size, self.batch_size, self.seed = min_pts, max_pts, list(grid_size), batch_size, seed
        self.K = kernels or _Kernels()
        self.traffic = self.comm.traffic
        self.n_z: Optional[Tensor] = None
        self.timing = False           # True: CUDA events at the phase boundaries (phase_ms() after a synchronise)
        self._marks: List[Tuple[str, object]] = []

    def _mark(self, name: str) -> None:
        if self.timing and self.chunk.is_cuda:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self._marks.append((name, ev))

    def phase_ms(self) -> Dict[str, float]:
        """Device time between consecutive marks (name = the phase that ENDS at the mark)."""
        torch.cuda.synchronize()
        out: Dict[str, float] = {}
        for (_, a), (name, b) in zip(self._marks[:-1], self._marks[1:]):
            out[name] = out.get(name, 0.0) + a.elapsed_time(b)
        return out

    # ---------------------------------------------------------------- reflectance ranks (src/preprocessing.py:18-30)
    def _rank_reflectance(self, keys: Tensor) -> Tensor:
        """quantile_normalize_reflectance over the whole plot without gathering it: rank r sorts the keys of ONE key
        range.  The ranges come from a sample (REFL_SAMPLE evenly spaced keys per rank, all-gathered and sorted on every
        rank: the same splitters everywhere, ranges of nearly equal counts; any splitters give the exact result).  Keys
        travel to their range's rank in point order and are sorted stably there, so equal reflectances rank by point
        index as one stable sort of the whole column would; a value's global rank is its range's offset + its sorted
        position, the normal scores go back the way the keys came, and min / max are reduced over the ranks before
        the affine map."""
        K, comm = self.K, self.comm
        W, r = comm.world, comm.rank
        dev, n = keys.device, keys.numel()
        if W == 1:
            _, order = K.stable_order(keys, 32)
            v, mnmx = K.reflectance_values(order, 0, n)
            return K.reflectance_scale(v, mnmx)
        if n:
            sample = keys[(torch.arange(REFL_SAMPLE, device=dev) * n) // REFL_SAMPLE]
        else:
            sample = torch.full((REFL_SAMPLE,), (1 << 32) - 1, device=dev, dtype=torch.int64)
        pool, _ = K.stable_order(comm.all_gather_equal(sample, "all-gather: reflectance sample").view(-1).contiguous(), 32)
        cuts = pool[(torch.arange(1, W, device=dev) * pool.numel()) // W].contiguous()      # first key of ranges 1 .. W-1
        dest = torch.searchsorted(cuts, keys, right=True)
        if n:
            sorted_dest, by_dest = K.stable_order(dest.contiguous(), _bits(W))
            by_dest = by_dest.long()
            send = _sorted_counts(sorted_dest, W)
        else:
            by_dest = torch.empty(0, device=dev, dtype=torch.int64)
            send = torch.zeros(W, device=dev, dtype=torch.int64)
        cm = comm.all_gather_equal(send, "all-gather: exchange sizes").cpu().numpy()                   # sync
        got = comm.all_to_all(keys[by_dest], cm[r].tolist(), cm[:, r].tolist(),
                              "all-to-all: reflectance keys")
        _, order = K.stable_order(got, 32) if got.numel() else (None, torch.empty(0, device=dev, dtype=torch.int32))
        rank0 = int(cm[:, :r].sum())
        v, mnmx = K.reflectance_values(order, rank0, self.total)
        mm = comm.all_reduce(torch.stack([mnmx[0], -mnmx[1]]), "MIN", "all-reduce: reflectance min / max")
        back = comm.all_to_all(v, cm[:, r].tolist(), cm[r].tolist(), "all-to-all: reflectance scores")
        home = torch.empty(n, device=dev, dtype=torch.float32)
        home[by_dest] = back
        return K.reflectance_scale(home, torch.stack([mm[0], -mm[1]]).contiguous())

    # ---------------------------------------------------------------- 1-3: tiling and the member exchange
    def tile(self) -> TileStore:
        K, comm, chunk = self.K, self.comm, self.chunk
        W, r = comm.world, comm.rank
        dev = chunk.device
        n = chunk.size(0)
        self._mark("start")
        # ---- global statistics (src/preprocessing.py:41-42,94; NaN checks :20-21)
        mn, mx = K.colminmax(chunk[:, :4])
        mm = comm.all_reduce(torch.cat([mn, -mx]), "MIN", "all-reduce: column min / max")
        flags = torch.stack([torch.tensor(n, device=dev), torch.isnan(chunk[:, 3]).sum(),
                             (~torch.isfinite(chunk[:, :3])).any(dim=1).sum()]).to(torch.int64)
        flags = comm.all_gather_equal(flags, "all-gather: row counts")
        rkeys = K.reflectance_keys(chunk) if n else torch.empty(0, device=dev, dtype=torch.int64)
        host = torch.cat([mm.double(), flags.view(-1).double()]).cpu().numpy()                        # sync 1
        ext = np.stack([host[:4], -host[4:8]]).astype(np.float32)
        flags_h = host[8: 8 + 3 * W].reshape(W, 3).astype(np.int64)
        if flags_h[:, 1].sum() > 0:
            raise ValueError("Input reflectance tensor contains NaN values.")
        if flags_h[:, 2].sum() > 0:
            raise _lib.P2WError("ShardedPlot: rows with non-finite coordinates must be removed before sharding")
        self.counts = flags_h[:, 0].tolist()
        self.offset = int(sum(self.counts[:r]))
        self.total = int(sum(self.counts))
        if self.total >= 2 ** 31:
            raise _lib.P2WError("ShardedPlot: point indices are 32-bit (fewer than 2^31 rows per plot)")
        self.ext = ext
        gmn = mm[:4].contiguous()
        # histogram of x for the vote's slabs, fetched with the next host copy
        hist = torch.histc(chunk[:, 0], bins=HIST_BINS, min=float(ext[0, 0]), max=float(ext[1, 0])) if n else \
            torch.zeros(HIST_BINS, device=dev)
        hist = comm.all_reduce(hist.to(torch.float64), "SUM", "all-reduce: x histogram")
        self._mark("tile: statistics")
        # ---- height above ground (:37-53)
        lo, hi = ext[0, :2], ext[1, :2] + np.float32(5.0)
        nb = [max(1, int(math.ceil((float(hi[d]) - float(lo[d])) / 5.0))) for d in range(2)]
        cell_min = comm.all_reduce(K.ground_min(chunk, gmn, nb[0], nb[1]), "MIN", "all-reduce: ground cells")
        n_z = K.ground_apply(chunk, gmn, nb[0], nb[1], cell_min)
        self.n_z = n_z
        # ---- reflectance (:18-30): ranks are global, so the column is ranked whole on every rank
        self.weighted = bool(ext[0, 3] != 0 or ext[1, 3] != 0)
        refl = self._rank_reflectance(rkeys) if self.weighted else None
        self._mark("tile: ground + reflectance ranks")
        feat = K.assemble5(chunk, refl, n_z)
        mn5, mx5 = K.colminmax(feat)
        mm5 = comm.all_reduce(torch.cat([mn5, -mx5]), "MIN", "all-reduce: column min / max")
        host = torch.cat([mm5.double(), hist]).cpu().numpy()                                          # sync 2
        ext5 = np.stack([host[:5], -host[5:10]]).astype(np.float32)
        self.bounds = slab_bounds(host[10:], float(ext[0, 0]), float(ext[1, 0]), W)
        start5, end5 = mm5[:5].contiguous(), (-mm5[5:]).contiguous()
        self.refl_min = float(ext5[0, 3])
        # ---- occupied voxels of this rank per grid size (:55-64)
        local = []
        for gi, size in enumerate(self.grid_size):
            cells = 1
            for d in range(5):
                cells *= int(np.float32(ext5[1, d] - ext5[0, d]) / np.float32(size)) + 1
            bits = max(1, int(cells).bit_length())
            if bits > GRID_FLAG_SHIFT:
                raise _lib.P2WError("ShardedPlot: the 5-D voxel grid has more than 2^58 cells")
            if n:
                keys, order = K.stable_order(K.grid_ids(feat, size, start5, end5), bits)
                starts, nuniq = K.segments(keys, order)
            else:
                keys = torch.empty(0, device=dev, dtype=torch.int64)
                order = torch.empty(0, device=dev, dtype=torch.int32)
                starts, nuniq = torch.zeros(1, device=dev, dtype=torch.int64), torch.zeros(1, device=dev, dtype=torch.int64)
            local.append((keys, order, starts, nuniq))
        self._mark("tile: voxel ids, sort, segments")
        nu = torch.cat([l[3] for l in local])
        nu_all = comm.all_gather_equal(nu, "all-gather: voxel list sizes").cpu().numpy()               # sync 3
        tables = []
        for gi, (keys, order, starts, _) in enumerate(local):
            u = int(nu_all[r, gi])
            uid = keys[starts[:u]] | (gi << GRID_FLAG_SHIFT)
            tables.append(torch.stack([uid, starts[1: u + 1] - starts[:u]], dim=1))
        table = comm.all_gather_v(torch.cat(tables), nu_all.sum(axis=1).tolist(), "all-gather: occupied voxels")
        gid, total, kept, ordinal = merge_voxel_tables(table, self.min_pts)
        # this rank's share of every kept voxel, per grid size (before the host copy, so that ONE copy brings the tile
        # table AND the number of members this rank will send)
        mine_u = []
        for gi, (keys, order, starts, _) in enumerate(local):
            u = int(nu_all[r, gi])
            g = torch.searchsorted(gid, keys[starts[:u]] | (gi << GRID_FLAG_SHIFT))
            t_u = torch.where(kept[g], ordinal[g], torch.full_like(g, -1)) if u else g
            cnt_u = starts[1: u + 1] - starts[:u]
            kept_cnt = torch.where(t_u >= 0, cnt_u, torch.zeros_like(cnt_u))
            mine_u.append((t_u, kept_cnt, starts[:u]))
        sent = torch.stack([m[1].sum() for m in mine_u])
        kept_host = torch.cat([total[kept], gid[kept], sent]).cpu().numpy()                           # sync 4
        T = (len(kept_host) - len(mine_u)) // 2
        full, gid_kept, sent_h = kept_host[:T], kept_host[T: 2 * T], kept_host[2 * T:]
        sizes = np.minimum(full, self.max_pts)
        ptr = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
        # ownership: batch b of the plot (tiles b*B .. b*B+B-1) belongs to rank b mod W.  Consecutive batches are
        # neighbours in the voxel order and cost about the same, so dealing them round gives every rank the same mix of
        # 2 m and 4 m tiles (contiguous ranges balanced by points left the ranks 10 % apart in time).
        B = self.batch_size
        nbatch = (T + B - 1) // B
        mine = [np.arange(b * B, min(b * B + B, T), dtype=np.int64) for b in range(r, nbatch, W)]
        local_tiles = np.concatenate(mine) if mine else np.zeros(0, np.int64)       # global tile ids of this rank, ascending
        self.local_tiles, self.tile_ptr, self.num_tiles = local_tiles, ptr, T
        self._mark("tile: merged voxel table + ownership")
        # ---- members of kept voxels, ascending (tile, point index)
        pts, tiles = [], []
        for gi, (keys, order, starts, _) in enumerate(local):
            t_u, kept_cnt, st_u = mine_u[gi]
            m = int(sent_h[gi])
            if m == 0:
                continue
            # member j of kept voxel u sits at order[starts[u] + j]; dropped voxels repeat zero times
            src0 = st_u - (torch.cumsum(kept_cnt, 0) - kept_cnt)
            src = torch.arange(m, device=dev) + torch.repeat_interleave(src0, kept_cnt, output_size=m)
            pts.append(order[src].long())
            tiles.append(torch.repeat_interleave(t_u, kept_cnt, output_size=m))
        pts = torch.cat(pts) if pts else torch.empty(0, device=dev, dtype=torch.int64)
        tiles = torch.cat(tiles) if tiles else torch.empty(0, device=dev, dtype=torch.int64)
        dest = (tiles // B) % W
        if pts.numel():                                # group by destination; stable, so (tile, point index) order survives
            sorted_dest, by_dest = K.stable_order(dest.contiguous(), _bits(W))
            send = _sorted_counts(sorted_dest, W)
            if W > 1:
                pts, tiles = pts[by_dest.long()], tiles[by_dest.long()]
        else:
            send = torch.zeros(W, device=dev, dtype=torch.int64)
        cm = comm.all_gather_equal(send, "all-gather: exchange sizes").cpu().numpy()                   # sync 5
        payload = torch.empty((pts.numel(), 6), device=dev, dtype=torch.int32)
        payload[:, :4] = feat[pts, :4].contiguous().view(torch.int32)
        payload[:, 4] = (pts + self.offset).to(torch.int32)
        payload[:, 5] = tiles.to(torch.int32)
        self._mark("tile: routing plan")
        recv = comm.all_to_all(payload, cm[r].tolist(), cm[:, r].tolist(), "all-to-all: tile members")
        self._mark("tile: member all-to-all")
        # ---- this rank's tiles: stable regroup by tile (sources arrive in rank order = ascending point index)
        full_loc = full[local_tiles].astype(np.int64)
        if int(full_loc.sum()) != recv.size(0):
            raise _lib.P2WError("ShardedPlot: the member exchange delivered a different number of rows than the tile table lists")
        if recv.size(0):
            t_recv = recv[:, 5].to(torch.int64)
            lt = ((t_recv // B) // W) * B + t_recv % B              # local ordinal of a global tile (all but the last batch are full)
            _, perm = K.stable_order(lt.contiguous(), _bits(len(local_tiles)))
            recv = recv[perm.long()]
        feat_loc = recv[:, :4].contiguous().view(torch.float32)
        gidx = recv[:, 4].contiguous()
        ptr_full = np.concatenate([[0], np.cumsum(full_loc)]).astype(np.int64)
        sizes_loc = sizes[local_tiles].astype(np.int64)
        off = np.concatenate([[0], np.cumsum(sizes_loc)]).astype(np.int64)
        gid_loc = gid_kept[local_tiles]
        grid_of_tile = np.asarray(self.grid_size, np.float32)[gid_loc >> GRID_FLAG_SHIFT]
        big = np.nonzero(full_loc > self.max_pts)[0]
        if len(big) == 0:
            members = torch.arange(recv.size(0), device=dev, dtype=torch.int64)
        else:
            plan = _to_dev(np.stack([ptr_full[:-1] - off[:-1], sizes_loc]), dev)
            members = torch.arange(int(off[-1]), device=dev) + torch.repeat_interleave(plan[0], plan[1], output_size=int(off[-1]))
            vox_all = gid_loc & ((1 << GRID_FLAG_SHIFT) - 1)
            gsel = gid_loc >> GRID_FLAG_SHIFT
            for gi in range(len(self.grid_size)):
                bg = big[gsel[big] == gi]
                if not len(bg):
                    continue
                cat = torch.cat([torch.arange(int(ptr_full[v]), int(ptr_full[v + 1]), device=dev, dtype=torch.int32) for v in bg])
                picks = K.thin(feat_loc, 3, cat, full_loc[bg], gidx, self.refl_min, self.weighted,
                               _to_dev(vox_all[bg], dev), self.max_pts, self.seed, gi)
                picks = picks.view(len(bg), self.max_pts).to(torch.int64)
                for j, v in enumerate(bg.tolist()):
                    members[off[v]: off[v + 1]] = picks[j]
        # global row id of every classified row of this rank (its position in the single-GPU, tile-major row order):
        # the vote sorts the rows it receives by it, so equal distances break ties exactly as on one GPU
        plan = _to_dev(np.stack([ptr[local_tiles] - off[:-1], sizes_loc]), dev)
        self.global_rows = (torch.arange(int(off[-1]), device=dev) +
                            torch.repeat_interleave(plan[0], plan[1], output_size=int(off[-1]))).to(torch.int32)
        if int(ptr[-1]) >= 2 ** 31:
            raise _lib.P2WError("ShardedPlot: classified rows are indexed in 32 bits (fewer than 2^31 tile points per plot)")
        self._mark("tile: regroup + thinning")
        return TileStore(feat=feat_loc, members=members, ptr=off, grid_of_tile=grid_of_tile)

    # ---------------------------------------------------------------- 5-6: the spatial vote by x-slabs
    def vote(self, xyz: Tensor, prob: Tensor, is_wood: float = 0.5, any_wood: float = 1, halo: float = DEFAULT_HALO):
        """xyz float32 [M,3], prob float32 [M]: this rank's classified rows in batch order.  Returns
        (label uint8 [n], pwood float64 [n]) for the rows of this rank's chunk."""
        K, comm, chunk = self.K, self.comm, self.chunk
        W, r = comm.world, comm.rank
        dev = chunk.device
        n = chunk.size(0)
        k = 32 if any_wood != 1 else 64                                                               # :137
        self._mark("classify")
        bounds = _to_dev(self.bounds, dev, np.float32)
        q = chunk[:, :3].contiguous()
        if W > 1 and n:
            dq = torch.searchsorted(bounds, q[:, 0].contiguous(), right=True)
            sorted_dq, qorder = K.stable_order(dq.contiguous(), _bits(W))
            qorder = qorder.long()
            qsend = _sorted_counts(sorted_dq, W)
        else:
            qorder = torch.arange(n, device=dev)
            qsend = torch.zeros(W, device=dev, dtype=torch.int64)
            qsend[r] = n
        # 20 B per classified row: x, y, z, prob, global row id
        rows = torch.cat([xyz.reshape(-1, 3), prob.reshape(-1, 1)], dim=1).contiguous().view(torch.int32)
        rows = torch.cat([rows, self.global_rows.view(-1, 1)], dim=1).contiguous()
        row_bits = max(1, int(self.tile_ptr[-1]).bit_length())
        width = float(self.ext[1, 0]) - float(self.ext[0, 0])
        queries = None
        pending = None                 # queries still to be answered (None: all of them)
        label = pwood = None
        self.vote_rounds = 0
        while True:
            self.vote_rounds += 1
            everything = W == 1 or halo >= width
            slots = W if (everything or W <= 3) else 3
            if everything:
                keys = torch.arange(W, device=dev).repeat(rows.size(0))
                span = torch.full((1,), W, device=dev, dtype=torch.int64)
            else:
                keys, span = halo_entries(xyz[:, 0].contiguous(), bounds, halo, W, slots)
            if keys.numel():
                sorted_keys, eorder = K.stable_order(keys, _bits(W + 1))
                rsend = _sorted_counts(sorted_keys, W)
            else:
                eorder = torch.empty(0, device=dev, dtype=torch.int32)
                rsend = torch.zeros(W, device=dev, dtype=torch.int64)
            top = span.max().view(1) if span.numel() else torch.zeros(1, device=dev, dtype=torch.int64)
            cm = comm.all_gather_equal(torch.cat([qsend, rsend, top]), "all-gather: exchange sizes").cpu().numpy()   # sync
            if not everything and int(cm[:, -1].max()) > slots:      # a row touches more slabs than slots: widen
                halo = width
                continue
            qm, rm = cm[:, :W], cm[:, W: 2 * W]
            if queries is None:
                queries = comm.all_to_all(q[qorder], qm[r].tolist(), qm[:, r].tolist(), "all-to-all: vote queries")
            entries = eorder[: int(rm[r].sum())].long()
            got = comm.all_to_all(rows[entries // slots], rm[r].tolist(), rm[:, r].tolist(), "all-to-all: classified rows")
            if W > 1 and got.size(0):                      # back into the single-GPU row order
                _, by_row = K.stable_order(got[:, 4].to(torch.int64).contiguous(), row_bits)
                got = got[by_row.long()]
            rx, rp = got[:, :3].contiguous().view(torch.float32), got[:, 3].contiguous().view(torch.float32)
            self._mark("vote: routing + all-to-alls")
            # later rounds only redo the queries that failed the bound (a handful of isolated points)
            todo = queries if pending is None else queries[pending]
            lab_t, pw_t, nbr = K.vote(rx, rp, (rp >= is_wood).to(torch.uint8), todo, k, float(any_wood))
            if pending is None:
                label, pwood = lab_t, pw_t
            else:
                label[pending] = lab_t
                pwood[pending] = pw_t
            self._mark("vote: search + vote")
            if everything:
                break
            # ---- exactness: the k-th neighbour must be closer than anything this slab was not given
            lo = float(self.bounds[r - 1]) - (halo - 1e-3) if r > 0 else -math.inf
            hi = float(self.bounds[r]) + (halo - 1e-3) if r < W - 1 else math.inf
            need = torch.zeros(2, device=dev, dtype=torch.float64)
            bad = None
            if todo.size(0):
                # the k-th neighbour sits in the first or in the last column (unordered / ordered table); FP32 is enough:
                # the bound carries a millimetre of slack
                ends = nbr[:, [0, k - 1]].long()
                if rx.size(0):
                    far = (todo[:, None, :] - rx[ends.clamp(min=0)]).pow(2).sum(2).sqrt().max(dim=1).values
                else:
                    far = torch.full((todo.size(0),), math.inf, device=dev, dtype=torch.float32)
                far = torch.where((ends < 0).any(dim=1), torch.full_like(far, math.inf), far)
                qx = todo[:, 0]
                margin = torch.minimum(qx - lo, hi - qx)
                bad = far >= margin
                need = torch.stack([bad.sum().double(), torch.where(bad, far - margin, torch.zeros_like(far)).max().double()])
            need = comm.all_reduce(need, "MAX", "all-reduce: halo check").cpu().numpy()                # sync
            self._mark("vote: halo check")
            if need[0] == 0:
                break
            if bad is None:
                pending = torch.empty(0, device=dev, dtype=torch.int64)
            else:
                sel = torch.nonzero(bad).view(-1)
                pending = sel if pending is None else pending[sel]
            halo = width if not np.isfinite(need[1]) else min(width, 1.25 * (halo + float(need[1])) + 0.01)
        self.halo = halo
        back_l = comm.all_to_all(label, qm[:, r].tolist(), qm[r].tolist(), "all-to-all: labels") if W > 1 else label
        back_p = comm.all_to_all(pwood, qm[:, r].tolist(), qm[r].tolist(), "all-to-all: pwood") if W > 1 else pwood
        out_l = torch.empty(n, device=dev, dtype=torch.uint8)
        out_p = torch.empty(n, device=dev, dtype=torch.float64)
        out_l[qorder] = back_l
        out_p[qorder] = back_p
        self._mark("vote: results home")
        return out_l, out_p


@torch.no_grad()
def classify_plot(net: torch.nn.Module, chunk: Tensor, min_pts: int = 128, max_pts: int = 16384,
                  grid_size=(2.0, 4.0), batch_size: int = 8, is_wood: float = 0.5, any_wood: float = 1,
                  max_points_per_launch: int = 1 << 21, halo: float = DEFAULT_HALO, comm: Optional[Comm] = None,
                  return_plot: bool = False, timing: bool = False):
    """chunk [n_r, >=4] (x, y, z, reflectance): this rank's contiguous rows of the plot, on its device
    -> (label uint8 [n_r], pwood float64 [n_r]) for the same rows.  Single process: the whole plot."""
    plot = ShardedPlot(chunk, comm, min_pts, max_pts, grid_size, batch_size)
    plot.timing = timing
    store = plot.tile()
    prob, _, xyz, _ = classify_tiles(net, store, batch_size, is_wood, max_points_per_launch=max_points_per_launch,
                                     want_xyz=True)
    dev = chunk.device
    if xyz is None or xyz.numel() == 0:
        xyz = torch.empty((0, 3), device=dev, dtype=torch.float32)
        prob = torch.empty(0, device=dev, dtype=torch.float32)
    label, pwood = plot.vote(xyz, prob, is_wood, any_wood, halo)
    plot.tile_points = int(store.ptr[-1])
    return (label, pwood, plot) if return_plot else (label, pwood)
