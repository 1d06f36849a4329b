"""Inference driver: the B200-native counterpart of src/predicter.py (SemanticSegmentation).

Per batch of `batch_size` tiles (src/predicter.py:193-215): K7 packs the tiles straight from the
on-device TileStore (mean shift, scale factor, concatenation -- :78-94 and PyG's collate), the
network runs on the libp2w ops, K8 turns logits into (x, y, z, pred, prob) rows (:199-214).
No `.pt` files, no per-tile torch.load, no Python loop over tiles inside a batch.

Pinned batch composition (the reference's BalancedBatchSampler shuffles with an unseeded RNG and
silently drops leftover tiles): tiles in TileStore order, consecutive
groups of `batch_size`, the last group may be short, nothing is dropped.  A tile's sub-sampled
representatives depend on its batch-mates (C.3), so the CPU oracle uses the same rule.

Multi-GPU: batches are independent; `shard_batches` deals whole batches to
ranks longest-first, there is no collective on the inference path, rank 0 gathers the rows.
"""
from __future__ import annotations

import os
from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from . import model as M
from . import ops
from .preprocessing import TileStore

__all__ = ["plan_batches", "plan_launches", "shard_batches", "classify_tiles", "PointCloudClassifier", "gather_rows", "SemanticSegmentation"]


def plan_batches(num_tiles: int, batch_size: int) -> List[Tuple[int, int]]:
    """[(first_tile, last_tile_exclusive)] -- consecutive groups of batch_size tiles."""
    return [(t, min(t + batch_size, num_tiles)) for t in range(0, num_tiles, batch_size)]


def shard_batches(batches: Sequence[Tuple[int, int]], ptr: np.ndarray, world_size: int, rank: int):
    """Greedy longest-first assignment of whole batches to ranks by point count; returns the
    indices (into `batches`) owned by `rank`, in ascending order.  Deterministic on every rank."""
    load = np.zeros(world_size, dtype=np.int64)
    owner = np.empty(len(batches), dtype=np.int64)
    pts = np.array([ptr[b] - ptr[a] for a, b in batches], dtype=np.int64)
    for i in np.argsort(-pts, kind="stable"):
        r = int(np.argmin(load))
        owner[i] = r
        load[r] += pts[i]
    return [i for i in range(len(batches)) if owner[i] == rank]


def plan_launches(batches: Sequence[Tuple[int, int]], ptr: np.ndarray, batch_ids: Sequence[int],
                  max_points: int) -> List[List[int]]:
    """Groups CONSECUTIVE reference batches into super-batches of at most `max_points` points (always
    at least one batch): one set of kernel launches serves a whole group, each batch keeping its own
    voxel-grid origin (engine.InferenceEngine)."""
    out: List[List[int]] = []
    cur: List[int] = []
    pts = 0
    for bi in batch_ids:
        a, b = batches[bi]
        n = int(ptr[b] - ptr[a])
        if cur and (bi != cur[-1] + 1 or pts + n > max_points):
            out.append(cur)
            cur, pts = [], 0
        cur.append(bi)
        pts += n
    if cur:
        out.append(cur)
    return out


@torch.no_grad()
def classify_tiles(net: torch.nn.Module, tiles: TileStore, batch_size: int = 8, is_wood: float = 0.5,
                   batch_ids: Optional[Iterable[int]] = None, want_rows: bool = False,
                   max_points_per_launch: int = 1 << 21, want_xyz: bool = False):
    """Runs the network over the tiles.  Returns (prob float32 [M'], pred uint8 [M'], rows float64
    [M',5] or None, spans) on the device, in batch order; M' covers the selected batches.  With
    want_xyz the third item is instead the un-shifted FP32 coordinates [M',3] the spatial vote searches.
    Batches of `batch_size` tiles are the reference's unit (their composition fixes the voxel-grid
    origin); up to `max_points_per_launch` points of consecutive batches share one launch set."""
    dev = tiles.feat.device
    batches = plan_batches(tiles.num_tiles, batch_size)
    batch_ids = list(range(len(batches)) if batch_ids is None else batch_ids)
    ptr_dev = _lib.to_device(tiles.ptr, dev, np.int64)
    probs, preds, rows, spans = [], [], [], []
    for group in plan_launches(batches, tiles.ptr, batch_ids, max_points_per_launch):
        t0, t1 = batches[group[0]][0], batches[group[-1]][1]
        lo, hi = int(tiles.ptr[t0]), int(tiles.ptr[t1])
        bptr = ptr_dev[t0: t1 + 1] - lo
        gptr = _lib.to_device([batches[b][0] - t0 for b in group] + [t1 - t0], dev, np.int64)
        pos, refl, batch, shift, sf = ops.pack_tiles(tiles.feat, tiles.members[lo:hi], bptr)
        data = M.make_data(pos, refl, batch, sf, local_shift=shift.reshape(-1), ptr=bptr,
                           group_ptr=gptr if len(group) > 1 else None)
        logits = net(data)
        out = ops.writeback(logits.float().reshape(-1), pos, bptr, shift, is_wood, want_rows=want_rows,
                            want_xyz=want_xyz and not want_rows)
        probs.append(out[0])
        preds.append(out[1])
        if want_rows or want_xyz:
            rows.append(out[2])
        spans.append((lo, hi))
    cat = (lambda xs: torch.cat(xs) if xs else torch.empty(0, device=dev))
    return cat(probs), cat(preds), (cat(rows) if (want_rows or want_xyz) else None), spans


def gather_rows(rows: torch.Tensor, batch_ids: Sequence[int], dst: int = 0):
    """The one exchange of the multi-GPU inference path: every rank sends the rows it classified
    (any [m, c] tensor, batch order) and the batch indices they belong to; `dst` gets the pieces of
    all ranks as a list of (batch_ids, rows) and re-orders them by batch.  Works on NCCL (device
    tensors) and gloo (host tensors); returns None on the other ranks."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    meta = [None] * world
    dist.all_gather_object(meta, (list(batch_ids), int(rows.size(0))))
    width = rows.size(1)
    longest = max(m[1] for m in meta)
    padded = rows.new_zeros((longest, width))
    padded[: rows.size(0)] = rows
    bucket = [torch.empty_like(padded) for _ in range(world)] if rank == dst else None
    dist.gather(padded, bucket, dst=dst)
    if rank != dst:
        return None
    return [(meta[r][0], bucket[r][: meta[r][1]]) for r in range(world)]


class PointCloudClassifier:
    """src/predicter.py:107-142: the spatial vote that turns the per-tile classifications (every point is
    classified once per tile that holds it) into one label / pwood per ORIGINAL point."""

    def __init__(self, is_wood, any_wood):
        self.is_wood = is_wood
        self.any_wood = any_wood

    def collect_predictions(self, classification, original, cell_size: float = 0.05):
        """classification: float64 [M,5] device rows (x, y, z, pred, prob) or a tuple
        (xyz float32 [M,3], pred uint8 [M], prob float32 [M]); original: [N, >=3] device array or a pandas
        DataFrame (x, y, z first).  Returns (label uint8 [N], pwood float64 [N]) on the device and, for a
        DataFrame, also stores them in its 'label' / 'pwood' columns as the reference does (:141)."""
        if isinstance(classification, tuple):
            xyz, pred, prob = classification
        else:
            xyz = classification[:, :3].to(torch.float32).contiguous()
            pred = classification[:, 3].to(torch.uint8)
            prob = classification[:, 4].to(torch.float32)
        frame = original if hasattr(original, "columns") else None
        if frame is not None:
            frame = frame.drop(columns=[c for c in frame.columns if c in ("label", "pwood", "pleaf")])
            org = torch.as_tensor(np.ascontiguousarray(frame.values[:, :3], dtype=np.float32)).to(xyz.device)
        else:
            org = original[:, :3].to(torch.float32).contiguous()
        k = 32 if self.any_wood != 1 else 64                            # :137
        label, pwood = ops.spatial_vote(xyz, prob.contiguous(), pred.contiguous(), org, k, float(self.any_wood),
                                        cell_size)
        if frame is not None:
            frame.loc[:, "label"] = label.cpu().numpy().astype(np.float64)
            frame.loc[:, "pwood"] = pwood.cpu().numpy()
            return frame
        return label, pwood


def SemanticSegmentation(args):
    """src/predicter.py:148-236 without the file output (:233-234, src/io.py).  Expects args.tiles (from
    preprocessing.preprocess), args.pc (the original cloud: DataFrame or [N, >=3] array) and the
    reference's flags (batch_size, is_wood, any_wood, model, wdir).  Sets args.classified_pc (float64
    [M,5] numpy: x, y, z, pred, prob, :217) and the voted per-point result: args.pc gains 'label' /
    'pwood' columns when it is a DataFrame, else args.label / args.pwood hold device tensors."""
    device = torch.device("cuda")
    net = getattr(args, "net", None)
    if net is None:
        net = M.Net(num_classes=1).to(device)
        path = os.path.join(getattr(args, "wdir", "."), "model", getattr(args, "model", "model.pth"))
        try:
            M.load_model(path, net, device)
        except KeyError:
            raise Exception(f"No model loaded at {path}")
    net.eval()
    _, _, rows, _ = classify_tiles(net, args.tiles, args.batch_size, args.is_wood, want_rows=True)
    args.classified_pc = rows.cpu().numpy()
    classifier = PointCloudClassifier(args.is_wood, any_wood=getattr(args, "any_wood", 1))
    pc = getattr(args, "pc", None)
    if pc is not None:
        finite = getattr(args.tiles, "finite_rows", None)
        cloud = torch.as_tensor(np.ascontiguousarray(pc.values[:, :3], dtype=np.float32) if hasattr(pc, "columns") else pc)
        cloud = cloud.to(device)
        if finite is None:
            label, pwood = classifier.collect_predictions(rows, cloud)
        else:             # rows with a non-finite coordinate were never tiled: label 0, pwood NaN
            lab_f, pw_f = classifier.collect_predictions(rows, cloud[finite])
            label = torch.zeros(cloud.size(0), device=device, dtype=torch.uint8).index_copy_(0, finite, lab_f)
            pwood = torch.full((cloud.size(0),), float("nan"), device=device, dtype=torch.float64).index_copy_(0, finite, pw_f)
        if hasattr(pc, "columns"):
            pc = pc.drop(columns=[c for c in pc.columns if c in ("label", "pwood", "pleaf")])
            pc.loc[:, "label"] = label.cpu().numpy().astype(np.float64)
            pc.loc[:, "pwood"] = pwood.cpu().numpy()
            args.pc = pc
        else:
            args.label, args.pwood = label, pwood
    return args
