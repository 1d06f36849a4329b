"""CPU restatement of PointsToWood's host pipeline around the network: tiling
(/root/reference/pointstowood/src/preprocessing.py:18-64, 79-127), batch packing
(src/predicter.py:78-94 + PyG collate) and write-back (src/predicter.py:199-214).

TEST INFRASTRUCTURE (checker + timed CPU arm), numpy / torch-CPU only.  PARITY PINNED to reference-executed code
since round 2: tests/golden/tiling.npz and predicter.npz are written by running the reference's OWN
src/preprocessing.py (one textual device='cuda' -> 'cpu' substitution at load time) and src/predicter.py over
oracle/shim (oracle/make_golden_tiling.py, make_golden_predicter.py), and this restatement reproduces them: n_z, the
normalised reflectance, tile membership and order bit for bit; packing to the ulp of the FP32 mean; vote labels and
pwood exactly (tests/test_oracle_tiling_golden.py, test_oracle_predicter_golden.py).  The third-party primitives under
the shim stay unpinned (oracle.py).  Where the reference is random or order-unstable the same
deterministic choices as the CUDA path are pinned:
  * stable sort for reflectance ranks (C.9);
  * > max_pts tiles: the reference's sampling laws on a counter-based hash instead of torch's generator (C.5):
    weighted without replacement (torch.multinomial, :118) as Efraimidis-Spirakis keys -log(u)/w, rows in
    draw order; without reflectance max_pts uniform draws with replacement (torch.randint, :120);
  * consecutive batches of `batch_size` tiles in tile order, nothing dropped (C.4);
  * local_shift = mean accumulated in float64 (the reference's fp32 torch.mean depends on its
    vectorised summation order).
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import oracle as O
from . import ref_model

SUBSAMPLE_SEED = 141190


def ground_normalize(xyz: np.ndarray) -> np.ndarray:
    """gpu_ground (:37-53): 5 m XY cells from bucketize on edges min + 5 i; n_z = z - cell min."""
    x, y, z = (np.ascontiguousarray(xyz[:, d], dtype=np.float32) for d in range(3))
    cells = []
    for v in (x, y):
        lo, hi = v.min(), np.float32(v.max() + np.float32(5.0))
        nb = max(1, int(math.ceil((float(hi) - float(lo)) / 5.0)))
        edges = (lo + np.float32(5.0) * np.arange(nb, dtype=np.float32)).astype(np.float32)
        cells.append((np.searchsorted(edges, v, side="left"), nb))
    cid = cells[0][0].astype(np.int64) * (cells[1][1] + 1) + cells[1][0]
    mins = np.full(int(cid.max()) + 1, np.inf, dtype=np.float32)
    np.minimum.at(mins, cid, z)
    return (z - mins[cid]).astype(np.float32)


def quantile_values(ranks: np.ndarray, n: int) -> np.ndarray:
    """:24-26: rank -> (rank + 1) / (n + 1) clamped -> erfinv(2q - 1) * sqrt(2), all in FP32."""
    q = (np.asarray(ranks).astype(np.float32) + np.float32(1.0)) / np.float32(n + 1)
    q = np.clip(q, np.float32(1e-7), np.float32(1.0) - np.float32(1e-7))
    return (torch.erfinv(torch.from_numpy(np.float32(2.0) * q - np.float32(1.0))) * torch.sqrt(torch.tensor(2.0))).numpy()


def quantile_normalize_reflectance(refl: np.ndarray) -> np.ndarray:
    """:18-30 with a stable sort."""
    refl = np.ascontiguousarray(refl, dtype=np.float32)
    n = len(refl)
    order = np.argsort(refl, kind="stable")
    ranks = np.empty(n, dtype=np.int64)
    ranks[order] = np.arange(n)
    v = quantile_values(ranks, n)
    mn, mx = v.min(), v.max()
    return (np.float32(2.0) * (v - mn) / (mx - mn) - np.float32(1.0)).astype(np.float32)


def _mix32(h: np.ndarray) -> np.ndarray:
    h = h.astype(np.uint32)
    h ^= h >> np.uint32(16)
    h = (h * np.uint32(0x85EBCA6B)).astype(np.uint32)
    h ^= h >> np.uint32(13)
    h = (h * np.uint32(0xC2B2AE35)).astype(np.uint32)
    h ^= h >> np.uint32(16)
    return h


def sampling_keys(refl_scaled: np.ndarray, idx: np.ndarray, refl_min: np.float32, seed: int) -> np.ndarray:
    """Efraimidis-Spirakis keys -log(u_i) / w_i (float64), w = refl - min + 1e-8 (:99,104, FP32),
    u = hash(seed, i) in (0, 1]: ascending key = draw order of weighted sampling without replacement."""
    with np.errstate(over="ignore"):
        w = (refl_scaled[idx] - refl_min + np.float32(1e-8)).astype(np.float32)
        h = _mix32((idx.astype(np.uint32) * np.uint32(0x9E3779B9) + np.uint32(seed & 0xFFFFFFFF)).astype(np.uint32))
    u = ((h >> np.uint32(8)).astype(np.float64) + 1.0) * 5.9604644775390625e-08
    return np.maximum(-np.log(u) / w.astype(np.float64), 0.0)


def _mix64(z: np.ndarray) -> np.ndarray:
    z = z.astype(np.uint64)
    with np.errstate(over="ignore"):
        z ^= z >> np.uint64(30)
        z = z * np.uint64(0xBF58476D1CE4E5B9)
        z ^= z >> np.uint64(27)
        z = z * np.uint64(0x94D049BB133111EB)
        z ^= z >> np.uint64(31)
    return z


def replacement_picks(idx: np.ndarray, voxel_id: int, max_pts: int, seed: int, grid_ordinal: int) -> np.ndarray:
    """:120: max_pts uniform draws WITH replacement; draw s takes member hash(seed, voxel, s) mod n."""
    sd = np.uint64((seed & 0xFFFFFFFF) | (grid_ordinal << 32))
    with np.errstate(over="ignore"):
        base = _mix64(np.array([sd ^ (np.uint64(voxel_id) * np.uint64(0x9E3779B97F4A7C15))], dtype=np.uint64))[0]
        h = _mix64(base + np.arange(max_pts, dtype=np.uint64))
    return idx[(h % np.uint64(len(idx))).astype(np.int64)]


def tile(feat5: np.ndarray, gridsize=(2.0, 4.0), min_pts=128, max_pts=16384, seed=SUBSAMPLE_SEED, weighted=True):
    """grid() + the > max_pts branch of write_voxels (:55-64, 116-120) -> list of index arrays.  `feat5` may hold more
    than five columns (further scalar fields of the file between reflectance and n_z): ALL of them are voxelised (:58)."""
    tiles, grids = [], []
    refl_min = feat5[:, 3].min()
    for gi, size in enumerate(gridsize):
        ids = O.grid(feat5, np.full(feat5.shape[1], size, np.float32))
        order = np.argsort(ids, kind="stable")
        sid = ids[order]
        starts = np.concatenate([[0], np.nonzero(sid[1:] != sid[:-1])[0] + 1, [len(sid)]])
        for a, b in zip(starts[:-1], starts[1:]):
            if b - a < min_pts:
                continue
            idx = order[a:b]
            if len(idx) > max_pts:
                if weighted:
                    idx = idx[np.argsort(sampling_keys(feat5[:, 3], idx, refl_min, seed), kind="stable")[:max_pts]]
                else:
                    idx = replacement_picks(idx, int(sid[a]), max_pts, seed, gi)
            tiles.append(idx.astype(np.int64))
            grids.append(size)
    return tiles, np.asarray(grids, np.float32)


def preprocess(cloud: np.ndarray, gridsize=(2.0, 4.0), min_pts=128, max_pts=16384, seed=SUBSAMPLE_SEED):
    """write_voxels (:79-127): returns (feat5 [N,5], tiles)."""
    cloud = np.ascontiguousarray(cloud, dtype=np.float32)
    n_z = ground_normalize(cloud[:, :3])
    refl = cloud[:, 3]
    weighted = not np.all(refl == 0)                                   # reflectance_not_zero (:94)
    if np.isnan(refl).any():
        raise ValueError("Input reflectance tensor contains NaN values.")          # :20-21
    if weighted:
        refl = quantile_normalize_reflectance(refl)
    feat5 = np.concatenate([cloud[:, :3], refl[:, None], n_z[:, None]], 1).astype(np.float32)
    grid_feat = np.concatenate([feat5[:, :4], cloud[:, 4:], feat5[:, 4:5]], 1).astype(np.float32) if cloud.shape[1] > 4 else feat5
    tiles, grids = tile(grid_feat, gridsize, min_pts, max_pts, seed, weighted)
    return feat5, tiles, grids


def pack(feat5: np.ndarray, tiles):
    """TestingDataset.__getitem__ + collate (src/predicter.py:78-94)."""
    pos, refl, batch, shift, sf = [], [], [], [], []
    for b, idx in enumerate(tiles):
        rows = feat5[idx]
        mean = (rows[:, :3].astype(np.float64).sum(0) / len(rows)).astype(np.float32)
        p = (rows[:, :3] - mean).astype(np.float32)
        sq = p * p
        sf.append(np.sqrt((sq[:, 0] + sq[:, 1]) + sq[:, 2]).max())
        pos.append(p)
        refl.append(rows[:, 3])
        batch.append(np.full(len(rows), b, np.int64))
        shift.append(mean)
    return (np.concatenate(pos), np.concatenate(refl), np.concatenate(batch), np.stack(shift),
            np.asarray(sf, np.float32))


def classify(sd, feat5, tiles, batch_size=8, is_wood=0.5, max_batches=None):
    """The inference loop (src/predicter.py:193-217) -> classified rows float64 [M,5]."""
    out = []
    nb = 0
    for t0 in range(0, len(tiles), batch_size):
        if max_batches is not None and nb >= max_batches:
            break
        group = tiles[t0:t0 + batch_size]
        pos, refl, batch, shift, sf = pack(feat5, group)
        logits = ref_model.net_forward(sd, torch.from_numpy(pos), torch.from_numpy(refl), torch.from_numpy(batch),
                                       torch.from_numpy(sf))
        prob = torch.sigmoid(torch.nan_to_num(logits)).numpy()
        pred = (prob >= is_wood).astype(np.float64)
        xyz = pos.astype(np.float64) + shift.astype(np.float64)[batch]
        out.append(np.concatenate([xyz, pred[:, None], prob.astype(np.float64)[:, None]], 1))
        nb += 1
    return np.concatenate(out) if out else np.zeros((0, 5))


def collect_predictions(classification, original_xyz, any_wood=1, workers=1):
    """PointCloudClassifier.collect_predictions + compute_labels (src/predicter.py:113-142): float64 KD-tree
    over the classified rows (pykdtree in the reference, scipy's cKDTree here: same exact k nearest
    neighbours up to ties), k = 64 (32 if any_wood != 1); pwood = np.median of the neighbours'
    probabilities; label = argmax over the class votes sum_{pred == c} prob (any_wood == 1), else
    any(pred > any_wood).  Returns (label float64 [N], pwood float64 [N])."""
    from scipy.spatial import cKDTree
    k = 32 if any_wood != 1 else 64
    tree = cKDTree(classification[:, :3])
    _, idx = tree.query(np.asarray(original_xyz, dtype=np.float64)[:, :3], k=k, workers=workers)
    nb = classification[idx]                                   # [N, k, 5]
    pwood = np.median(nb[:, :, -1], axis=1)
    if any_wood != 1:
        label = (nb[:, :, -2] > any_wood).any(axis=1).astype(np.float64)
    else:
        votes = np.stack([((nb[:, :, -2] == c) * nb[:, :, -1]).sum(axis=1) for c in (0, 1)], axis=1)
        # the reference sizes class_votes by k: classes >= 2 never occur, argmax = first max
        label = np.argmax(votes, axis=1).astype(np.float64)
    return label, pwood
