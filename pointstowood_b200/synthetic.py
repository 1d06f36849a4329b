"""Synthetic TLS-like inputs for parity tests and benchmarks.

Everything is float32, metres, plot origin (0, 0, 0), generated with
``numpy.random.Generator(PCG64(seed))`` so the same cloud is reproduced on the CPU
container and on the GPU box.  The reference ships no data (its ``.gitignore``
excludes ``data/``) and no checkpoint, so these clouds and seeded weights are the
workload every number in this repo is quoted on.
"""
from __future__ import annotations

import numpy as np

__all__ = ["tls_plot", "tls_plot_blocks", "uniform_tiles", "write_ply"]


def _cylinder(rng, n, base, axis, length, radius, noise):
    """n points on the surface of a cylinder starting at `base` along unit `axis`."""
    axis = axis / np.linalg.norm(axis)
    ref = np.array([1.0, 0.0, 0.0]) if abs(axis[0]) < 0.9 else np.array([0.0, 1.0, 0.0])
    u = np.cross(axis, ref)
    u /= np.linalg.norm(u)
    v = np.cross(axis, u)
    t = rng.random(n) * length
    th = rng.random(n) * 2.0 * np.pi
    r = radius + rng.normal(0.0, noise, n)
    return (base[None, :] + t[:, None] * axis[None, :]
            + (r * np.cos(th))[:, None] * u[None, :] + (r * np.sin(th))[:, None] * v[None, :])


def tls_plot(n_points: int, seed: int = 1, side: float | None = None):
    """TLS-like forest plot: returns (xyzr float32 [N,4], label uint8 [N]; wood=1).

    10 % ground, one tree per jittered 4 m cell: stem 25 %, branches 15 %, leaf discs 50 %.
    Constant areal density of 2 500 points / m^2 (1 M points -> 20 x 20 m).
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    if side is None:
        side = float(np.sqrt(n_points / 2500.0))
    ncell = max(1, int(round(side / 4.0)))
    cell = side / ncell
    ntree = ncell * ncell
    n_ground = n_points // 10
    per_tree = (n_points - n_ground) // ntree
    out = np.empty((n_points, 4), dtype=np.float32)
    lab = np.zeros(n_points, dtype=np.uint8)

    gx = rng.random(n_ground) * side
    gy = rng.random(n_ground) * side
    gz = 0.3 * np.sin(gx / 7.0) * np.cos(gy / 9.0) + rng.normal(0.0, 0.02, n_ground)
    out[:n_ground, 0], out[:n_ground, 1], out[:n_ground, 2] = gx, gy, gz
    out[:n_ground, 3] = rng.normal(-8.0, 2.0, n_ground)
    o = n_ground
    for ti in range(ntree):
        cx = (ti % ncell + 0.5 + rng.uniform(-0.3, 0.3)) * cell
        cy = (ti // ncell + 0.5 + rng.uniform(-0.3, 0.3)) * cell
        cz = 0.3 * np.sin(cx / 7.0) * np.cos(cy / 9.0)
        n_t = per_tree if ti < ntree - 1 else n_points - o
        n_stem = int(n_t * 25 / 90)
        n_br = int(n_t * 15 / 90)
        n_leaf = n_t - n_stem - n_br
        h = rng.uniform(10.0, 25.0)
        r = rng.uniform(0.1, 0.4)
        base = np.array([cx, cy, cz])
        stem = _cylinder(rng, n_stem, base, np.array([rng.normal(0, .02), rng.normal(0, .02), 1.0]), h, r, 0.003)
        nb = int(rng.integers(5, 11))
        parts = []
        left = n_br
        for b in range(nb):
            nbp = left if b == nb - 1 else n_br // nb
            left -= nbp
            zb = rng.uniform(0.35, 0.95) * h
            az = rng.uniform(0, 2 * np.pi)
            axis = np.array([np.cos(az), np.sin(az), rng.uniform(0.1, 0.8)])
            parts.append(_cylinder(rng, nbp, base + np.array([0, 0, zb]), axis,
                                   rng.uniform(1.0, 3.0), rng.uniform(0.01, 0.05), 0.003))
        branches = np.concatenate(parts, 0) if parts else np.empty((0, 3))
        # leaves: discs of ~20 points, 3 cm radius, inside a crown ellipsoid
        ndisc = max(1, n_leaf // 20)
        cen = rng.normal(0.0, 1.0, (ndisc, 3))
        cen /= np.maximum(np.linalg.norm(cen, axis=1, keepdims=True), 1e-9)
        cen *= (rng.random((ndisc, 1)) ** (1 / 3))
        cen = cen * np.array([1.8, 1.8, 0.3 * h]) + np.array([cx, cy, cz + 0.65 * h])
        nrm = rng.normal(0.0, 1.0, (ndisc, 3))
        nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
        which = rng.integers(0, ndisc, n_leaf)
        ref = np.where(np.abs(nrm[:, :1]) < 0.9, np.array([[1.0, 0, 0]]), np.array([[0, 1.0, 0]]))
        du = np.cross(nrm, ref)
        du /= np.linalg.norm(du, axis=1, keepdims=True)
        dv = np.cross(nrm, du)
        rr = 0.03 * np.sqrt(rng.random(n_leaf))
        th = rng.random(n_leaf) * 2 * np.pi
        leaves = cen[which] + (rr * np.cos(th))[:, None] * du[which] + (rr * np.sin(th))[:, None] * dv[which]
        nw = n_stem + n_br
        out[o:o + n_stem, :3] = stem
        out[o + n_stem:o + nw, :3] = branches
        out[o + nw:o + n_t, :3] = leaves
        out[o:o + nw, 3] = rng.normal(-6.0, 1.5, nw)
        out[o + nw:o + n_t, 3] = rng.normal(-11.0, 2.0, n_leaf)
        lab[o:o + nw] = 1
        o += n_t
    perm = rng.permutation(n_points)          # scanners do not deliver points tree by tree
    return out[perm], lab[perm]


def tls_plot_blocks(n_points: int, seed: int = 2, rows: tuple | None = None, block_points: int = 1_000_000):
    """Large plots (BASELINE.json configs[3]: 100 M points) as a square arrangement of 20 x 20 m blocks of
    `block_points` points, block b = tls_plot(block_points, seed + b) moved to its place; the rows of a block are
    contiguous, like the scan positions merged into one TLS file.  Same areal density as tls_plot.
    `rows = (lo, hi)` returns only rows lo .. hi of the plot (a rank of a sharded run generates its own chunk:
    only the blocks that overlap the range are built).  Returns (xyzr float32 [hi-lo, 4], label uint8)."""
    nblk = max(1, -(-n_points // block_points))
    per_side = int(np.ceil(np.sqrt(nblk)))
    side = float(np.sqrt(block_points / 2500.0))
    lo, hi = (0, n_points) if rows is None else rows
    parts, labs = [], []
    for b in range(lo // block_points, min(nblk, -(-hi // block_points))):
        nb = min(block_points, n_points - b * block_points)
        cloud, lab = tls_plot(nb, seed + b, side=side)
        cloud[:, 0] += np.float32((b % per_side) * side)
        cloud[:, 1] += np.float32((b // per_side) * side)
        a0, a1 = max(lo - b * block_points, 0), min(hi - b * block_points, nb)
        parts.append(cloud[a0:a1])
        labs.append(lab[a0:a1])
    if not parts:
        return np.zeros((0, 4), np.float32), np.zeros(0, np.uint8)
    return np.concatenate(parts), np.concatenate(labs)


def uniform_tiles(n_tiles: int, n_per_tile: int = 16384, side: float = 2.0, seed: int = 3):
    """Micro-bench tiles (config 3): `n_tiles` x `n_per_tile` points uniform in a cube.

    Returns (pos float32 [B*n,3], ptr int64 [B+1]).
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    pos = (rng.random((n_tiles * n_per_tile, 3), dtype=np.float32) * np.float32(side)).astype(np.float32)
    ptr = np.arange(n_tiles + 1, dtype=np.int64) * n_per_tile
    return pos, ptr


def write_ply(path: str, xyzr: np.ndarray) -> None:
    """Binary little-endian PLY with `property float x,y,z,reflectance` (parses with the
    reference reader, /root/reference/pointstowood/src/io.py:11-47)."""
    xyzr = np.ascontiguousarray(xyzr, dtype="<f4")
    with open(path, "wb") as f:
        f.write(b"ply\nformat binary_little_endian 1.0\n")
        f.write(f"element vertex {xyzr.shape[0]}\n".encode())
        for name in ("x", "y", "z", "reflectance"):
            f.write(f"property float {name}\n".encode())
        f.write(b"end_header\n")
        xyzr.tofile(f)
