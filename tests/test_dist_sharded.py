"""World-size-2/3 gloo tests (CPU) of the sharded-plot plan (pointstowood_b200/distributed.py): statistics,
voxel-table merge, tile ownership, the member all-to-all, thinning at the owner, slab routing with halo, the
exactness check and the way back.  The per-point kernels are replaced by an oracle-backed stand-in (test
infrastructure); what is under test is the exchange plan, which must reproduce the single-process tiling and
vote exactly."""
import math
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

PARAMS = dict(min_pts=64, max_pts=1500, grid_size=(2.0, 4.0), batch_size=8)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _oracle_kernels():
    from oracle import oracle as O
    from oracle import ref_pipeline
    from pointstowood_b200.distributed import _Kernels

    class OracleKernels(_Kernels):
        def colminmax(self, a):
            if a.size(0) == 0:
                return torch.full((a.size(1),), math.inf), torch.full((a.size(1),), -math.inf)
            return a.min(0).values.clone(), a.max(0).values.clone()

        @staticmethod
        def _cells(v, lo, nb):
            edges = (np.float32(lo) + np.float32(5.0) * np.arange(nb, dtype=np.float32)).astype(np.float32)
            return np.searchsorted(edges, v, side="left")

        def ground_min(self, cloud, mn_xy, nbx, nby):
            c = cloud.numpy()
            cid = self._cells(c[:, 0], mn_xy[0].item(), nbx) * (nby + 1) + self._cells(c[:, 1], mn_xy[1].item(), nby)
            mins = np.full((nbx + 1) * (nby + 1), np.inf, dtype=np.float32)
            np.minimum.at(mins, cid, c[:, 2])
            return torch.from_numpy(mins)

        def ground_apply(self, cloud, mn_xy, nbx, nby, cell_min):
            c = cloud.numpy()
            cid = self._cells(c[:, 0], mn_xy[0].item(), nbx) * (nby + 1) + self._cells(c[:, 1], mn_xy[1].item(), nby)
            return torch.from_numpy((c[:, 2] - cell_min.numpy()[cid]).astype(np.float32))

        def reflectance_keys(self, cloud):
            b = cloud[:, 3].contiguous().numpy().view(np.uint32).astype(np.int64)
            return torch.from_numpy(np.where(b & 0x80000000, (~b) & 0xFFFFFFFF, b | 0x80000000))

        def reflectance_values(self, order, rank0, n_total):
            o = order.numpy().astype(np.int64)
            v = np.empty(len(o), np.float32)
            v[o] = ref_pipeline.quantile_values(rank0 + np.arange(len(o)), n_total)
            mnmx = np.array([v.min() if len(v) else np.inf, v.max() if len(v) else -np.inf], np.float32)
            return torch.from_numpy(v), torch.from_numpy(mnmx)

        def reflectance_scale(self, v, mnmx):
            mn, mx = mnmx.numpy().astype(np.float32)
            return torch.from_numpy((np.float32(2.0) * (v.numpy() - mn) / (mx - mn) - np.float32(1.0)).astype(np.float32))

        def assemble5(self, cloud, refl, n_z):
            r = cloud[:, 3] if refl is None else refl
            return torch.cat([cloud[:, :3], r[:, None], n_z[:, None]], 1).contiguous()

        def grid_ids(self, feat, size, start, end):
            return torch.from_numpy(O.grid(feat.numpy(), np.full(5, size, np.float32), start.numpy(), end.numpy()))

        def stable_order(self, keys, bits):
            assert keys.numel() == 0 or int(keys.max()) < (1 << bits)
            s, o = torch.sort(keys, stable=True)
            return s, o.to(torch.int32)

        def segments(self, sorted_keys, order):
            k = sorted_keys.numpy()
            heads = np.concatenate([[0], np.nonzero(k[1:] != k[:-1])[0] + 1, [len(k)]]).astype(np.int64)
            buf = np.zeros(len(k) + 1, np.int64)
            buf[: len(heads)] = heads
            return torch.from_numpy(buf), torch.tensor([len(heads) - 1])

        def thin(self, feat, refl_col, members, sizes, global_index, refl_min, weighted, voxel_ids, maxpoints, seed, grid_ordinal):
            m, g = members.numpy().astype(np.int64), global_index.numpy().astype(np.int64)
            out, o = [], 0
            for t, s in enumerate(sizes):
                rows = m[o:o + s]
                o += s
                if weighted:
                    # keys are a function of the POINT index; the weights come from this rank's rows
                    lookup = np.zeros(int(g.max()) + 1, np.float32)
                    lookup[g[rows]] = feat.numpy()[rows, refl_col]
                    keys = ref_pipeline.sampling_keys(lookup, g[rows], np.float32(refl_min), seed)
                    out.append(rows[np.argsort(keys, kind="stable")[:maxpoints]])
                else:
                    out.append(ref_pipeline.replacement_picks(rows, int(voxel_ids[t]), maxpoints, seed, grid_ordinal))
            return torch.from_numpy(np.concatenate(out).astype(np.int32))

        def vote(self, rows_xyz, prob, pred, queries, k, any_wood):
            return _vote(rows_xyz.numpy(), prob.numpy(), pred.numpy(), queries.numpy(), k)

    return OracleKernels()


def _vote(rows, prob, pred, queries, k):
    """(label, pwood, nbr) from the C oracle's exact k-NN, ties by (d2, index)."""
    from oracle import oracle as O
    nbr = O.knn(rows, queries, k) if len(rows) and len(queries) else np.full((len(queries), k), -1, np.int64)
    ok = nbr >= 0
    p = np.where(ok, prob[np.maximum(nbr, 0)].astype(np.float64), np.nan)
    pwood = np.nanmedian(p, axis=1) if len(queries) else np.zeros(0)
    w = np.where(ok & (pred[np.maximum(nbr, 0)] == 1), p, 0.0).sum(1)
    l = np.where(ok & (pred[np.maximum(nbr, 0)] == 0), p, 0.0).sum(1)
    return (torch.from_numpy((w > l).astype(np.uint8)), torch.from_numpy(pwood.astype(np.float64)),
            torch.from_numpy(nbr.astype(np.int32)))


def _fake_prob(xyz, rows):
    """A stand-in for the network: depends on the coordinates AND on the global row, so duplicated points
    (one row per tile that holds them) carry different values, as real classifications do."""
    rows = np.asarray(rows, dtype=np.int64)
    return ((np.sin(xyz[:, 0] * 3.1) * np.cos(xyz[:, 2] * 1.7) * 0.5 + 0.5) * 0.8 + 0.2 * ((rows * 2654435761) % 1000) / 1000.0).astype(np.float32)


def _cloud(weighted=True):
    from pointstowood_b200.synthetic import tls_plot
    cloud = tls_plot(24000, 11, side=6.5)[0]
    if not weighted:
        cloud[:, 3] = 0
    return cloud


def _single(cloud, weighted):
    """The single-process answer: oracle tiling, fake probabilities in tile-major row order, exact vote."""
    from oracle import ref_pipeline
    feat5, tiles, _ = ref_pipeline.preprocess(cloud, PARAMS["grid_size"], PARAMS["min_pts"], PARAMS["max_pts"])
    members = np.concatenate(tiles)
    xyz = feat5[members, :3]
    prob = _fake_prob(xyz, np.arange(len(xyz)))
    label, pwood, _ = _vote(xyz, prob, (prob >= 0.5).astype(np.uint8), cloud[:, :3].copy(), 64)
    return feat5, tiles, label.numpy(), pwood.numpy()


def _worker(rank, world, port, out, weighted, halo):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pointstowood_b200.distributed import Comm, ShardedPlot
    cloud = _cloud(weighted)
    n = len(cloud)
    cuts = [0, n // 3, n] if world == 2 else [0, n // 4, n // 4, n]            # uneven chunks; rank 1 of 3 is EMPTY
    chunk = torch.from_numpy(cloud[cuts[rank]:cuts[rank + 1]].copy())
    plot = ShardedPlot(chunk, Comm(), kernels=_oracle_kernels(), **PARAMS)
    store = plot.tile()
    feat = store.feat.numpy()
    members = store.members.numpy()
    xyz = feat[members, :3]
    prob = torch.from_numpy(_fake_prob(xyz, plot.global_rows.numpy()))
    label, pwood = plot.vote(torch.from_numpy(xyz.copy()), prob, 0.5, 1, halo)
    out.put(dict(rank=rank, local_tiles=plot.local_tiles.copy(), ptr=store.ptr.copy(), rows=feat[members].copy(), label=label.numpy(),
                 pwood=pwood.numpy(), lo=cuts[rank], hi=cuts[rank + 1], num_tiles=plot.num_tiles, rounds=plot.vote_rounds,
                 traffic=dict(plot.traffic), n_z=plot.n_z.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def _run(world, weighted, halo):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out, weighted, halo)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted([out.get(timeout=300) for _ in range(world)], key=lambda d: d["rank"])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    return got


@pytest.mark.parametrize("world,weighted,halo", [(2, True, 0.5), (3, True, 0.02), (2, False, 0.5)])
def test_sharded_plot_equals_single_process(world, weighted, halo):
    cloud = _cloud(weighted)
    feat5, tiles, label, pwood = _single(cloud, weighted)
    sizes = [len(t) for t in tiles]
    assert max(sizes) == PARAMS["max_pts"], "the fixture must exercise the thinning of oversized tiles"
    got = _run(world, weighted, halo)
    assert all(g["num_tiles"] == len(tiles) for g in got)
    assert sorted(np.concatenate([g["local_tiles"] for g in got]).tolist()) == list(range(len(tiles)))     # a partition
    for g in got:
        B = PARAMS["batch_size"]
        assert all((t // B) % world == g["rank"] for t in g["local_tiles"])  # whole batches, dealt round
        mine = [tiles[t] for t in g["local_tiles"]]
        assert np.array_equal(np.diff(g["ptr"]), [len(t) for t in mine])
        want = feat5[np.concatenate(mine), :4] if mine else np.zeros((0, 4), np.float32)
        assert np.array_equal(g["rows"], want)                              # same members, same order, same values
        assert np.array_equal(g["n_z"], feat5[g["lo"]:g["hi"], 4])
        assert np.array_equal(g["label"], label[g["lo"]:g["hi"]])
        assert np.array_equal(g["pwood"], pwood[g["lo"]:g["hi"]])
    if halo < 0.1:                 # a 2 cm halo cannot hold the 64 nearest rows: the bound check must have widened it
        assert got[0]["rounds"] > 1
    assert any("all-to-all: tile members" in g["traffic"] for g in got)
    if weighted:
        assert all("all-to-all: reflectance keys" in g["traffic"] for g in got)


def test_slab_bounds_and_halo_entries():
    from pointstowood_b200.distributed import halo_entries, slab_bounds
    rng = np.random.default_rng(0)
    x = rng.random(100000).astype(np.float32) * 50
    hist = np.histogram(x, bins=4096, range=(0, 50))[0]
    b = slab_bounds(hist, 0.0, 50.0, 4)
    counts = np.bincount(np.searchsorted(b, x, side="right"), minlength=4)
    assert counts.min() > 0.9 * 25000 and counts.max() < 1.1 * 25000
    keys, span = halo_entries(torch.from_numpy(x), torch.from_numpy(b), 1.0, 4, 3)
    keys = keys.view(-1, 3).numpy()
    for i in rng.integers(0, len(x), 200):
        want = [s for s in range(4) if (b[s - 1] if s else -np.inf) - 1.0 <= x[i] + 1e-4 and x[i] - 1e-4 < (b[s] if s < 3 else np.inf) + 1.0]
        have = sorted(k for k in keys[i] if k < 4)
        assert set(have) <= set(want) and len(have) == int(span[i]) and len(have) >= len(want) - 1
        assert int(np.searchsorted(b, x[i], side="right")) in have
