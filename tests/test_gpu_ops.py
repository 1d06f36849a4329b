"""GPU parity tests: libp2w (through the C ABI, via pointstowood_b200.ops) against the CPU
oracle on the same seeded inputs.  Integer outputs must be bit-exact."""
import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from pointstowood_b200 import ops as _ops
    return _ops


def _dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def _tiles(rng, sizes_x, sizes_y, dup=False, lattice=False):
    px = np.concatenate([[0], np.cumsum(sizes_x)]).astype(np.int64)
    py = np.concatenate([[0], np.cumsum(sizes_y)]).astype(np.int64)
    x = rng.random((px[-1], 3), dtype=np.float32) * 2
    y = rng.random((py[-1], 3), dtype=np.float32) * 2
    if lattice:      # many exact distance ties
        x = np.round(x * 8) / 8
        y = np.round(y * 8) / 8
    if dup and len(x) > 10:
        x[5:10] = x[0:5]
    return x.astype(np.float32), y.astype(np.float32), px, py


KNN_CASES = [
    dict(sx=[1000], sy=[300], k=32),
    dict(sx=[5000, 3000, 17, 0, 2100], sy=[700, 900, 40, 5, 0], k=32),       # short + empty tiles
    dict(sx=[4097, 2049], sy=[513, 1025], k=16, lattice=True),                # ties
    dict(sx=[3000], sy=[3000], k=2, dup=True),
    dict(sx=[2500, 2500], sy=[100, 100], k=64),
    dict(sx=[1500], sy=[64], k=100),
    dict(sx=[20], sy=[33], k=32),
    dict(sx=[1], sy=[1], k=1),
]


@pytest.mark.parametrize("method", ["sweep", "grid"])
@pytest.mark.parametrize("case", KNN_CASES)
def test_knn_table_bit_exact(ops, case, method):
    rng = np.random.default_rng(hash(str(case)) % 2**32)
    x, y, px, py = _tiles(rng, case["sx"], case["sy"], case.get("dup", False), case.get("lattice", False))
    ref, ref_d = O.knn(x, y, case["k"], px, py, return_d2=True)
    nbr, d2 = ops.knn_table(_dev(x), _dev(y), case["k"], _dev(px), _dev(py), return_d2=True, method=method)
    assert np.array_equal(nbr.cpu().numpy().astype(np.int64), ref)
    assert np.array_equal(d2.cpu().numpy(), ref_d)


def test_knn_misaligned_base_pointer(ops):
    rng = np.random.default_rng(5)
    x, y, px, py = _tiles(rng, [3001], [257])
    big = _dev(np.concatenate([np.zeros((1, 3), np.float32), x]))
    xs = big[1:]                                   # 12-byte offset: the non-TMA staging path
    assert xs.data_ptr() % 16 != 0
    nbr = ops.knn_table(xs, _dev(y), 8, _dev(px), _dev(py))
    assert np.array_equal(nbr.cpu().numpy().astype(np.int64), O.knn(x, y, 8, px, py))


def test_knn_edge_index_matches_upstream_layout(ops):
    rng = np.random.default_rng(6)
    x, y, px, py = _tiles(rng, [900, 10], [100, 50])
    bx = np.repeat(np.arange(2), [900, 10])
    by = np.repeat(np.arange(2), [100, 50])
    e = ops.knn(_dev(x), _dev(y), 16, _dev(bx), _dev(by))
    assert e.dtype == torch.int64
    assert np.array_equal(e.cpu().numpy(), O.table_to_edges(O.knn(x, y, 16, px, py)))


@pytest.mark.parametrize("case", [
    dict(sx=[6000], sy=[2000], r=0.15, m=32),
    dict(sx=[3000, 0, 4500], sy=[500, 3, 800], r=0.3, m=32),     # heavy truncation
    dict(sx=[2048 * 3 + 5], sy=[777], r=0.05, m=8),
    dict(sx=[4000], sy=[1000], r=0.25, m=64, lattice=True),
])
@pytest.mark.parametrize("method", ["sweep", "grid"])
def test_radius_table_bit_exact(ops, case, method):
    rng = np.random.default_rng(hash(str(case)) % 2**32)
    x, y, px, py = _tiles(rng, case["sx"], case["sy"], lattice=case.get("lattice", False))
    ref, ref_cnt = O.radius(x, y, case["r"], px, py, case["m"])
    nbr, cnt = ops.radius_table(_dev(x), _dev(y), case["r"], _dev(px), _dev(py), case["m"], method=method)
    assert np.array_equal(cnt.cpu().numpy(), ref_cnt)
    assert np.array_equal(nbr.cpu().numpy().astype(np.int64), ref)
    bx = np.repeat(np.arange(len(px) - 1), np.diff(px))
    by = np.repeat(np.arange(len(py) - 1), np.diff(py))
    e = ops.radius(_dev(x), _dev(y), case["r"], _dev(bx), _dev(by), case["m"], batch_size=len(px) - 1)
    assert np.array_equal(e.cpu().numpy(), O.table_to_edges(ref))


@pytest.mark.parametrize("method", ["sweep", "grid"])
def test_radius_sa1_subset_queries(ops, method):
    """The SA1 call shape: y = x[idx], r = 0.08, max 32 on a dense cloud (src/model.py:118)."""
    from pointstowood_b200.synthetic import tls_plot
    p, _ = tls_plot(40000, 21, side=4.0)
    x = p[:, :3] - p[:, :3].mean(0)
    idx = np.sort(np.random.default_rng(1).choice(len(x), 9000, replace=False))
    px, py = np.array([0, len(x)]), np.array([0, len(idx)])
    ref, ref_cnt = O.radius(x, x[idx], 0.08, px, py, 32)
    nbr, cnt = ops.radius_table(_dev(x), _dev(x[idx]), 0.08, _dev(px), _dev(py), 32, method=method)
    assert np.array_equal(nbr.cpu().numpy().astype(np.int64), ref)
    assert np.array_equal(cnt.cpu().numpy(), ref_cnt)
    assert (ref_cnt == 32).mean() > 0.1          # truncation really happens


def _tls_tiles(n, seed, side, n_side):
    from pointstowood_b200.synthetic import tls_plot
    p, _ = tls_plot(n, seed, side=side)
    cell = side / n_side
    tid = np.minimum((p[:, 0] / cell).astype(int), n_side - 1) * n_side + np.minimum((p[:, 1] / cell).astype(int),
                                                                                     n_side - 1)
    order = np.argsort(tid, kind="stable")
    x = np.ascontiguousarray(p[order, :3])
    ptr = np.concatenate([[0], np.cumsum(np.bincount(tid, minlength=n_side * n_side))]).astype(np.int64)
    return x, ptr


@pytest.mark.parametrize("k", [2, 16, 32])
def test_grid_knn_on_tls_tiles_matches_sweep_and_oracle(ops, k):
    """Surface-like, strongly non-uniform tiles (the data the cell size is planned for); queries are a
    subset of the sources (SA levels) or a superset (FP levels)."""
    x, ptr = _tls_tiles(70000, 3, 6.0, 3)
    rng = np.random.default_rng(k)
    keep = np.sort(rng.choice(len(x), len(x) // 3, replace=False))
    sub = x[keep]
    ptr_sub = np.searchsorted(keep, ptr).astype(np.int64)
    for (src, psrc, qry, pqry) in ((x, ptr, sub, ptr_sub), (sub, ptr_sub, x, ptr)):
        grid, gd = ops.knn_table(_dev(src), _dev(qry), k, _dev(psrc), _dev(pqry), return_d2=True, method="grid")
        sweep, sd = ops.knn_table(_dev(src), _dev(qry), k, _dev(psrc), _dev(pqry), return_d2=True, method="sweep")
        assert torch.equal(grid, sweep) and torch.equal(gd, sd)
    ref = O.knn(sub, x[:5000], k, ptr_sub, np.array([0] + [5000] * 9))         # first tile against the oracle
    got = ops.knn_table(_dev(sub), _dev(x[:5000]), k, _dev(ptr_sub), _dev(np.array([0] + [5000] * 9)), method="grid")
    assert np.array_equal(got.cpu().numpy().astype(np.int64), ref)


def test_grid_knn_queries_far_outside_and_sparse_outliers(ops):
    """Queries outside the sources' bounding box and isolated sources many empty cells away force
    the shell expansion well past the first ring."""
    rng = np.random.default_rng(11)
    core = rng.normal(0, 0.05, (6000, 3)).astype(np.float32)
    far = (rng.random((40, 3), dtype=np.float32) * 8 - 4).astype(np.float32)
    x = np.concatenate([core, far]).astype(np.float32)
    y = np.concatenate([far[:20] + 0.01, rng.random((300, 3), dtype=np.float32) * 30 - 15,
                        core[:200]]).astype(np.float32)
    px, py = np.array([0, len(x)]), np.array([0, len(y)])
    for k in (1, 8, 32):
        ref, ref_d = O.knn(x, y, k, px, py, return_d2=True)
        nbr, d2 = ops.knn_table(_dev(x), _dev(y), k, _dev(px), _dev(py), return_d2=True, method="grid")
        assert np.array_equal(nbr.cpu().numpy().astype(np.int64), ref)
        assert np.array_equal(d2.cpu().numpy(), ref_d)
    ref, ref_cnt = O.radius(x, y, 0.6, px, py, 32)
    nbr, cnt = ops.radius_table(_dev(x), _dev(y), 0.6, _dev(px), _dev(py), 32, method="grid")
    assert np.array_equal(nbr.cpu().numpy().astype(np.int64), ref) and np.array_equal(cnt.cpu().numpy(), ref_cnt)


def test_grid_knn_degenerate_tiles(ops):
    """All sources identical / collinear / k larger than the tile."""
    same = np.full((500, 3), 0.25, np.float32)
    line = np.zeros((700, 3), np.float32)
    line[:, 0] = np.linspace(0, 1, 700, dtype=np.float32)
    few = np.random.default_rng(2).random((40, 3), dtype=np.float32)
    x = np.concatenate([same, line, few]).astype(np.float32)
    px = np.array([0, 500, 1200, 1240])
    y = np.random.default_rng(3).random((90, 3), dtype=np.float32)
    py = np.array([0, 30, 60, 90])
    for k in (4, 64):
        ref = O.knn(x, y, k, px, py)
        nbr = ops.knn_table(_dev(x), _dev(y), k, _dev(px), _dev(py), method="grid")
        assert np.array_equal(nbr.cpu().numpy().astype(np.int64), ref)


@pytest.mark.parametrize("sizes,ratio", [([3000], 0.05), ([1000, 1, 17000, 250], 0.01), ([64], 1.0)])
def test_fps_bit_exact(ops, sizes, ratio):
    rng = np.random.default_rng(3)
    ptr = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    src = rng.random((ptr[-1], 3), dtype=np.float32)
    ref = O.fps(src, ptr, ratio)
    out = ops.fps(_dev(src), ratio=ratio, random_start=False, ptr=_dev(ptr))
    assert np.array_equal(out.cpu().numpy(), ref)


@pytest.mark.parametrize("n,dim,size", [(50000, 3, 0.04), (20000, 5, 2.0), (1000, 1, 0.5)])
def test_grid_cluster_bit_exact(ops, n, dim, size):
    rng = np.random.default_rng(n)
    pos = (rng.random((n, dim), dtype=np.float32) * 7 - 3).astype(np.float32)
    sz = np.full(dim, size, np.float32)
    ids = ops.grid_cluster(_dev(pos), _dev(sz))
    assert np.array_equal(ids.cpu().numpy(), O.grid(pos, sz))


def test_voxel_grid_and_consecutive_cluster(ops):
    rng = np.random.default_rng(9)
    sizes = [7000, 9000, 300]
    pos = (rng.normal(0, 0.7, (sum(sizes), 3))).astype(np.float32)
    batch = np.repeat(np.arange(3), sizes).astype(np.int64)
    ref_ids = O.voxel_grid(pos, 0.08, batch)
    ids = ops.voxel_grid(_dev(pos), 0.08, _dev(batch))
    assert np.array_equal(ids.cpu().numpy(), ref_ids)
    ref_inv, ref_perm = O.consecutive_cluster(ref_ids)
    inv, perm = ops.consecutive_cluster(ids)
    assert np.array_equal(perm.cpu().numpy(), ref_perm)
    assert np.array_equal(inv.cpu().numpy(), ref_inv)
    idx = ops.voxel_sample(_dev(pos), 0.08, _dev(batch))
    assert np.array_equal(idx.cpu().numpy(), ref_perm)
    idx = ops.voxel_sample(_dev(pos), 0.08, _dev(batch), key_bits=8)     # overflow -> 64-bit retry
    assert np.array_equal(idx.cpu().numpy(), ref_perm)


@pytest.mark.parametrize("n,bits", [(1, 8), (2049, 11), (300000, 37), (70000, 64)])
def test_sort_pairs_is_a_stable_sort(ops, n, bits):
    rng = np.random.default_rng(n)
    keys = rng.integers(0, 2 ** min(bits, 62), n, dtype=np.int64)
    keys[: n // 3] = keys[n // 3: 2 * (n // 3)][: n // 3] if n >= 3 else keys[: n // 3]   # duplicates
    ks, vs = ops.sort_pairs(_dev(keys), bits)
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(vs.cpu().numpy(), order.astype(np.int32))
    assert np.array_equal(ks.cpu().numpy(), keys[order])


def test_scatter_max_min_and_global_max_pool(ops):
    rng = np.random.default_rng(4)
    n, c, m = 5000, 7, 40
    src = rng.normal(size=(n, c)).astype(np.float32)
    index = np.sort(rng.integers(0, m - 3, n)).astype(np.int64)       # slots m-3.. stay empty
    out, arg = ops.scatter_max(_dev(src), _dev(index), dim=0, dim_size=m)
    ref = O.scatter_max(src, index, m)
    assert np.array_equal(out.cpu().numpy(), ref)
    a = arg.cpu().numpy()
    filled = np.isin(np.arange(m), index)
    assert (a[~filled] == n).all()
    assert np.array_equal(src[a[filled], np.arange(c)[None, :]], ref[filled])
    out_min, _ = ops.scatter_min(_dev(src[:, 0]), _dev(index), dim_size=m)
    assert np.array_equal(out_min.cpu().numpy(), -O.scatter_max(-src[:, 0], index, m))
    ptr = np.searchsorted(index, np.arange(m + 1)).astype(np.int64)
    pooled = ops.global_max_pool(_dev(src), _dev(index), size=m)
    assert np.array_equal(pooled.cpu().numpy(), ref)
    assert ptr[-1] == n


def test_knn_interpolate_matches_oracle(ops):
    from oracle import ref_model
    rng = np.random.default_rng(8)
    sx, sy = [400, 1, 90], [3000, 50, 700]
    px = rng.random((sum(sx), 3), dtype=np.float32)
    py = rng.random((sum(sy), 3), dtype=np.float32)
    py[:50] = px[:50]                                   # zero distances -> clamp 1e-16
    bx = np.repeat(np.arange(3), sx).astype(np.int64)
    by = np.repeat(np.arange(3), sy).astype(np.int64)
    x = rng.normal(size=(sum(sx), 96)).astype(np.float32)
    ref = ref_model.knn_interpolate(torch.from_numpy(x), torch.from_numpy(px), torch.from_numpy(py),
                                    torch.from_numpy(bx), torch.from_numpy(by), 2, 3).numpy()
    out = ops.knn_interpolate(_dev(x), _dev(px), _dev(py), _dev(bx), _dev(by), k=2)
    assert np.abs(out.cpu().numpy() - ref).max() <= 1e-5 * max(1.0, np.abs(ref).max())


def test_pack_and_writeback(ops):
    rng = np.random.default_rng(10)
    n = 30000
    cloud = np.concatenate([rng.random((n, 3)) * 20 + 100, rng.normal(size=(n, 2))], 1).astype(np.float32)
    sizes = [9000, 1, 5000]
    index = rng.permutation(n)[: sum(sizes)].astype(np.int64)
    ptr = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    pos, refl, batch, shift, sf = ops.pack_tiles(_dev(cloud), _dev(index), _dev(ptr))
    for b in range(3):
        rows = cloud[index[ptr[b]:ptr[b + 1]]]
        mean = (rows[:, :3].astype(np.float64).sum(0) / len(rows)).astype(np.float32)
        assert np.array_equal(shift[b].cpu().numpy(), mean)
        p = rows[:, :3] - mean
        assert np.array_equal(pos[ptr[b]:ptr[b + 1]].cpu().numpy(), p)
        nrm = np.sqrt((p * p)[:, 0] + (p * p)[:, 1] + (p * p)[:, 2])
        assert sf[b].item() == nrm.max()
        assert np.array_equal(refl[ptr[b]:ptr[b + 1]].cpu().numpy(), rows[:, 3])
        assert (batch[ptr[b]:ptr[b + 1]] == b).all()
    logits = rng.normal(size=sum(sizes)).astype(np.float32) * 4
    logits[:3] = [np.nan, np.inf, -np.inf]
    prob, pred, rows = ops.writeback(_dev(logits), pos, _dev(ptr), shift, 0.5, want_rows=True)
    ref_p = torch.sigmoid(torch.nan_to_num(torch.from_numpy(logits))).numpy()
    assert np.abs(prob.cpu().numpy() - ref_p).max() < 1e-6
    assert np.array_equal(pred.cpu().numpy(), (prob.cpu().numpy() >= 0.5).astype(np.uint8))
    r = rows.cpu().numpy()
    b_of = np.repeat(np.arange(3), sizes)
    xyz = pos.cpu().numpy().astype(np.float64) + shift.cpu().numpy().astype(np.float64)[b_of]
    assert np.array_equal(r[:, :3], xyz)
    assert np.array_equal(r[:, 3], pred.cpu().numpy().astype(np.float64))


def test_errors_are_loud(ops):
    from pointstowood_b200._lib import P2WError
    x = torch.rand(10, 3, device="cuda")
    with pytest.raises(P2WError):
        ops.knn(x, x, 101)
    with pytest.raises(P2WError):
        ops.knn_table(x.cpu(), x, 4, torch.tensor([0, 10]).cuda(), torch.tensor([0, 10]).cuda())
    with pytest.raises(P2WError):
        ops.knn(x, x, 4, cosine=True)


def test_torch_ops_namespace_matches_upstream_schemas(ops):
    """torch.ops.p2w.* take the dispatcher-level arguments of torch_cluster / torch_scatter."""
    import pointstowood_b200.torch_ops  # noqa: F401
    rng = np.random.default_rng(12)
    x, y, px, py = _tiles(rng, [700, 300], [90, 40])
    e = torch.ops.p2w.knn(_dev(x), _dev(y), _dev(px), _dev(py), 8, False, 1)
    assert np.array_equal(e.cpu().numpy(), O.table_to_edges(O.knn(x, y, 8, px, py)))
    e = torch.ops.p2w.radius(_dev(x), _dev(y), _dev(px), _dev(py), 0.2, 16, 1, False)
    assert np.array_equal(e.cpu().numpy(), O.table_to_edges(O.radius(x, y, 0.2, px, py, 16)[0]))
    out = torch.ops.p2w.fps(_dev(x), _dev(px), torch.tensor([0.1]).cuda(), False)
    assert np.array_equal(out.cpu().numpy(), O.fps(x, px, 0.1))
    sz = np.array([0.1, 0.1, 0.1], np.float32)
    assert np.array_equal(torch.ops.p2w.grid(_dev(x), _dev(sz), None, None).cpu().numpy(), O.grid(x, sz))
    src = rng.normal(size=(1000, 1)).astype(np.float32)
    idx = np.sort(rng.integers(0, 50, 1000)).astype(np.int64)
    m, _ = torch.ops.p2w.scatter_max(_dev(src), _dev(idx), 0, None, 50)
    assert np.array_equal(m.cpu().numpy(), O.scatter_max(src, idx, 50))


def test_fps_random_start_follows_torchs_generator(ops):
    """random_start=True (upstream's default): the start row of every example is (rand(B) * deg).long() from
    torch's CUDA generator; the rest is the deterministic farthest-point iteration from that row."""
    rng = np.random.default_rng(12)
    sizes = [700, 1, 333, 2048]
    src = rng.random((sum(sizes), 3)).astype(np.float32)
    ptr = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    torch.manual_seed(123)
    out = ops.fps(_dev(src), ratio=0.25, random_start=True, ptr=_dev(ptr)).cpu().numpy()
    torch.manual_seed(123)
    deg = torch.tensor(sizes, device="cuda")
    start = ptr[:-1] + (torch.rand(len(sizes), device="cuda") * deg.float()).long().cpu().numpy()
    assert (start != ptr[:-1]).any()
    perm = np.arange(len(src))
    perm[ptr[:-1]], perm[start] = start, ptr[:-1].copy()
    want = perm[O.fps(src[perm], ptr, 0.25)]
    assert np.array_equal(out, want)
    m = np.ceil(np.asarray(sizes, np.float32) * np.float32(0.25)).astype(np.int64)
    assert np.array_equal(out[np.concatenate([[0], np.cumsum(m)[:-1]])], start)       # first pick = the drawn row
    out2 = ops.fps(_dev(src), ratio=0.25, random_start=True, ptr=_dev(ptr)).cpu().numpy()
    assert not np.array_equal(out, out2)                                                # the generator moved on


@pytest.mark.parametrize("k", [16, 32, 64])
def test_unordered_knn_table_is_the_same_set(ops, k):
    """P2W_KNN_UNORDERED: every row holds the same neighbours as the ordered table, and its k-th neighbour sits in
    the first or in the last column (the heap kernel leaves its root first; the ordered kernels may ignore the flag)."""
    from pointstowood_b200.synthetic import tls_plot
    cloud, _ = tls_plot(120_000, 9, side=7.0)
    rng = np.random.default_rng(4)
    tid = (cloud[:, 0] > 3.5).astype(np.int64) * 2 + (cloud[:, 1] > 3.5)
    order = np.argsort(tid, kind="stable")
    x = np.ascontiguousarray(cloud[order, :3])
    x[:40] = x[0]                                                 # a cluster of identical points: d = 0 ties
    ptr = np.concatenate([[0], np.cumsum(np.bincount(tid, minlength=4))]).astype(np.int64)
    qsel = np.sort(rng.choice(len(x), 40_000, replace=False))
    y = x[qsel] + (rng.random((len(qsel), 3)) < 0.5) * np.float32(1e-3)      # half of the queries ARE sources
    y = y.astype(np.float32)
    ptr_y = np.searchsorted(qsel, ptr).astype(np.int64)
    dx, dy, dpx, dpy = _dev(x), _dev(y), _dev(ptr), _dev(ptr_y)
    want, d2 = ops.knn_table(dx, dy, k, dpx, dpy, return_d2=True, method="grid")
    got = ops.knn_table(dx, dy, k, dpx, dpy, method="grid", unordered=True)
    assert torch.equal(torch.sort(got, dim=1).values, torch.sort(want, dim=1).values)
    kth = want[:, k - 1]
    assert bool(((got[:, 0] == kth) | (got[:, k - 1] == kth)).all())
    ref = O.knn(x, y[:2000], k, ptr, np.minimum(ptr_y, 2000))
    assert np.array_equal(np.sort(got[:2000].cpu().numpy(), 1), np.sort(ref, 1))
