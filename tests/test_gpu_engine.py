"""GPU tests of the eval-mode executor (pointstowood_b200/engine.py): super-batches give the results
of the reference's batch-by-batch loop, the folded network matches the module-by-module forward and
the oracle, and the dense-block helper kernels match torch."""
import os

import numpy as np
import pytest
import torch

from oracle import oracle as O
from oracle import ref_model

pytestmark = pytest.mark.gpu

PROB_TOL_FP32 = 1e-3      # BASELINE.json north_star
PROB_TOL_BF16 = 1e-2


@pytest.fixture(scope="module")
def p2w():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import pointstowood_b200.model as model
    from pointstowood_b200 import ops
    return model, ops


def _tiles_from_plot(n_points, seed, side, n_side):
    """A plot cut into n_side x n_side tiles (tile-major rows), mean-shifted per tile like
    TestingDataset.__getitem__ (src/predicter.py:84-86)."""
    from pointstowood_b200.synthetic import tls_plot
    p, _ = tls_plot(n_points, seed, side=side)
    cell = side / n_side
    tid = (np.minimum((p[:, 0] / cell).astype(int), n_side - 1) * n_side +
           np.minimum((p[:, 1] / cell).astype(int), n_side - 1)).astype(np.int64)
    order = np.argsort(tid, kind="stable")
    p, tid = p[order], tid[order]
    uniq, tile = np.unique(tid, return_inverse=True)
    pos = torch.from_numpy(p[:, :3].copy())
    sf = []
    for b in range(len(uniq)):
        m = torch.from_numpy(tile == b)
        pos[m] = pos[m] - pos[m].mean(0)
        sf.append(torch.sqrt((pos[m] ** 2).sum(1)).max())
    return pos, torch.from_numpy((p[:, 3] / 10).astype(np.float32)), torch.from_numpy(tile.astype(np.int64)), \
        torch.stack(sf)


def test_grouped_voxel_sample_equals_batch_by_batch(p2w):
    """Each reference batch keeps its own grid origin inside a super-batch: bit-exact representatives."""
    _, ops = p2w
    pos, _, batch, sf = _tiles_from_plot(60000, 17, 6.0, 4)            # 16 tiles
    T = sf.numel()
    ptr = np.concatenate([[0], np.cumsum(np.bincount(batch.numpy(), minlength=T))]).astype(np.int64)
    groups = [0, 3, 8, 9, 16]                                           # ragged batches of 3, 5, 1, 7 tiles
    want = []
    for g0, g1 in zip(groups[:-1], groups[1:]):
        lo, hi = ptr[g0], ptr[g1]
        ids = O.voxel_grid(pos[lo:hi].numpy(), 0.08, batch[lo:hi].numpy() - g0)
        want.append(O.consecutive_cluster(ids)[1] + lo)
    got = ops.voxel_sample(pos.cuda(), 0.08, batch.cuda(), ptr=torch.from_numpy(ptr).cuda(),
                           group_ptr=torch.tensor(groups).cuda())
    assert np.array_equal(got.cpu().numpy(), np.concatenate(want))
    # a tiny spatial_bits forces the overflow retry path
    got = ops.voxel_sample(pos.cuda(), 0.08, batch.cuda(), ptr=torch.from_numpy(ptr).cuda(),
                           group_ptr=torch.tensor(groups).cuda(), spatial_bits=6)
    assert np.array_equal(got.cpu().numpy(), np.concatenate(want))


def _net(model_mod, precision="fp32"):
    net = model_mod.Net(num_classes=1)
    net.load_state_dict(ref_model.seeded_state_dict(), strict=True)
    return net.cuda().eval().set_precision(precision)


def test_engine_matches_module_forward_and_oracle(p2w):
    model_mod, _ = p2w
    pos, refl, batch, sf = _tiles_from_plot(30000, 23, 4.0, 2)
    data = lambda: model_mod.make_data(pos.cuda(), refl.cuda(), batch.cuda(), sf.cuda())
    net = _net(model_mod)
    with torch.no_grad():
        fast = net(data())
        net.use_engine = False
        slow = net(data())
    want = ref_model.net_forward(ref_model.seeded_state_dict(), pos, refl, batch, sf)
    p_fast, p_slow, p_ref = (torch.sigmoid(t.float().cpu()).numpy() for t in (fast, slow, want))
    assert np.abs(p_fast - p_slow).max() <= 1e-4
    assert np.abs(p_fast - p_ref).max() <= PROB_TOL_FP32
    assert ((p_fast >= 0.5) == (p_ref >= 0.5)).mean() >= 0.999


@pytest.mark.parametrize("precision,tol", [("bf16", PROB_TOL_BF16), ("bf16-conv", PROB_TOL_BF16)])
def test_engine_bf16_within_tolerance_of_oracle(p2w, precision, tol):
    model_mod, _ = p2w
    pos, refl, batch, sf = _tiles_from_plot(30000, 29, 4.0, 2)
    net = _net(model_mod, precision)
    with torch.no_grad():
        got = net(model_mod.make_data(pos.cuda(), refl.cuda(), batch.cuda(), sf.cuda()))
    want = ref_model.net_forward(ref_model.seeded_state_dict(), pos, refl, batch, sf)
    p, p_ref = torch.sigmoid(got.float().cpu()).numpy(), torch.sigmoid(want).numpy()
    assert np.abs(p - p_ref).max() <= tol
    assert ((p >= 0.5) == (p_ref >= 0.5)).mean() >= 0.999


def test_super_batch_is_invariant_to_launch_size(p2w):
    """classify_tiles over many reference batches in ONE launch set == one launch set per batch."""
    model_mod, _ = p2w
    from pointstowood_b200.predicter import classify_tiles
    from pointstowood_b200.preprocessing import Voxelise
    from pointstowood_b200.synthetic import tls_plot
    cloud, _ = tls_plot(150_000, 43, side=8.0)
    store = Voxelise(cloud, minpoints=256, maxpoints=16384, gridsize=(2.0, 4.0)).write_voxels()
    assert store.num_tiles > 24
    net = _net(model_mod)
    one = classify_tiles(net, store, 8, 0.5, max_points_per_launch=1)           # reference batch per launch
    big = classify_tiles(net, store, 8, 0.5, max_points_per_launch=1 << 30)     # everything at once
    assert len(one[3]) > 3 and len(big[3]) == 1
    assert np.abs(one[0].cpu().numpy() - big[0].cpu().numpy()).max() <= 1e-5     # GEMM blocking may differ
    assert (one[1] == big[1]).float().mean().item() >= 0.9999


def test_affine_relu_matches_torch(p2w):
    _, ops = p2w
    g = torch.Generator(device="cuda").manual_seed(3)
    for dtype, tol in ((torch.float32, 1e-6), (torch.bfloat16, 2e-2)):
        x = torch.randn(1237, 512, device="cuda", generator=g).to(dtype)
        s1, t1, s2, t2 = (torch.randn(512, device="cuda", generator=g) for _ in range(4))
        want1 = torch.relu(x.float() * s1 + t1)
        got1 = ops.affine_relu_(x.clone(), s1, t1)
        assert (got1.float() - want1).abs().max().item() <= tol * (1 + want1.abs().max().item())
        want2 = torch.relu(want1.to(dtype).float() * s2 + t2) if dtype == torch.float32 else torch.relu(want1 * s2 + t2)
        got2 = ops.affine_relu_(x.clone(), s1, t1, s2, t2)
        assert (got2.float() - want2).abs().max().item() <= tol * (1 + want2.abs().max().item())


def test_global_max_pool_with_affine_and_bf16_rows(p2w):
    """GlobalSAModule's BatchNorm + global_max_pool in one pass equals the tensor expression, bit for bit."""
    _, ops = p2w
    g = torch.Generator(device="cuda").manual_seed(4)
    ptr = torch.tensor([0, 5, 5, 1300, 1301, 4000], device="cuda")             # an empty and a one-row segment
    batch = torch.repeat_interleave(torch.arange(5, device="cuda"), ptr[1:] - ptr[:-1])
    s, t = torch.randn(512, device="cuda", generator=g), torch.randn(512, device="cuda", generator=g)
    for dtype in (torch.float32, torch.bfloat16):
        x = torch.randn(4000, 512, device="cuda", generator=g).to(dtype)
        want = ops.global_max_pool(x.float() * s + t, batch, ptr=ptr)
        got = ops.global_max_pool(x, batch, ptr=ptr, scale=s, shift=t)
        assert torch.equal(got, want)
        assert torch.equal(ops.global_max_pool(x, batch, ptr=ptr, scale=None, shift=None),
                           ops.global_max_pool(x.float(), batch, ptr=ptr))
    assert (got[1] == 0).all()


def test_add_relu_matches_torch(p2w):
    _, ops = p2w
    g = torch.Generator(device="cuda").manual_seed(6)
    for dtype in (torch.float32, torch.bfloat16):
        a = torch.randn(3001, 136, device="cuda", generator=g).to(dtype)
        b = torch.randn(3001, 136, device="cuda", generator=g).to(dtype)
        want = torch.relu(a.clone().add_(b))
        got = ops.add_relu_(a, b)
        assert got.data_ptr() == a.data_ptr() and torch.equal(got, want)
    with pytest.raises(RuntimeError):
        ops.add_relu_(torch.zeros(8, 8, device="cuda").t()[:, :4], torch.zeros(8, 4, device="cuda"))


def test_rowdot_matches_torch(p2w):
    """The 1-channel head (conv2, src/model.py:243) as a streaming row dot."""
    _, ops = p2w
    g = torch.Generator(device="cuda").manual_seed(9)
    for c in (8, 32, 64, 128, 200, 512, 1024):
        w = torch.randn(c, device="cuda", generator=g)
        for dtype, n in ((torch.float32, 4099), (torch.bfloat16, 70001), (torch.float32, 0)):
            x = torch.randn(n, c, device="cuda", generator=g).to(dtype)
            got = ops.rowdot(x, w, 0.25)
            want = (x.double() @ w.double() + 0.25).float()
            assert got.shape == (n,) and got.dtype == torch.float32
            if n:
                assert (got - want).abs().max().item() <= 1e-5 * c ** 0.5 * (1 + want.abs().max().item())
    with pytest.raises(RuntimeError):
        ops.rowdot(torch.zeros(4, 12, device="cuda"), torch.zeros(12, device="cuda"), 0.0)


def test_conv_tc_bf16_rows_in_and_out(p2w, golden_dir):
    """The tensor-core PointNetConv with bf16 feature rows in / out agrees with its fp32-row form."""
    _, ops = p2w
    g = np.load(os.path.join(golden_dir, "conv.npz"))
    tag = "sa2"
    sd = {k[len(tag) + 1:]: torch.from_numpy(g[k]).cuda() for k in g.files
          if k.startswith(tag + ".") and k[len(tag) + 1].isdigit()}
    scale = sd["1.2.weight"] / torch.sqrt(sd["1.2.running_var"] + 1e-5)
    shift = sd["1.2.bias"] - sd["1.2.running_mean"] * scale
    w = [sd["0.0.weight"], sd["0.0.bias"], sd["1.0.weight"], sd["1.0.bias"], scale, shift]
    pos = torch.from_numpy(g[tag + ".pos"]).cuda()
    idx = torch.from_numpy(g[tag + ".idx"].astype(np.int64)).cuda()
    x = torch.from_numpy(g[tag + ".x"]).cuda()
    nbr = torch.from_numpy(g[tag + ".nbr"]).cuda()
    ws = ops.pointnet_conv_ws(x.size(1), w[0].size(0), w[2].size(0), ops.CONV_BF16_TC, x.device)
    a = ops.pointnet_conv_max(x, pos, pos[idx], nbr, *w, mode=ops.CONV_BF16_TC, ws=ws)
    b = ops.pointnet_conv_max(x.bfloat16(), pos, pos[idx], nbr, *w, mode=ops.CONV_BF16_TC, ws=ws, packed=True,
                              out_dtype=torch.bfloat16)
    assert b.dtype == torch.bfloat16
    ref = torch.from_numpy(g[tag + ".out"]).cuda()
    scale_ref = ref.abs().max().item()
    assert (a - ref).abs().max().item() <= 3e-2 * scale_ref
    assert (b.float() - a).abs().max().item() <= 1e-2 * scale_ref


def test_knn_interpolate_bf16_rows(p2w):
    _, ops = p2w
    rng = np.random.default_rng(5)
    px = torch.from_numpy(rng.random((500, 3), dtype=np.float32)).cuda()
    py = torch.from_numpy(rng.random((2000, 3), dtype=np.float32)).cuda()
    x = torch.from_numpy(rng.normal(size=(500, 64)).astype(np.float32)).cuda()
    ptr_x, ptr_y = torch.tensor([0, 500]).cuda(), torch.tensor([0, 2000]).cuda()
    want = ops.knn_interpolate(x, px, py, k=2, ptr_x=ptr_x, ptr_y=ptr_y)
    buf = torch.zeros((2000, 96), device="cuda", dtype=torch.bfloat16)
    ops.knn_interpolate(x.bfloat16(), px, py, k=2, ptr_x=ptr_x, ptr_y=ptr_y, out=buf)
    assert (buf[:, :64].float() - want).abs().max().item() <= 3e-2
    assert (buf[:, 64:] == 0).all()


def test_knn_interpolate_cat_equals_interpolate_then_cat(p2w):
    """FPModule front half in one pass: same arithmetic as knn_interpolate + torch.cat (bit-exact in fp32)."""
    _, ops = p2w
    rng = np.random.default_rng(8)
    px = torch.from_numpy(rng.random((700, 3), dtype=np.float32)).cuda()
    py = torch.from_numpy(rng.random((3000, 3), dtype=np.float32)).cuda()
    x = torch.from_numpy(rng.normal(size=(700, 64)).astype(np.float32)).cuda()
    skip = torch.from_numpy(rng.normal(size=(3000, 32)).astype(np.float32)).cuda()
    ptr_x, ptr_y = torch.tensor([0, 300, 700]).cuda(), torch.tensor([0, 1000, 3000]).cuda()
    want = torch.cat([ops.knn_interpolate(x, px, py, k=2, ptr_x=ptr_x, ptr_y=ptr_y), skip], dim=1)
    got = ops.knn_interpolate_cat(x, px, py, skip, 2, ptr_x, ptr_y)
    assert torch.equal(got, want)
    got_bf = ops.knn_interpolate_cat(x.bfloat16(), px, py, skip, 2, ptr_x, ptr_y, out_dtype=torch.bfloat16)
    assert got_bf.dtype == torch.bfloat16 and (got_bf.float() - want).abs().max().item() <= 3e-2
    none = ops.knn_interpolate_cat(x, px, py, None, 2, ptr_x, ptr_y)
    assert torch.equal(none, want[:, :64])


def test_interpolate_add_is_interpolation_after_the_linear():
    """p2w_knn_interpolate_add: relu(interp(x Wc^T) + (x_skip Ws^T + b)) equals relu([interp(x), x_skip] W^T + b) -- the
    identity the engine uses to run an FPModule's first Linear over the coarse rows (src/model.py:149-152)."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from pointstowood_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(11)
    sizes_c, sizes_f = [300, 0, 700], [900, 40, 2100]
    ptr_c = torch.tensor([0, 300, 300, 1000], device="cuda")
    ptr_f = torch.tensor([0, 900, 940, 3040], device="cuda")
    pos_c = torch.rand(1000, 3, device="cuda", generator=g)
    pos_f = torch.rand(3040, 3, device="cuda", generator=g)
    pos_f[:50] = pos_c[:50]                                   # coincident points: the 1e-16 clamp of the weights
    x = torch.randn(1000, 64, device="cuda", generator=g)
    skip = torch.randn(3040, 24, device="cuda", generator=g)
    w = torch.randn(48, 88, device="cuda", generator=g) * 0.2
    b = torch.randn(48, device="cuda", generator=g)
    want = torch.relu(torch.cat([ops.knn_interpolate(x, pos_c, pos_f, k=2, ptr_x=ptr_c, ptr_y=ptr_f), skip], 1) @ w.t() + b)
    y = x @ w[:, :64].t()
    z = skip @ w[:, 64:].t() + b
    got = ops.knn_interpolate_add_(y, pos_c, pos_f, z.clone(), 2, ptr_c, ptr_f, relu=True)
    rows = torch.ones(3040, dtype=torch.bool, device="cuda")
    rows[900:940] = False                                     # tile 1 has no coarse rows: interpolation of nothing is 0 in both
    assert torch.allclose(got[rows], want[rows], atol=1e-4, rtol=1e-4)         # FP32 GEMMs in another association
    assert torch.equal(got[~rows], torch.relu(z[~rows]))
    got16 = ops.knn_interpolate_add_(y.bfloat16(), pos_c, pos_f, z.bfloat16(), 2, ptr_c, ptr_f, relu=True)
    assert (got16.float() - want)[rows].abs().max().item() <= 0.05 * want[rows].abs().max().item()


@pytest.mark.parametrize("n,k,c_out", [(1000, 128, 512), (12345, 256, 1024), (777, 512, 2048), (300, 64, 256), (129, 128, 200),
                                       (0, 128, 512)])
def test_dense_expand_matches_gemm_then_affine(p2w, n, k, c_out):
    """p2w_dense_expand (tcgen05): relu(relu(x W^T + b) * a + c) in one pass equals the library GEMM with its bias + ReLU
    epilogue followed by p2w_affine_relu, to bf16 rounding of the intermediate (which the fused kernel does not round)."""
    _, ops = p2w
    g = torch.Generator(device="cuda").manual_seed(n + k)
    x = torch.randn(n, k, device="cuda", generator=g).bfloat16()
    w = (torch.randn(c_out, k, device="cuda", generator=g) / k ** 0.5).bfloat16().float()
    b = torch.randn(c_out, device="cuda", generator=g) * 0.1
    a = torch.randn(c_out, device="cuda", generator=g)
    c = torch.randn(c_out, device="cuda", generator=g) * 0.1
    ws = ops.dense_expand_ws(k, c_out, x.device)
    got = ops.dense_expand(x, w, b, a, c, ws)
    again = ops.dense_expand(x, w, b, a, c, ws, packed=True)             # the packed weights are re-used
    assert torch.equal(got, again)
    want = torch.relu(torch.relu(x.double() @ w.double().t() + b.double()) * a.double() + c.double())
    assert got.shape == (n, c_out) and got.dtype == torch.bfloat16
    if n:
        err = (got.double() - want).abs()
        assert float((err / (1.0 + want.abs())).max()) < 8e-3             # one bf16 rounding of the result
        plain = ops.dense_expand(x, w, b, None, None)
        want1 = torch.relu(x.double() @ w.double().t() + b.double())
        assert float(((plain.double() - want1).abs() / (1.0 + want1)).max()) < 8e-3
