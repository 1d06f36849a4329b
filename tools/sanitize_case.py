"""Small invocations of the hand-written kernels for compute-sanitizer (memcheck / racecheck / synccheck):
conv_tc_kernel at the three SA widths, the brute-force sweep, the warp-per-query and heap cell-list kernels, the
radix sort + unique pass, pack / write-back, the vote and the fused expand GEMM (dense_tc_kernel).  Results are checked against the oracle / the FP32 kernel so
that a sanitizer-clean run is also a correct one.  Usage: compute-sanitizer --tool memcheck python tools/sanitize_case.py
[dense]  ("dense": only the fused expand case).
(P2W_KNN_HEAP=1 in the environment sends every k >= 5 search through the heap kernel, P2W_KNN_WARP=1 through the warp kernel.)"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as O  # noqa: E402
from pointstowood_b200 import ops  # noqa: E402

rng = np.random.default_rng(0)
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()


def dense_case():
    """dense_tc_kernel: resident (K = 128) and streamed (K = 256, 512) weight rings, ragged last tile, c_out not a multiple of 128"""
    g = torch.Generator(device="cuda").manual_seed(2)
    for n, k, co in ((1000, 128, 512), (700, 256, 1024), (300, 512, 2048), (129, 64, 200)):
        xs = torch.randn(n, k, device="cuda", generator=g).bfloat16()
        w = (torch.randn(co, k, device="cuda", generator=g) / k ** 0.5).bfloat16().float()
        b, a, c = (torch.randn(co, device="cuda", generator=g) * 0.3 for _ in range(3))
        got = ops.dense_expand(xs, w, b, a, c).double()
        want = torch.relu(torch.relu(xs.double() @ w.double().t() + b.double()) * a.double() + c.double())
        assert float(((got - want).abs() / (1.0 + want.abs())).max()) < 8e-3, (n, k, co)
    torch.cuda.synchronize()


if sys.argv[1:] == ["dense"]:
    dense_case()
    print("sanitize_case ok (dense)")
    sys.exit(0)
# ---- neighbour searches (sweep_kernel, grid_query_kernel, grid_query_heap_kernel, grid_query_small_kernel)
sizes = [1500, 0, 700, 2300]
x = rng.random((sum(sizes), 3)).astype(np.float32)
ptr = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
y = x[::3].copy()
bx = np.repeat(np.arange(len(sizes)), sizes)
ptr_y = np.searchsorted(bx[::3], np.arange(len(sizes) + 1)).astype(np.int64)
for k in (2, 16, 32, 64):
    ref = O.knn(x, y, k, ptr, ptr_y)
    for method in ("sweep", "grid"):
        got = ops.knn_table(dev(x), dev(y), k, dev(ptr), dev(ptr_y), method=method).cpu().numpy()
        assert np.array_equal(got, ref), (k, method)
ref, cnt = O.radius(x, y, 0.08, ptr, ptr_y, 32)
for method in ("sweep", "grid"):
    got, c = ops.radius_table(dev(x), dev(y), 0.08, dev(ptr), dev(ptr_y), 32, method=method)
    assert np.array_equal(got.cpu().numpy(), ref) and np.array_equal(c.cpu().numpy(), cnt), method
# ---- voxel sampling (radix sort, unique), fps
batch = dev(bx.astype(np.int64))
idx = ops.voxel_sample(dev(x), 0.1, batch)
assert np.array_equal(idx.cpu().numpy(), O.consecutive_cluster(O.voxel_grid(x, 0.1, bx))[1])
f = ops.fps(dev(x), ratio=0.25, random_start=False, ptr=dev(ptr))
assert np.array_equal(f.cpu().numpy(), O.fps(x, ptr, 0.25))
# ---- fused conv: tcgen05 kernel against the FP32 kernel at the three SA shapes
g = torch.Generator(device="cuda").manual_seed(1)
for (C, H, Co, ns, nt) in ((32, 64, 128, 6000, 1500), (128, 192, 256, 3000, 701), (256, 384, 512, 2000, 333)):
    xs = torch.randn(ns, C, device="cuda", generator=g)
    ps = torch.rand(ns, 4, device="cuda", generator=g)
    tg = torch.randperm(ns, device="cuda", generator=g)[:nt].sort().values
    nbr = torch.randint(0, ns, (nt, 32), device="cuda", generator=g, dtype=torch.int32)
    nbr[::7, 20:] = -1
    nbr[5] = -1
    w1 = torch.randn(H, C + 4, device="cuda", generator=g) * 0.1
    w2 = torch.randn(Co, H, device="cuda", generator=g) * 0.1
    b1, b2 = torch.randn(H, device="cuda", generator=g) * 0.1, torch.randn(Co, device="cuda", generator=g) * 0.1
    sc, sh = torch.randn(Co, device="cuda", generator=g), torch.randn(Co, device="cuda", generator=g) * 0.1
    want = ops.pointnet_conv_max(xs, ps, ps[tg], nbr, w1, b1, w2, b2, sc, sh, ops.CONV_FP32)
    got = ops.pointnet_conv_max(xs.bfloat16(), ps, ps, nbr, w1, b1, w2, b2, sc, sh, ops.CONV_BF16_TC, tgt_index=tg)
    err = (got - want).abs().max().item() / max(want.abs().max().item(), 1e-6)
    assert err < 3e-2, (C, err)
# ---- pack / write-back / vote
cloud = dev(np.concatenate([x, rng.normal(size=(len(x), 1)).astype(np.float32)], 1))
pos, refl, b, shift, sf = ops.pack_tiles(cloud, None, dev(ptr))
prob, pred, xyz = ops.writeback(torch.randn(len(x), device="cuda"), pos, dev(ptr), shift, 0.5, want_xyz=True)
label, pwood = ops.spatial_vote(xyz, prob, pred, cloud[:, :3].contiguous(), 64, 1.0)
assert bool(((pwood >= 0) & (pwood <= 1)).all())
dense_case()
torch.cuda.synchronize()
print("sanitize_case ok")
