"""CPU tests of the host-side logic around the kernels (no GPU, no compute through libp2w)."""
import numpy as np
import pytest
import torch

from oracle import oracle as O
from oracle import ref_model, ref_pipeline


def test_ops_fail_loudly_on_cpu_tensors():
    from pointstowood_b200 import ops
    from pointstowood_b200._lib import P2WError
    x = torch.rand(16, 3)
    ptr = torch.tensor([0, 16])
    with pytest.raises(P2WError, match="CUDA"):
        ops.knn_table(x, x, 4, ptr, ptr)
    with pytest.raises(P2WError, match="CUDA"):
        ops.voxel_grid(x, 0.1, torch.zeros(16, dtype=torch.long))
    with pytest.raises(P2WError):
        ops.knn(x, x, 101)


def test_product_never_imports_the_oracle():
    import os
    import re
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pointstowood_b200")
    for name in os.listdir(root):
        if name.endswith(".py"):
            text = open(os.path.join(root, name)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), name


def test_state_dict_schema_matches_reference_keys():
    """Module tree / key names of SURVEY.md Appendix D: the oracle's reference-format state dict
    (strict-loaded into the real reference Net by oracle/make_golden.py) loads strictly."""
    from pointstowood_b200 import model as M
    net = M.Net(num_classes=1)
    res = net.load_state_dict(ref_model.seeded_state_dict(), strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert sum(p.numel() for p in net.parameters()) == 18_158_852


def test_batch_to_ptr_and_edge_table_roundtrip():
    from pointstowood_b200 import ops
    from pointstowood_b200.model import PointNetConv
    batch = torch.tensor([0, 0, 0, 2, 2, 3])
    assert ops.batch_to_ptr(batch, 5).tolist() == [0, 3, 3, 5, 6, 6]
    assert np.array_equal(O.batch_to_ptr(batch.numpy(), 5), [0, 3, 3, 5, 6, 6])
    nbr = np.array([[4, 1, -1, -1], [-1, -1, -1, -1], [0, 2, 3, 5]], dtype=np.int64)
    edges = torch.from_numpy(O.table_to_edges(nbr))                      # row 0 = target, row 1 = source
    table = PointNetConv.edge_index_to_table(torch.stack([edges[1], edges[0]]), 3, k=4)
    assert np.array_equal(table.numpy(), nbr)


def test_plan_and_shard_batches():
    from pointstowood_b200.predicter import plan_batches, shard_batches
    assert plan_batches(17, 8) == [(0, 8), (8, 16), (16, 17)]
    assert plan_batches(0, 8) == []
    rng = np.random.default_rng(1)
    ptr = np.concatenate([[0], np.cumsum(rng.integers(128, 16384, 101))])
    batches = plan_batches(101, 8)
    for world in (1, 2, 4, 8):
        parts = [shard_batches(batches, ptr, world, r) for r in range(world)]
        assert sorted(i for p in parts for i in p) == list(range(len(batches)))
        load = [sum(ptr[batches[i][1]] - ptr[batches[i][0]] for i in p) for p in parts]
        assert max(load) - min(load) <= 8 * 16384          # greedy longest-first is within one batch


def test_oracle_pipeline_invariants():
    from pointstowood_b200.synthetic import tls_plot
    cloud, _ = tls_plot(60_000, 5, side=5.0)
    feat5, tiles, grids = ref_pipeline.preprocess(cloud, (2.0, 4.0), 128, 2048)
    assert feat5.shape == (60_000, 5) and feat5[:, 4].min() == 0.0
    assert abs(feat5[:, 3].min() + 1) < 1e-6 and abs(feat5[:, 3].max() - 1) < 1e-6
    assert all(128 <= len(t) <= 2048 for t in tiles) and any(len(t) == 2048 for t in tiles)
    assert list(grids) == sorted(grids)                                  # 2 m tiles first, then 4 m
    for t in tiles:
        assert len(np.unique(t)) == len(t)
    small = [t for t in tiles if len(t) < 2048]
    assert all(np.all(np.diff(t) > 0) for t in small)                    # ascending index inside a tile
    pos, refl, batch, shift, sf = ref_pipeline.pack(feat5, tiles[:3])
    assert np.abs(pos[batch == 1].mean(0)).max() < 1e-4 and sf.shape == (3,)
    # reflectance ranks are stable: equal inputs keep their order
    r = ref_pipeline.quantile_normalize_reflectance(np.array([3, 1, 1, 2, 1], np.float32))
    assert r[1] < r[2] < r[4] < r[3] < r[0]


def test_plan_launches_groups_consecutive_batches_under_a_point_budget():
    from pointstowood_b200.predicter import plan_batches, plan_launches
    sizes = np.array([100] * 20)
    ptr = np.concatenate([[0], np.cumsum(sizes)])
    batches = plan_batches(20, 8)                       # (0,8) (8,16) (16,20): 800, 800, 400 points
    assert plan_launches(batches, ptr, [0, 1, 2], 1 << 30) == [[0, 1, 2]]
    assert plan_launches(batches, ptr, [0, 1, 2], 1600) == [[0, 1], [2]]
    assert plan_launches(batches, ptr, [0, 1, 2], 1) == [[0], [1], [2]]      # a batch is never split
    assert plan_launches(batches, ptr, [0, 2], 1 << 30) == [[0], [2]]        # only consecutive batches merge
    assert plan_launches(batches, ptr, [], 10) == []


def test_poly1_focal_loss_matches_oracle_restatement():
    import torch
    from oracle import ref_model
    from pointstowood_b200.trainer import Poly1FocalLoss
    g = torch.Generator().manual_seed(0)
    z = torch.randn(4000, generator=g) * 6
    y = (torch.rand(4000, generator=g) > 0.7).float()
    got, gamma = Poly1FocalLoss(reduction="mean", gamma=2.0, alpha=None, label_smoothing=0.1)(z, y)
    assert gamma == 2.0 and torch.equal(got, ref_model.poly1_focal_loss(z, y))
    none, _ = Poly1FocalLoss(alpha=0.25)(z, y)
    assert none.shape == z.shape and (none >= 0).all()
