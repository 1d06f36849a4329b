"""Known-answer and property tests of the CPU oracle's primitives.  The
reference holds no golden vectors for them (parity unpinned), so these pin the oracle to an
independent NumPy evaluation of the same published semantics."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import oracle as O


def _d2(x, y):
    """fma(dz,dz,fma(dy,dy,dx*dx)) in float32 (exact products in float64, one rounding per step is
    reproduced by rounding after every add; the lattice inputs below make all steps exact)."""
    d = (x[None, :, :] - y[:, None, :]).astype(np.float32)
    acc = (d[..., 0].astype(np.float64) ** 2).astype(np.float32)
    for i in (1, 2):
        acc = (d[..., i].astype(np.float64) ** 2 + acc.astype(np.float64)).astype(np.float32)
    return acc


def test_knn_known_answer_with_ties():
    x = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [-1, 0, 0], [2, 0, 0], [1, 0, 0]], np.float32)
    y = np.array([[0, 0, 0]], np.float32)
    nbr, d2 = O.knn(x, y, 4, return_d2=True)
    assert nbr.tolist() == [[0, 1, 2, 3]]            # distance ties resolved by the lower index
    assert d2.tolist() == [[0, 1, 1, 1]]
    nbr = O.knn(x[:2], y, 4)
    assert nbr.tolist() == [[0, 1, -1, -1]]          # fewer sources than k: -1 padding
    with pytest.raises(RuntimeError):
        O.knn(x, y, 101)


def test_radius_keeps_lowest_indices_not_nearest():
    x = np.array([[0.05, 0, 0], [0.01, 0, 0], [0.5, 0, 0], [0.02, 0, 0], [0.0, 0, 0]], np.float32)
    nbr, cnt = O.radius(x, np.zeros((1, 3), np.float32), 0.08, max_num_neighbors=2)
    assert nbr.tolist() == [[0, 1]] and cnt.tolist() == [2]
    nbr, cnt = O.radius(x, np.zeros((1, 3), np.float32), 0.05, max_num_neighbors=8)     # strict <
    assert nbr[0, :3].tolist() == [1, 3, 4] and cnt.tolist() == [3]


@settings(max_examples=25, deadline=None)
@given(st.integers(1, 40), st.integers(1, 30), st.integers(1, 12), st.integers(0, 2**31 - 1))
def test_knn_matches_lexicographic_sort(nx, ny, k, seed):
    rng = np.random.default_rng(seed)
    x = rng.integers(0, 4, (nx, 3)).astype(np.float32) / 4          # lattice: many exact ties
    y = rng.integers(0, 4, (ny, 3)).astype(np.float32) / 4
    nbr = O.knn(x, y, k)
    d = _d2(x, y)
    order = np.lexsort((np.broadcast_to(np.arange(nx), d.shape), d), axis=1)[:, :k]
    want = np.full((ny, k), -1)
    want[:, : order.shape[1]] = order
    assert np.array_equal(nbr, want)


def test_batched_search_never_crosses_tiles():
    rng = np.random.default_rng(2)
    x = rng.random((300, 3), dtype=np.float32)
    ptr_x = np.array([0, 100, 100, 300])
    ptr_y = np.array([0, 5, 9, 20])
    nbr = O.knn(x, x[:20], 8, ptr_x, ptr_y)
    assert ((nbr[:5] >= 0) & (nbr[:5] < 100)).all()
    assert (nbr[5:9] == -1).all()                                     # empty source tile
    assert ((nbr[9:] >= 100) & (nbr[9:] < 300)).all()


def test_fps_known_answer():
    src = np.array([[0, 0, 0], [1, 0, 0], [10, 0, 0], [5, 0, 0], [10, 0, 0]], np.float32)
    assert O.fps(src, None, 0.6).tolist() == [0, 2, 3]               # farthest first, ties -> lowest index


def test_grid_and_consecutive_cluster_known_answer():
    pos = np.array([[0.0, 0.0], [0.9, 0.0], [1.0, 0.0], [0.0, 2.5], [1.99, 2.9]], np.float32)
    ids = O.grid(pos, np.array([1.0, 1.0], np.float32))
    assert ids.tolist() == [0, 0, 1, 4, 5]                            # 2 cells in x (stride 1), 3 in y (stride 2)
    inv, perm = O.consecutive_cluster(ids)
    assert inv.tolist() == [0, 0, 1, 2, 3] and perm.tolist() == [1, 2, 3, 4]    # highest member index
    ids = O.voxel_grid(pos, 1.0, np.array([0, 0, 0, 1, 1]))
    assert ids[3] - O.voxel_grid(pos, 1.0, np.zeros(5))[3] == 6       # batch is the slowest-varying axis
