"""GPU parity of the spatial vote (src/predicter.py:107-142) against the float64 oracle
(oracle/ref_pipeline.collect_predictions) and of the plot-wide cell-list search behind it."""
import numpy as np
import pytest
import torch

from oracle import oracle as O
from oracle import ref_pipeline

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from pointstowood_b200 import ops as _ops
    return _ops


def _classified(n, seed):
    """A plot, every point 'classified' twice with jittered probabilities (as the 2 m and 4 m tiles do)."""
    from pointstowood_b200.synthetic import tls_plot
    cloud, label = tls_plot(n, seed, side=8.0)
    rng = np.random.default_rng(seed)
    xyz = np.concatenate([cloud[:, :3], cloud[:, :3]]).astype(np.float64)
    xyz += rng.normal(0, 1e-7, xyz.shape)                     # un-shifted float64 sums are not fp32 exact
    prob = np.clip(np.concatenate([label, label]) * 0.6 + rng.random(2 * n) * 0.4, 0, 1).astype(np.float32)
    pred = (prob >= 0.5).astype(np.float64)
    perm = rng.permutation(2 * n)
    rows = np.concatenate([xyz, pred[:, None], prob.astype(np.float64)[:, None]], axis=1)[perm]
    return cloud, rows


def test_plotwide_knn_with_cell_hint_is_exact(ops):
    """One 'tile' of 100 k points, k = 64, caller-given cell size: bit-exact against the C oracle."""
    cloud, rows = _classified(50_000, 5)
    x = rows[:, :3].astype(np.float32)
    y = cloud[:3000, :3].astype(np.float32)
    px, py = np.array([0, len(x)]), np.array([0, len(y)])
    ref, ref_d = O.knn(x, y, 64, px, py, return_d2=True)
    nbr, d2 = ops.knn_table(torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda(), 64, torch.from_numpy(px).cuda(),
                            torch.from_numpy(py).cuda(), return_d2=True, method="grid", cell_size=0.05)
    assert np.array_equal(nbr.cpu().numpy().astype(np.int64), ref)
    assert np.array_equal(d2.cpu().numpy(), ref_d)


@pytest.mark.parametrize("any_wood", [1, 0.5])
def test_spatial_vote_matches_float64_oracle(ops, any_wood):
    from pointstowood_b200.predicter import PointCloudClassifier
    cloud, rows = _classified(60_000, 9)
    want_label, want_pwood = ref_pipeline.collect_predictions(rows, cloud[:, :3], any_wood)
    clf = PointCloudClassifier(0.5, any_wood)
    label, pwood = clf.collect_predictions(torch.from_numpy(rows).cuda(), torch.from_numpy(cloud).cuda())
    label, pwood = label.cpu().numpy(), pwood.cpu().numpy()
    # FP32 distances can swap the 64th / 65th neighbour at a near-tie: the vote must agree on >= 99.9 %
    assert (label == want_label).mean() >= 0.999
    assert np.mean(np.abs(pwood - want_pwood) <= 1e-6) >= 0.995
    assert np.abs(pwood - want_pwood).max() <= 0.05


def test_spatial_vote_kernel_exact_on_given_neighbours(ops):
    """compute_labels alone (the neighbour table given): median and votes are exact."""
    rng = np.random.default_rng(3)
    m, n, k = 5000, 2000, 64
    prob = rng.random(m).astype(np.float32)
    prob[:50] = 0.25                                          # ties inside the median
    pred = (prob >= 0.5).astype(np.uint8)
    nbr = np.stack([rng.choice(m, k, replace=False) for _ in range(n)]).astype(np.int32)
    from pointstowood_b200 import _lib
    label = torch.empty(n, dtype=torch.uint8, device="cuda")
    pwood = torch.empty(n, dtype=torch.float64, device="cuda")
    d = lambda a: torch.from_numpy(a).cuda()
    nb_d, pr_d, pd_d = d(nbr), d(prob), d(pred)
    _lib.check(_lib.lib().p2w_spatial_vote(nb_d.data_ptr(), n, k, pr_d.data_ptr(), pd_d.data_ptr(), 1.0, label.data_ptr(),
                                           pwood.data_ptr(), torch.cuda.current_stream().cuda_stream))
    p = prob[nbr].astype(np.float64)
    assert np.array_equal(pwood.cpu().numpy(), np.median(p, axis=1))
    w1 = (p * (pred[nbr] == 1)).sum(1)
    w0 = (p * (pred[nbr] == 0)).sum(1)
    clear = np.abs(w1 - w0) > 1e-9
    assert np.array_equal(label.cpu().numpy()[clear], (w1 > w0)[clear].astype(np.uint8))
