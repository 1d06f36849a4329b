"""The packing and spatial-vote oracles (oracle/ref_pipeline.py) against tests/golden/predicter.npz, written by EXECUTING the
reference's own src/predicter.py (TestingDataset.__getitem__, PointCloudClassifier.collect_predictions / compute_labels;
imported unmodified through oracle/shim by oracle/make_golden_predicter.py)."""
import os

import numpy as np

from oracle import ref_pipeline


def _fixtures(golden_dir):
    g = np.load(os.path.join(golden_dir, "tiling.npz"))
    p = np.load(os.path.join(golden_dir, "predicter.npz"))
    feat5 = np.concatenate([g["cloud"][:, :3], g["reflectance"][:, None], g["n_z"][:, None]], 1).astype(np.float32)
    return g, p, feat5


def vote_rows(g):
    """The classified rows oracle/make_golden_predicter.py voted on: every tile point, probabilities that vary over the plot (+ seeded noise)."""
    rng = np.random.default_rng(7)
    xyz = g["cloud"][g["members"], :3].astype(np.float64)
    prob = np.clip(0.5 + 0.45 * np.sin(3.0 * xyz[:, 0]) * np.cos(2.0 * xyz[:, 1]) + 0.1 * rng.normal(size=len(xyz)), 0.0, 1.0)
    return np.concatenate([g["cloud"][g["members"], :3].astype(np.float64), (prob >= 0.5)[:, None].astype(np.float64),
                           prob[:, None]], 1)


def test_packing_matches_the_reference_dataset(golden_dir):
    g, p, feat5 = _fixtures(golden_dir)
    tiles = [g["members"][g["ptr"][t]:g["ptr"][t + 1]] for t in p["tiles"]]
    pos, refl, batch, shift, sf = ref_pipeline.pack(feat5, tiles)
    o = 0
    for i, idx in enumerate(tiles):
        n = len(idx)
        # local_shift: the reference's FP32 torch.mean vs the oracle's FP64-accumulated mean (a pinned choice: the FP32
        # result depends on torch's vectorised summation order) -- equal to within an ulp of the coordinates
        assert np.abs(shift[i] - p[f"shift{i}"]).max() <= 2e-6
        assert np.abs(pos[o:o + n] - p[f"pos{i}"]).max() <= 4e-6
        assert np.array_equal(refl[o:o + n], p[f"refl{i}"])
        assert abs(float(sf[i]) - float(p[f"sf{i}"])) <= 4e-6
        assert (batch[o:o + n] == i).all()
        o += n


def test_spatial_vote_matches_the_reference_classifier(golden_dir):
    g, p, _ = _fixtures(golden_dir)
    rows = vote_rows(g)
    xyz = g["cloud"][:20000, :3].astype(np.float64)
    for name, any_wood in (("vote", 1), ("vote_any", 0.9)):
        label, pwood = ref_pipeline.collect_predictions(rows, xyz, any_wood)
        assert np.array_equal(label, p[f"{name}_label"]), name
        assert np.array_equal(pwood, p[f"{name}_pwood"]), name
    assert 0.2 < p["vote_label"].mean() < 0.8 and p["vote_any_label"].mean() > p["vote_label"].mean()
