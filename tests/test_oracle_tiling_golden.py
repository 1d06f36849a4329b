"""The tiling oracle (oracle/ref_pipeline.py) against tests/golden/tiling.npz, which oracle/make_golden_tiling.py wrote by
EXECUTING the reference's own src/preprocessing.py (gpu_ground, quantile_normalize_reflectance, grid, write_voxels; one
textual device='cuda' -> 'cpu' substitution at load time, third-party calls through oracle/shim).  This pins the tiling
rows of SURVEY.md 8(a) -- n_z, the normalised reflectance, tile membership and order -- to reference-executed code."""
import os

import numpy as np

from oracle import ref_pipeline


def _golden(golden_dir):
    return np.load(os.path.join(golden_dir, "tiling.npz"))


def test_oracle_tiling_matches_the_reference_run(golden_dir):
    g = _golden(golden_dir)
    cloud = g["cloud"]
    feat5, tiles, grids = ref_pipeline.preprocess(cloud, (2.0, 4.0), 128, 10 ** 9)
    assert np.array_equal(feat5[:, :3], cloud[:, :3])
    assert np.array_equal(feat5[:, 4], g["n_z"]), "height above ground differs from the reference's gpu_ground"
    assert np.array_equal(feat5[:, 3], g["reflectance"]), "normalised reflectance differs from the reference's"
    ptr = g["ptr"]
    assert len(tiles) == len(ptr) - 1 == 131
    for t, idx in enumerate(tiles):
        assert np.array_equal(idx, g["members"][ptr[t]:ptr[t + 1]]), f"tile {t}: members or their order differ"
    n2 = int((grids == 2.0).sum())
    assert 0 < n2 < len(tiles) and (grids[:n2] == 2.0).all() and (grids[n2:] == 4.0).all()      # 2 m list, then 4 m list
    rows0 = np.concatenate([cloud[tiles[0], :3], g["reflectance"][tiles[0], None], g["n_z"][tiles[0], None]], 1)
    assert np.array_equal(rows0, g["rows_tile0"])                # what the reference saved as voxel_0.pt


def test_min_pts_filter_and_thinning_cap_apply_on_top(golden_dir):
    """The same cloud with the shipped limits: every reference voxel with >= min_pts members is a tile; thinning only
    shortens the ones above max_pts (here none of the 2 m voxels, so those tiles are the reference's, row for row)."""
    g = _golden(golden_dir)
    _, tiles, grids = ref_pipeline.preprocess(g["cloud"], (2.0, 4.0), 128, 4096)
    ptr = g["ptr"]
    assert len(tiles) == len(ptr) - 1
    for t, idx in enumerate(tiles):
        ref = g["members"][ptr[t]:ptr[t + 1]]
        if len(ref) <= 4096:
            assert np.array_equal(idx, ref)
        else:
            assert len(idx) == 4096 and len(np.unique(idx)) == 4096 and np.isin(idx, ref).all()
