"""Generate tests/golden/predicter.npz by running the REFERENCE's own src/predicter.py code, imported unmodified through
oracle/shim: TestingDataset.__getitem__ (tile packing, :78-94) on voxel files written by the reference's write_voxels, and
PointCloudClassifier.collect_predictions / compute_labels (the spatial vote, :107-142; numba-compiled in the reference) on a
seeded set of classified rows, for both vote rules (--any-wood 1 and != 1).

Run in the build container only (the GPU box has no /root/reference):   python oracle/make_golden_predicter.py
"""
import glob
import os
import sys
import tempfile

import numpy as np
import pandas as pd
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "shim"))
sys.path.insert(0, "/root/reference/pointstowood")

import src.predicter as refpred  # noqa: E402  (the reference, unmodified)


def main():
    g = np.load(os.path.join(ROOT, "tests", "golden", "tiling.npz"))
    cloud, refl, n_z, members, ptr = g["cloud"], g["reflectance"], g["n_z"], g["members"], g["ptr"]
    feat5 = np.concatenate([cloud[:, :3], refl[:, None], n_z[:, None]], 1).astype(np.float32)
    out = {}
    tiles = [3, 40, 90, 130]                                    # two 2 m and two 4 m tiles of the tiling fixture
    with tempfile.TemporaryDirectory() as tmp:
        for i, t in enumerate(tiles):
            torch.save(torch.from_numpy(feat5[members[ptr[t]:ptr[t + 1]]]), os.path.join(tmp, f"voxel_{i}.pt"))
        ds = refpred.TestingDataset(voxels=tmp, max_pts=16384, device="cpu")
        assert len(ds) == len(tiles)
        for i in range(len(tiles)):
            d = ds[i]
            out[f"pos{i}"] = d.pos.numpy()
            out[f"refl{i}"] = d.reflectance.numpy()
            out[f"shift{i}"] = d.local_shift.numpy()
            out[f"sf{i}"] = np.float32(d.sf.item())
    out["tiles"] = np.asarray(tiles)
    # the vote: classified rows = every tile point with a seeded probability; both rules
    rng = np.random.default_rng(7)
    rows_idx = members
    xyz = cloud[rows_idx, :3].astype(np.float64)
    prob = np.clip(0.5 + 0.45 * np.sin(3.0 * xyz[:, 0]) * np.cos(2.0 * xyz[:, 1]) + 0.1 * rng.normal(size=len(rows_idx)), 0.0, 1.0)
    rows = np.concatenate([cloud[rows_idx, :3].astype(np.float64), (prob >= 0.5)[:, None].astype(np.float64), prob[:, None]], 1)
    original = pd.DataFrame(cloud[:20000].astype(np.float64), columns=["x", "y", "z", "reflectance"])
    for name, any_wood in (("vote", 1), ("vote_any", 0.9)):
        clf = refpred.PointCloudClassifier(is_wood=0.5, any_wood=any_wood)
        res = clf.collect_predictions(rows, original.copy())
        out[f"{name}_label"] = res["label"].to_numpy()
        out[f"{name}_pwood"] = res["pwood"].to_numpy()
    # (the rows are rebuilt from tiling.npz and the seed in the tests: tests/test_oracle_predicter_golden.py)
    path = os.path.join(ROOT, "tests", "golden", "predicter.npz")
    np.savez_compressed(path, **out)
    print(path, {k: np.shape(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
