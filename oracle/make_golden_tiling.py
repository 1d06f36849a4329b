"""Generate tests/golden/tiling.npz by running the REFERENCE's own tiling code (src/preprocessing.py: gpu_ground,
quantile_normalize_reflectance, grid and write_voxels) on a seeded cloud.

The reference hard-codes device='cuda' (:43-44, 83, 86) and imports torch_geometric / torch_scatter; neither exists in the
build container.  So the module is loaded from its source text with ONE textual substitution, device='cuda' ->
device='cpu' (the file under /root/reference is not touched), on top of oracle/shim (voxel_grid, consecutive_cluster,
scatter_min forwarded to the CPU oracle).  Everything else -- bucketize edges, the unique / scatter_min ground cells, the
sort-based ranks and erfinv, the per-voxel member lists and their order, the row filter and torch.save of write_voxels --
is the reference's code, executed.  max_pts is set above every voxel size so that the random thinning (:116-120) does not run.

Run in the build container only (the GPU box has no /root/reference):   python oracle/make_golden_tiling.py
"""
import glob
import os
import sys
import tempfile
import types

import numpy as np
import pandas as pd
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "shim"))

from pointstowood_b200.synthetic import tls_plot  # noqa: E402

SRC = "/root/reference/pointstowood/src/preprocessing.py"


def load_reference_preprocessing():
    text = open(SRC).read()
    assert text.count("device='cuda'") >= 3
    mod = types.ModuleType("ref_preprocessing")
    mod.__file__ = SRC
    exec(compile(text.replace("device='cuda'", "device='cpu'"), SRC, "exec"), mod.__dict__)
    return mod


def main():
    ref = load_reference_preprocessing()
    cloud, _ = tls_plot(60_000, 21, side=6.5)
    rng = np.random.default_rng(0)
    while len(np.unique(cloud[:, 3])) < len(cloud):          # equal reflectances would make torch.sort's tie order matter
        _, first = np.unique(cloud[:, 3], return_index=True)
        dup = np.setdiff1d(np.arange(len(cloud)), first)
        cloud[dup, 3] += rng.normal(0, 1e-3, len(dup)).astype(np.float32)
    df = pd.DataFrame(cloud, columns=["x", "y", "z", "reflectance"])
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        # step by step, to record the intermediate columns
        v = ref.Voxelise(df.copy(), vxpath=tmp, minpoints=128, maxpoints=10 ** 9, gridsize=[2.0, 4.0])
        v.pos = torch.tensor(v.pos.values, dtype=torch.float)
        v.pos = v.gpu_ground()
        out["n_z"] = v.pos[:, 4].numpy().copy()
        v.pos[:, 3] = v.quantile_normalize_reflectance()
        out["reflectance"] = v.pos[:, 3].numpy().copy()
        voxels = v.grid()
        out["members"] = np.concatenate([t.numpy() for t in voxels]).astype(np.int64)
        out["ptr"] = np.concatenate([[0], np.cumsum([len(t) for t in voxels])]).astype(np.int64)
        # the whole of write_voxels on a fresh object: the files it saves must hold exactly those rows
        w = ref.Voxelise(df.copy(), vxpath=tmp, minpoints=128, maxpoints=10 ** 9, gridsize=[2.0, 4.0])
        nz_returned = w.write_voxels()
        files = sorted(glob.glob(os.path.join(tmp, "voxel_*.pt")), key=lambda f: int(f.split("_")[-1][:-3]))
        assert len(files) == len(voxels)
        rows = [torch.load(f).numpy() for f in files]
        for t, (idx, r) in enumerate(zip(voxels, rows)):
            assert r.shape == (len(idx), 5)
            assert np.array_equal(r[:, :3], cloud[idx.numpy(), :3]) and np.array_equal(r[:, 4], out["n_z"][idx.numpy()])
            assert np.array_equal(r[:, 3], out["reflectance"][idx.numpy()])
        assert np.array_equal(nz_returned.numpy(), out["n_z"])
        out["rows_tile0"] = rows[0]
    out["cloud"] = cloud
    out["params"] = np.array([128, 2.0, 4.0])
    path = os.path.join(ROOT, "tests", "golden", "tiling.npz")
    np.savez_compressed(path, **out)
    print(path, {k: v.shape for k, v in out.items()}, "tiles:", len(out["ptr"]) - 1)


if __name__ == "__main__":
    main()
