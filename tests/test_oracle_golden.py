"""Pins the oracle's model restatement (oracle/ref_model.py) to fixtures produced by the
reference's OWN model code (oracle/make_golden.py ran /root/reference/pointstowood/src/
model.py + pointnet.py unmodified through oracle/shim)."""
import os

import numpy as np
import pytest
import torch

from oracle import oracle as O
from oracle import ref_model


@pytest.mark.parametrize("name", ["net_a", "net_b"])
def test_net_forward_matches_reference_fixture(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    sd = ref_model.seeded_state_dict()
    trace = {}
    logits = ref_model.net_forward(sd, torch.from_numpy(g["pos"]), torch.from_numpy(g["reflectance"]),
                                   torch.from_numpy(g["batch"].astype(np.int64)), torch.from_numpy(g["sf"]), trace)
    for lvl in (1, 2, 3):
        t = trace[f"sa{lvl}_module"]
        assert np.array_equal(t["idx"], g[f"idx{lvl}"]), f"voxel representatives differ at level {lvl}"
        assert np.array_equal(O.table_to_edges(t["nbr"]), g[f"edges{lvl}"]), f"edges differ at level {lvl}"
    # fp32 tolerance on logits: the restatement folds nothing, only reorders Conv1d(k=1) as Linear
    assert np.abs(logits.numpy() - g["logits"]).max() < 2e-5


@pytest.mark.parametrize("tag", ["sa1", "sa2", "sa3"])
def test_pointnet_conv_matches_reference_fixture(golden_dir, tag):
    g = np.load(os.path.join(golden_dir, "conv.npz"))
    sd = {k[len(tag) + 1:]: torch.from_numpy(g[k]) for k in g.files if k.startswith(tag + ".") and k[len(tag) + 1].isdigit()}
    sd = {"local_nn." + k: v for k, v in sd.items()}
    pos = torch.from_numpy(g[tag + ".pos"])
    idx = torch.from_numpy(g[tag + ".idx"].astype(np.int64))
    out = ref_model.pointnet_conv(sd, "", torch.from_numpy(g[tag + ".x"]), pos, pos[idx], g[tag + ".nbr"].astype(np.int64))
    assert np.abs(out.numpy() - g[tag + ".out"]).max() < 1e-5
