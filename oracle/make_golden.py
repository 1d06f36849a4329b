"""Generate tests/golden/*.npz by running the REFERENCE's own model code, imported
unmodified from /root/reference through oracle/shim, on seeded inputs and weights.

Run in the build container only (the GPU box has no /root/reference):
    python oracle/make_golden.py
The fixtures are committed; tests replay them against oracle/ref_model.py (CPU) and the
CUDA path (GPU).
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "shim"))
sys.path.insert(0, "/root/reference/pointstowood")

from oracle import ref_model  # noqa: E402
from oracle import oracle as O  # noqa: E402
from pointstowood_b200.synthetic import tls_plot  # noqa: E402
import src.model as refmodel  # noqa: E402  (the reference, unmodified)
from src.pointnet import PointNetConv  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def make_batch(n_points, side, seed, n_tiles):
    """Cut `n_tiles` square columns out of a small synthetic plot, mean-shift each
    (predicter.py:84-86) and collate (Appendix A.10)."""
    p, _ = tls_plot(n_points, seed, side=side)
    tiles = []
    edges = np.linspace(0, side, n_tiles + 1)
    for t in range(n_tiles):
        m = (p[:, 0] >= edges[t]) & (p[:, 0] < edges[t + 1])
        tiles.append(p[m])
    pos, refl, batch, sf = [], [], [], []
    for t, tile in enumerate(tiles):
        xyz = torch.from_numpy(tile[:, :3].copy())
        xyz = xyz - torch.mean(xyz, axis=0)
        sf.append(torch.sqrt((xyz ** 2).sum(dim=1)).max())
        pos.append(xyz)
        r = tile[:, 3]
        refl.append(torch.from_numpy(((r - r.min()) / (r.max() - r.min()) * 2 - 1).astype(np.float32)))
        batch.append(torch.full((len(tile),), t, dtype=torch.long))
    return torch.cat(pos), torch.cat(refl), torch.cat(batch), torch.stack(sf)


def run_reference_net(sd, pos, refl, batch, sf):
    net = refmodel.Net(num_classes=1)
    missing = net.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    net.eval()
    rec = {"idx": [], "edges": []}
    orig_cc, orig_knn, orig_radius = refmodel.consecutive_cluster, refmodel.knn, refmodel.radius

    def cc(v):
        out = orig_cc(v)
        rec["idx"].append(out[1].numpy().copy())
        return out

    def knn(*a, **k):
        out = orig_knn(*a, **k)
        rec["edges"].append(out.numpy().copy())
        return out

    def radius(*a, **k):
        out = orig_radius(*a, **k)
        rec["edges"].append(out.numpy().copy())
        return out

    refmodel.consecutive_cluster, refmodel.knn, refmodel.radius = cc, knn, radius
    try:
        data = types.SimpleNamespace(pos=pos.clone(), batch=batch.clone(), reflectance=refl.clone(), sf=sf.clone())
        with torch.no_grad():
            logits = net(data)
    finally:
        refmodel.consecutive_cluster, refmodel.knn, refmodel.radius = orig_cc, orig_knn, orig_radius
    return logits.numpy(), rec


def golden_net():
    sd = ref_model.seeded_state_dict()
    for name, (n, side, seed, tiles) in {"net_a": (6000, 3.0, 11, 2), "net_b": (9000, 2.5, 12, 3)}.items():
        pos, refl, batch, sf = make_batch(n, side, seed, tiles)
        logits, rec = run_reference_net(sd, pos, refl, batch, sf)
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), pos=pos.numpy(), reflectance=refl.numpy(),
                            batch=batch.numpy().astype(np.int32), sf=sf.numpy(), logits=logits,
                            idx1=rec["idx"][0].astype(np.int32), idx2=rec["idx"][1].astype(np.int32),
                            idx3=rec["idx"][2].astype(np.int32),
                            edges1=rec["edges"][0].astype(np.int32), edges2=rec["edges"][1].astype(np.int32),
                            edges3=rec["edges"][2].astype(np.int32))
        print(name, pos.shape, [len(i) for i in rec["idx"]], logits[:4])


def golden_conv():
    """The reference PointNetConv (src/pointnet.py) on seeded inputs at SA1/SA2/SA3 widths."""
    g = torch.Generator().manual_seed(77)
    out = {}
    for tag, (C, H, Co, ns, nt, K) in {"sa1": (32, 64, 128, 700, 200, 32), "sa2": (128, 192, 256, 300, 90, 32),
                                       "sa3": (256, 384, 512, 120, 40, 32)}.items():
        local_nn = refmodel.MLP([C + 4, H, Co])
        refmodel.initialize_weights(local_nn)
        sd = {k: v.clone() for k, v in local_nn.state_dict().items()}
        ref_model.randomise_bn(sd, 9)
        local_nn.load_state_dict(sd)
        conv = PointNetConv(local_nn=local_nn, global_nn=None, add_self_loops=False, radius=0.1).eval()
        conv.local_nn.load_state_dict(sd)       # reset_parameters() re-initialised it
        x = torch.randn(ns, C, generator=g)
        pos = torch.cat([torch.rand(ns, 3, generator=g), torch.randn(ns, 1, generator=g)], 1)
        idx = torch.sort(torch.randperm(ns, generator=g)[:nt]).values
        nbr = O.knn(pos[:, :3].numpy(), pos[idx, :3].numpy(), K)
        # ragged: drop a random tail of every third target, as radius() truncation does
        cnt = np.full(nt, K)
        cnt[::3] = torch.randint(1, K, (len(cnt[::3]),), generator=g).numpy()
        nbr[np.arange(K)[None, :] >= cnt[:, None]] = -1
        edges = torch.from_numpy(O.table_to_edges(nbr))
        with torch.no_grad():
            y = conv(x, (pos, pos[idx]), torch.stack([edges[1], edges[0]]))
        for k, v in sd.items():
            out[f"{tag}.{k}"] = v.numpy()
        out.update({f"{tag}.x": x.numpy(), f"{tag}.pos": pos.numpy(), f"{tag}.idx": idx.numpy().astype(np.int32),
                    f"{tag}.nbr": nbr.astype(np.int32), f"{tag}.out": y.numpy()})
        print(tag, y.shape, float(y.abs().mean()))
    np.savez_compressed(os.path.join(GOLD, "conv.npz"), **out)


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    golden_net()
    golden_conv()
