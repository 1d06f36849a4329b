"""bf16 (the benchmarked precision) against fp32 on EVERY tile point of the 1 M-point plot, and against the CPU oracle on
sampled full batches: the distribution of |dp| and the label agreement at --is-wood 0.5 (north_star: 1e-2, 99.9 %).
Seeded random weights (the checkpoint is not shipped)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_model, ref_pipeline  # noqa: E402
from pointstowood_b200 import model as M  # noqa: E402
from pointstowood_b200.predicter import classify_tiles  # noqa: E402
from pointstowood_b200.preprocessing import Voxelise  # noqa: E402
from pointstowood_b200.synthetic import tls_plot  # noqa: E402

cloud, _ = tls_plot(1_000_000, 1)
dev = torch.from_numpy(cloud).cuda()
sd = ref_model.seeded_state_dict()
net = M.Net(num_classes=1)
net.load_state_dict(sd, strict=True)
net = net.cuda().eval()
store = Voxelise(dev, minpoints=128, maxpoints=16384, gridsize=(2.0, 4.0)).write_voxels()
res = {}
for prec in ("fp32", "bf16"):
    net.set_precision(prec)
    for launch in (1 << 21, 1 << 19):
        prob, pred, _, _ = classify_tiles(net, store, 8, 0.5, max_points_per_launch=launch)
        res[(prec, launch)] = (prob.float().cpu().numpy(), pred.cpu().numpy())


def cmp(a, b, name):
    d = np.abs(a[0] - b[0])
    out = dict(pair=name, n=int(d.size), max=float(d.max()), p999=float(np.quantile(d, 0.999)), p99=float(np.quantile(d, 0.99)),
               mean=float(d.mean()), within_1e2=float((d <= 1e-2).mean()), within_1e3=float((d <= 1e-3).mean()),
               labels=float((a[1] == b[1]).mean()))
    print(json.dumps(out), flush=True)


cmp(res[("bf16", 1 << 21)], res[("fp32", 1 << 21)], "bf16 vs fp32, all tile points")
cmp(res[("bf16", 1 << 21)], res[("bf16", 1 << 19)], "bf16: one launch set vs four")
cmp(res[("fp32", 1 << 21)], res[("fp32", 1 << 19)], "fp32: one launch set vs four")
# oracle on 4 sampled full batches
feat = store.feat.cpu().numpy()
members = store.members.cpu().numpy()
tiles = [members[store.ptr[t]:store.ptr[t + 1]] for t in range(store.num_tiles)]
nb = (store.num_tiles + 7) // 8
for b in np.linspace(0, nb - 1, 4).round().astype(int):
    group = tiles[b * 8:(b + 1) * 8]
    ref = ref_pipeline.classify(sd, feat, group, 8, 0.5)
    lo, hi = int(store.ptr[b * 8]), int(store.ptr[min((b + 1) * 8, store.num_tiles)])
    want = (ref[:, 4].astype(np.float32), ref[:, 3].astype(np.uint8))
    for prec in ("fp32", "bf16"):
        got = (res[(prec, 1 << 21)][0][lo:hi], res[(prec, 1 << 21)][1][lo:hi])
        cmp(got, want, f"{prec} vs oracle, batch {int(b)} ({hi - lo} points)")
