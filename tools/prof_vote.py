"""Timing of the spatial vote's pieces on the bench plot (1 M points, ~2 M classified rows)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointstowood_b200 import _lib, ops  # noqa: E402
from pointstowood_b200.preprocessing import Voxelise  # noqa: E402
from pointstowood_b200.synthetic import tls_plot  # noqa: E402


def timeit(fn, iters=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


cloud, _ = tls_plot(1_000_000, 1)
dev = torch.from_numpy(cloud).cuda()
store = Voxelise(dev, minpoints=128, maxpoints=16384, gridsize=(2.0, 4.0)).write_voxels()
xyz = store.feat[store.members, :3].contiguous()
org = dev[:, :3].contiguous()
px = torch.tensor([0, xyz.size(0)], device="cuda")
py = torch.tensor([0, org.size(0)], device="cuda")
prob = torch.rand(xyz.size(0), device="cuda")
pred = (prob >= 0.5).to(torch.uint8)
for k in (64, 32):
    for c in (0.03, 0.05, 0.08, 0.12, 0.2):
        ms = timeit(lambda: ops.knn_table(xyz, org, k, px, py, method="grid", cell_size=c))
        print(json.dumps(dict(op="plot_knn", k=k, cell=c, ms=ms)))
nbr = ops.knn_table(xyz, org, 64, px, py, method="grid", cell_size=0.05)
label = torch.empty(org.size(0), device="cuda", dtype=torch.uint8)
pw = torch.empty(org.size(0), device="cuda", dtype=torch.float64)
ms = timeit(lambda: _lib.check(_lib.lib().p2w_spatial_vote(nbr.data_ptr(), org.size(0), 64, prob.data_ptr(), pred.data_ptr(), 1.0,
                                                           label.data_ptr(), pw.data_ptr(), torch.cuda.current_stream().cuda_stream)))
print(json.dumps(dict(op="vote_kernel", ms=ms)))
