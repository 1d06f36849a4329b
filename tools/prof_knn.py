"""Small driver for ncu captures of the neighbour-search kernels (config 3 shapes)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointstowood_b200 import ops  # noqa: E402
from pointstowood_b200.synthetic import tls_plot, uniform_tiles  # noqa: E402

method = sys.argv[1] if len(sys.argv) > 1 else "grid"
pos, ptr = uniform_tiles(64, 16384, 2.0, 3)
x, p = torch.from_numpy(pos).cuda(), torch.from_numpy(ptr).cuda()
cloud, _ = tls_plot(16 * 16384, 7, side=8.0)
tid = np.minimum((cloud[:, 0] / 2.0).astype(int), 3) * 4 + np.minimum((cloud[:, 1] / 2.0).astype(int), 3)
order = np.argsort(tid, kind="stable")
xt = torch.from_numpy(np.ascontiguousarray(cloud[order, :3])).cuda()
pt = torch.from_numpy(np.concatenate([[0], np.cumsum(np.bincount(tid, minlength=16))]).astype(np.int64)).cuda()
for _ in range(2):
    ops.knn_table(x, x, 32, p, p, method=method)
    ops.knn_table(x, x, 2, p, p, method=method)
    ops.knn_table(xt, xt, 32, pt, pt, method=method)
    ops.knn_table(xt, xt, 2, pt, pt, method=method)
    ops.radius_table(xt, xt, 0.08, pt, pt, 32, method=method)
torch.cuda.synchronize()
