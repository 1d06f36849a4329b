"""One pass of the inference path on a synthetic plot (for ncu captures at the bench's real shapes)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointstowood_b200 import model as M  # noqa: E402
from pointstowood_b200.predicter import classify_tiles  # noqa: E402
from pointstowood_b200.preprocessing import Voxelise  # noqa: E402
from pointstowood_b200.synthetic import tls_plot  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
cloud, _ = tls_plot(n, 1)
dev = torch.from_numpy(cloud).cuda()
torch.manual_seed(141190)
net = M.randomise_bn_(M.Net(num_classes=1), 5).cuda().eval().set_precision("bf16")
for _ in range(reps):
    store = Voxelise(dev, minpoints=128, maxpoints=16384, gridsize=(2.0, 4.0)).write_voxels()
    classify_tiles(net, store, 8, 0.5)
torch.cuda.synchronize()
