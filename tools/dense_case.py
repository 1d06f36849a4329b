"""One launch set of the fused expand kernel at the bench's SA1 / SA2 / SA3 row counts (for ncu)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointstowood_b200 import ops  # noqa: E402

g = torch.Generator(device="cuda").manual_seed(0)
for n, c in ((870_000, 128), (427_000, 256), (107_000, 512)):
    e = 4 * c
    x = torch.randn(n, c, device="cuda", generator=g).bfloat16()
    w = torch.randn(e, c, device="cuda", generator=g) * 0.05
    b, a, cc = (torch.randn(e, device="cuda", generator=g) * 0.1 for _ in range(3))
    ws = ops.dense_expand_ws(c, e, x.device)
    for i in range(3):
        ops.dense_expand(x, w, b, a, cc, ws, packed=i > 0)
torch.cuda.synchronize()
