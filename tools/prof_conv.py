"""conv_tc timing at the bench's shapes with spatially local neighbour tables."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointstowood_b200 import ops  # noqa: E402


def timeit(fn, iters=5, warm=2):
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


g = torch.Generator(device="cuda").manual_seed(1)
only = os.environ.get("P2W_PROF_LAYERS", "32,128,256").split(",")
for (C, H, Co, ns, nt) in ((32, 64, 128, 930000, 400000), (128, 192, 256, 400000, 211000), (256, 384, 512, 211000, 87000)):
    if str(C) not in only:
        continue
    for dt in (torch.bfloat16,):
        xs = torch.randn(ns, C, device="cuda", generator=g).to(dt)
        ps = torch.rand(ns, 4, device="cuda", generator=g)
        idx = torch.linspace(0, ns - 1, nt, device="cuda").long()
        off = torch.randint(-300, 300, (nt, 32), device="cuda", generator=g)
        nbr = (idx[:, None] + off).clamp_(0, ns - 1).to(torch.int32)
        w1 = torch.randn(H, C + 4, device="cuda", generator=g) * 0.1
        w2 = torch.randn(Co, H, device="cuda", generator=g) * 0.1
        b1, b2 = torch.zeros(H, device="cuda"), torch.zeros(Co, device="cuda")
        sc, sh = torch.ones(Co, device="cuda"), torch.zeros(Co, device="cuda")
        ws = ops.pointnet_conv_ws(C, H, Co, 1, "cuda")
        ops.pointnet_conv_max(xs, ps, ps[idx], nbr, w1, b1, w2, b2, sc, sh, 1, ws=ws)
        ms = timeit(lambda: ops.pointnet_conv_max(xs, ps, ps[idx], nbr, w1, b1, w2, b2, sc, sh, 1, ws=ws, packed=True,
                                                  out_dtype=dt))
        fl = nt * 32 * (2 * (C + 4) * H + 2 * H * Co)
        print(json.dumps(dict(C=C, dtype=str(dt), n_tgt=nt, ms=ms, tflops=fl / ms / 1e9, us_per_tile=ms * 1e3 / (nt / 4 / 148),
                              debug=os.environ.get("P2W_CONV_DEBUG", "0"))))
