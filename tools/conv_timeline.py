"""Timing experiment: clock64 timeline of conv_tc's MMA warp (CTA 0).  Needs an instrumented build
(P2W_CONV_INSTRUMENT=1 python -m pointstowood_b200.build --force) and P2W_CONV_DEBUG including 16."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointstowood_b200 import _lib, ops  # noqa: E402

C, H, Co, ns, nt = (256, 384, 512, 211000, 87000) if len(sys.argv) < 2 or sys.argv[1] == "sa3" else (128, 192, 256, 400000, 211000)
g = torch.Generator(device="cuda").manual_seed(1)
xs = torch.randn(ns, C, device="cuda", generator=g).bfloat16()
ps = torch.rand(ns, 4, device="cuda", generator=g)
idx = torch.linspace(0, ns - 1, nt, device="cuda").long()
nbr = (idx[:, None] + torch.randint(-300, 300, (nt, 32), device="cuda", generator=g)).clamp_(0, ns - 1).to(torch.int32)
w1 = torch.randn(H, C + 4, device="cuda", generator=g) * 0.1
w2 = torch.randn(Co, H, device="cuda", generator=g) * 0.1
z = lambda n: torch.zeros(n, device="cuda")
ws = ops.pointnet_conv_ws(C, H, Co, 1, "cuda")
for _ in range(3):
    ops.pointnet_conv_max(xs, ps, ps[idx], nbr, w1, z(H), w2, z(Co), torch.ones(Co, device="cuda"), z(Co), 1, ws=ws,
                          out_dtype=torch.bfloat16)
torch.cuda.synchronize()
buf = np.zeros(8192, dtype=np.int64)
rc = _lib.lib().p2wdbg_conv_timeline(buf.ctypes.data_as(ctypes.c_void_p), 8192)
n = int(buf[0])
ev = buf[2:2 + n]
tags, clk = ev & 255, ev >> 8
names = {1: "L1 wait b1_full", 2: "L2 wait b2_full", 3: "L1 go", 4: "L2 go", 5: "blk wait acc_empty", 6: "blk go", 7: "blk issued"}
t0 = clk[0]
# print tiles 20..22
tile, start = -1, []
for i in range(n):
    if tags[i] == 1:
        tile += 1
    if 20 <= tile < 23:
        print(f"tile {tile:3d} {names[int(tags[i])]:22s} {int(clk[i] - t0):10d}  (+{int(clk[i] - clk[i - 1]) if i else 0})")
