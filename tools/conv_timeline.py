"""Timing experiment: clock64 timeline of every role of conv_tc's CTA 0 over its tiles 20..27.  Needs an
instrumented build (P2W_CONV_INSTRUMENT=1 python -m pointstowood_b200.build --force) and P2W_CONV_DEBUG=16.
Usage: python tools/conv_timeline.py [sa1|sa2|sa3]"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointstowood_b200 import _lib, ops  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "sa1"
C, H, Co, ns, nt = {"sa1": (32, 64, 128, 930000, 400000), "sa2": (128, 192, 256, 400000, 211000),
                    "sa3": (256, 384, 512, 211000, 87000)}[which]
g = torch.Generator(device="cuda").manual_seed(1)
xs = torch.randn(ns, C, device="cuda", generator=g).bfloat16()
ps = torch.rand(ns, 4, device="cuda", generator=g)
idx = torch.linspace(0, ns - 1, nt, device="cuda").long()
nbr = (idx[:, None] + torch.randint(-300, 300, (nt, 32), device="cuda", generator=g)).clamp_(0, ns - 1).to(torch.int32)
w1 = torch.randn(H, C + 4, device="cuda", generator=g) * 0.1
w2 = torch.randn(Co, H, device="cuda", generator=g) * 0.1
z = lambda n: torch.zeros(n, device="cuda")
ws = ops.pointnet_conv_ws(C, H, Co, 1, "cuda")
buf = np.zeros(8192, dtype=np.int64)
fn = _lib.lib().p2wdbg_conv_timeline
fn.restype = ctypes.c_int
for _ in range(3):
    ops.pointnet_conv_max(xs, ps, ps[idx], nbr, w1, z(H), w2, z(Co), torch.ones(Co, device="cuda"), z(Co), 1, ws=ws,
                          out_dtype=torch.bfloat16)
    torch.cuda.synchronize()
    n = fn(buf.ctypes.data_as(ctypes.c_void_p), 8192)
ev = np.sort(buf[:n])
clk, it, role, tag = ev >> 16, (ev >> 8) & 255, (ev >> 4) & 15, ev & 15
roles = {1: "mma L1", 2: "mma L2", 3: "gather", 5: "epi1", 6: "epi2"}
tags = {1: {0: "wait msg_full", 1: "go", 2: "acc free", 3: "issued+commit"}, 2: {0: "wait hid_full", 1: "go", 2: "acc free", 3: "issued+commit"},
        3: {0: "wait msg_empty", 1: "go", 2: "arrived msg_full", 3: "rows ready"},
        5: {0: "wait acc1_full", 1: "go", 2: "hid free", 3: "arrived hid_full"}, 6: {0: "wait acc2_full", 1: "go", 2: "arrived acc2_empty"}}
t0 = clk[0]
for i in range(n):
    r = int(role[i])
    print(f"{int(clk[i] - t0):8d}  tile {int(it[i]):3d}  {roles.get(r, r):8s} {tags.get(r, {}).get(int(tag[i]), int(tag[i]))}")
