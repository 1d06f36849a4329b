"""Per-kernel timings on the GPU box (CUDA events, after warm-up); prints one JSON line each."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointstowood_b200 import ops  # noqa: E402
from pointstowood_b200.synthetic import uniform_tiles, tls_plot  # noqa: E402


def timeit(fn, iters=10, warm=3):
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


def main():
    out = []
    for B in (8, 64):
        pos, ptr = uniform_tiles(B, 16384, 2.0, 3)
        x = torch.from_numpy(pos).cuda()
        p = torch.from_numpy(ptr).cuda()
        for method in ("sweep", "grid"):
            for k in (16, 32):
                ms = timeit(lambda: ops.knn_table(x, x, k, p, p, method=method))
                byt = 12 * 2 * x.size(0) + 16 * x.size(0) * k + 16 * (B + 1)
                out.append(dict(op="knn", method=method, tiles=B, k=k, ms=ms, us_per_tile=ms * 1e3 / B,
                                alg_GBs=byt / ms / 1e6, frac_hbm=byt / ms / 1e6 / 6542.7))
            ms = timeit(lambda: ops.radius_table(x, x, 0.08, p, p, 32, method=method))
            out.append(dict(op="radius", method=method, tiles=B, ms=ms))
            ms = timeit(lambda: ops.knn_table(x, x, 2, p, p, method=method))
            out.append(dict(op="knn", method=method, tiles=B, k=2, ms=ms))
        batch = torch.repeat_interleave(torch.arange(B, device="cuda"), 16384)
        ms = timeit(lambda: ops.voxel_sample(x, 0.04, batch))
        out.append(dict(op="voxel_sample", tiles=B, ms=ms))
    # TLS-like tiles (surfaces): 16 tiles of ~16k points cut from a synthetic plot
    cloud, _ = tls_plot(16 * 16384, 7, side=8.0)
    tid = np.minimum((cloud[:, 0] / 2.0).astype(int), 3) * 4 + np.minimum((cloud[:, 1] / 2.0).astype(int), 3)
    order = np.argsort(tid, kind="stable")
    xt = torch.from_numpy(np.ascontiguousarray(cloud[order, :3])).cuda()
    pt = torch.from_numpy(np.concatenate([[0], np.cumsum(np.bincount(tid, minlength=16))]).astype(np.int64)).cuda()
    for method in ("sweep", "grid"):
        for k in (2, 32):
            ms = timeit(lambda: ops.knn_table(xt, xt, k, pt, pt, method=method))
            byt = 12 * 2 * xt.size(0) + 16 * xt.size(0) * k + 16 * 17
            out.append(dict(op="knn_tls", method=method, tiles=16, k=k, ms=ms, alg_GBs=byt / ms / 1e6,
                            frac_hbm=byt / ms / 1e6 / 6542.7))
        ms = timeit(lambda: ops.radius_table(xt, xt, 0.08, pt, pt, 32, method=method))
        out.append(dict(op="radius_tls", method=method, tiles=16, ms=ms))
    pos, ptr = uniform_tiles(8, 16384, 2.0, 3)
    x = torch.from_numpy(pos).cuda()
    p = torch.from_numpy(ptr).cuda()
    ms = timeit(lambda: ops.fps(x, ratio=0.25, random_start=False, ptr=p), iters=2, warm=1)
    out.append(dict(op="fps", tiles=8, ratio=0.25, ms=ms))
    # fused conv at the three SA widths
    g = torch.Generator(device="cuda").manual_seed(1)
    for (C, H, Co, ns, nt) in ((32, 64, 128, 131072, 40000), (128, 192, 256, 40000, 13000), (256, 384, 512, 13000, 4000)):
        xs = torch.randn(ns, C, device="cuda", generator=g)
        ps = torch.rand(ns, 4, device="cuda", generator=g)
        idx = torch.randperm(ns, device="cuda", generator=g)[:nt].sort().values
        nbr = torch.randint(0, ns, (nt, 32), device="cuda", generator=g, dtype=torch.int32)
        w1 = torch.randn(H, C + 4, device="cuda", generator=g) * 0.1
        w2 = torch.randn(Co, H, device="cuda", generator=g) * 0.1
        b1, b2 = torch.zeros(H, device="cuda"), torch.zeros(Co, device="cuda")
        sc, sh = torch.ones(Co, device="cuda"), torch.zeros(Co, device="cuda")
        for mode in (0, 1):
            try:
                ms = timeit(lambda: ops.pointnet_conv_max(xs, ps, ps[idx], nbr, w1, b1, w2, b2, sc, sh, mode), iters=5)
            except Exception as ex:   # noqa: BLE001
                out.append(dict(op="conv", mode=mode, C=C, error=str(ex)))
                continue
            fl = nt * 32 * (2 * (C + 4) * H + 2 * H * Co)
            out.append(dict(op="conv", mode=mode, C=C, H=H, Co=Co, n_tgt=nt, ms=ms, tflops=fl / ms / 1e9))
    for o in out:
        print(json.dumps(o))


if __name__ == "__main__":
    main()
