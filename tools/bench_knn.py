"""Cell-list searches alone, at the shapes the bench step runs them (CUDA events after warm-up): 16 384-point tiles
(BASELINE.json configs[2]), TLS-like tiles, and the plot-wide 64-NN of the spatial vote.  Prints one JSON line per
case with the time, the algorithmic GB/s, the distance evaluations per second and their share of the
FP32 issue bound (6 FP32 instructions per evaluation on 148 x 128 lanes at the SM clock).
P2W_KNN_WARP=1 selects the round-1 warp-per-query kernel (identical results) for an A/B on the same box."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointstowood_b200 import _lib, ops  # noqa: E402
from pointstowood_b200.synthetic import tls_plot, uniform_tiles  # noqa: E402

FP32_PAIR_BOUND = 148 * 128 * 1.965e9 / 6.0


def timed(fn, iters=10, warm=3):
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


def search(x, y, px, py, k, radius=None, cell=0.0, unordered=False):
    """One cell-list search through the C ABI with a caller-held workspace -> (ms, pair evaluations)."""
    L = _lib.lib()
    T = px.numel() - 1
    ws = torch.empty(int(L.p2w_grid_search_ws_bytes(x.size(0), y.size(0), T)), device="cuda", dtype=torch.uint8)
    nbr = torch.empty((y.size(0), k), device="cuda", dtype=torch.int32)
    cnt = torch.empty(y.size(0), device="cuda", dtype=torch.int32)
    st = torch.cuda.current_stream().cuda_stream

    def run():
        if radius is None:
            _lib.check(L.p2w_knn_grid_ex(x.data_ptr(), y.data_ptr(), px.data_ptr(), py.data_ptr(), T, x.size(0), y.size(0), k,
                                         float(cell), 1 if unordered else 0, nbr.data_ptr(), None, ws.data_ptr(), ws.numel(), st))
        else:
            _lib.check(L.p2w_radius_grid(x.data_ptr(), y.data_ptr(), px.data_ptr(), py.data_ptr(), T, x.size(0), y.size(0),
                                         float(radius), k, nbr.data_ptr(), cnt.data_ptr(), ws.data_ptr(), ws.numel(), st))
    ms = timed(run)
    evals = int(ops.grid_search_pair_evals(ws, x.size(0), y.size(0), T).item())
    return ms, evals


def line(name, x, y, px, py, k, **kw):
    ms, evals = search(x, y, px, py, k, **kw)
    byt = 12 * (x.size(0) + y.size(0)) + 16 * y.size(0) * k + 16 * px.numel()
    out = dict(case=name, kernel="warp-per-query (P2W_KNN_WARP=1)" if os.environ.get("P2W_KNN_WARP") == "1" else
               ("thread-per-query heap (P2W_KNN_HEAP=1)" if os.environ.get("P2W_KNN_HEAP") == "1" else
                "default dispatch (heap: radius, and kNN up to k = 32 on tiles of >= 8 192 sources; warp kernel otherwise)"),
               nx=x.size(0), ny=y.size(0), tiles=px.numel() - 1, k=k, ms=round(ms, 4), us_per_tile=round(ms * 1e3 / (px.numel() - 1), 2),
               alg_GBs=round(byt / ms / 1e6, 1), frac_hbm=round(byt / ms / 1e6 / 6542.7, 4))
    if evals:
        out.update(pair_evals=evals, evals_per_query=round(evals / y.size(0), 1), pair_evals_per_s=evals / ms * 1e3,
                   frac_fp32_issue=round(evals / ms * 1e3 / FP32_PAIR_BOUND, 4))
    print(json.dumps(out), flush=True)


def main():
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    only = sys.argv[1] if len(sys.argv) > 1 else ""
    if only == "uniform32":                     # one case for ncu
        pos, ptr = uniform_tiles(64, 16384, 2.0, 3)
        x, p = dev(pos), dev(ptr)
        return line("uniform 16384-point tiles, queries = sources", x, x, p, p, 32)
    if only == "vote":
        plot, _ = tls_plot(1_000_000, 1)
        q = dev(plot[:, :3])
        rows = dev(np.concatenate([plot[:, :3], plot[::-1, :3]]))
        one = lambda n: torch.tensor([0, n], device="cuda", dtype=torch.int64)
        return line("spatial vote, unordered table", rows, q, one(rows.size(0)), one(q.size(0)), 64, cell=0.05, unordered=True)
    for B in (8, 64):
        pos, ptr = uniform_tiles(B, 16384, 2.0, 3)
        x, p = dev(pos), dev(ptr)
        for k in (16, 32):
            line(f"uniform 16384-point tiles, queries = sources", x, x, p, p, k)
        line("uniform 16384-point tiles, radius 0.08 max 32", x, x, p, p, 32, radius=0.08)
    cloud, _ = tls_plot(16 * 16384, 7, side=8.0)
    tid = np.minimum((cloud[:, 0] / 2.0).astype(int), 3) * 4 + np.minimum((cloud[:, 1] / 2.0).astype(int), 3)
    order = np.argsort(tid, kind="stable")
    xt = dev(cloud[order, :3])
    pt = dev(np.concatenate([[0], np.cumsum(np.bincount(tid, minlength=16))]).astype(np.int64))
    line("TLS-like tiles, k = 32", xt, xt, pt, pt, 32)
    line("TLS-like tiles, radius 0.08 max 32", xt, xt, pt, pt, 32, radius=0.08)
    plot, _ = tls_plot(1_000_000, 1)
    q = dev(plot[:, :3])
    rows = dev(np.concatenate([plot[:, :3], plot[::-1, :3]]))          # every point classified twice, as by 2 m + 4 m tiles
    one = lambda n: torch.tensor([0, n], device="cuda", dtype=torch.int64)
    line("spatial vote: 64-NN of 1 M points among 2 M rows", rows, q, one(rows.size(0)), one(q.size(0)), 64, cell=0.05)
    line("spatial vote, unordered table", rows, q, one(rows.size(0)), one(q.size(0)), 64, cell=0.05, unordered=True)


if __name__ == "__main__":
    main()
