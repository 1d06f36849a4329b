"""The dense blocks of the eval engine alone, at the bench's row counts (1 M-point plot: N1 = 870 k, N2 = 427 k, N3 = 107 k rows):
every cuBLASLt GEMM of the three InvertedResidualBlocks with its bias / ReLU epilogue, and the streaming passes between them.
Prints per call: ms, TFLOP/s, the HBM bytes the call must move (read A once, write C once, weights negligible) and GB/s,
and the two lower bounds -- FLOPs / burst bf16 peak and bytes / measured HBM copy bandwidth.  A call that takes about the
SUM of the two bounds is doing its memory phase and its math phase one after the other."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointstowood_b200 import ops  # noqa: E402

PEAK_TF, PEAK_GBS = 1667.1, 6542.7


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


def main():
    g = torch.Generator(device="cuda").manual_seed(0)
    flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)            # > L2: every call starts cold
    for name, n, c in (("SA1 block", 870_000, 128), ("SA2 block", 427_000, 256), ("SA3 block", 107_000, 512)):
        e = 4 * c
        for tag, k, m in (("expand", c, e), ("pointwise", e, e), ("project", e, c)):
            a = torch.randn(n, k, device="cuda", generator=g).bfloat16()
            w = (torch.randn(k, m, device="cuda", generator=g) * 0.05).bfloat16()
            b = torch.zeros(m, device="cuda", dtype=torch.bfloat16)

            def call():
                flush.zero_()
                return torch._addmm_activation(b, a, w, use_gelu=False)
            ms = timed(call) - timed(lambda: flush.zero_())
            flops, nbytes = 2.0 * n * k * m, 2.0 * n * (k + m)
            print(json.dumps(dict(block=name, gemm=tag, rows=n, k=k, n_out=m, ms=round(ms, 4), tflops=round(flops / ms / 1e9, 1),
                                  hbm_gbs=round(nbytes / ms / 1e6, 1), bound_math_ms=round(flops / PEAK_TF / 1e9, 4),
                                  bound_hbm_ms=round(nbytes / PEAK_GBS / 1e6, 4))), flush=True)
        # the fused expand (csrc/dense_tc.cu): GEMM + ReLU + affine + ReLU, [n, 4c] written once
        a = torch.randn(n, c, device="cuda", generator=g).bfloat16()
        w32 = torch.randn(e, c, device="cuda", generator=g) * 0.05
        b32, a32, c32 = (torch.randn(e, device="cuda", generator=g) * 0.1 for _ in range(3))
        ws = ops.dense_expand_ws(c, e, a.device)
        ops.dense_expand(a, w32, b32, a32, c32, ws)

        def fused():
            flush.zero_()
            return ops.dense_expand(a, w32, b32, a32, c32, ws, packed=True)
        ms = timed(fused) - timed(lambda: flush.zero_())
        nbytes = 2.0 * n * (c + e)
        print(json.dumps(dict(block=name, op="fused expand + affine (tcgen05)", rows=n, k=c, n_out=e, ms=round(ms, 4),
                              tflops=round(2.0 * n * c * e / ms / 1e9, 1), hbm_gbs=round(nbytes / ms / 1e6, 1),
                              bound_hbm_ms=round(nbytes / PEAK_GBS / 1e6, 4))), flush=True)
        x = torch.randn(n, e, device="cuda", generator=g).bfloat16()
        s1 = torch.rand(e, device="cuda", generator=g) + 0.5
        t1 = torch.randn(e, device="cuda", generator=g) * 0.1

        def ew():
            flush.zero_()
            return ops.affine_relu_(x, s1, t1)
        ms = timed(ew) - timed(lambda: flush.zero_())
        print(json.dumps(dict(block=name, op="affine_relu", rows=n, c=e, ms=round(ms, 4), hbm_gbs=round(4.0 * n * e / ms / 1e6, 1),
                              bound_hbm_ms=round(4.0 * n * e / PEAK_GBS / 1e6, 4))), flush=True)


if __name__ == "__main__":
    main()
