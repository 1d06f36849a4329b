"""fps at BASELINE.json configs[2]: B x 16 384-point tiles, ratio 0.25; P2W_FPS_CLUSTER=0 / 1 forces the single-CTA / cluster kernel."""
import os, sys, json, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointstowood_b200 import ops
from pointstowood_b200.synthetic import uniform_tiles
from oracle import oracle as O
def timed(fn, iters=3, warm=1):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); s,e=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True); s.record()
    for _ in range(iters): fn()
    e.record(); torch.cuda.synchronize(); return s.elapsed_time(e)/iters
for B in (8, 64):
    pos, ptr = uniform_tiles(B, 16384, 2.0, 3)
    x, p = torch.from_numpy(pos).cuda(), torch.from_numpy(ptr).cuda()
    ms = timed(lambda: ops.fps(x, ratio=0.25, random_start=False, ptr=p))
    print(json.dumps(dict(op="fps", kernel={"0": "single CTA per tile", "1": "cluster of 8 CTAs per tile (DSMEM)"}.get(os.environ.get("P2W_FPS_CLUSTER"), "default (cluster up to 37 tiles)"), tiles=B, ratio=0.25, ms=round(ms,3), ms_per_tile_concurrent=round(ms,3))))
# ragged + long tile exactness
rng=np.random.default_rng(1); sizes=[20000, 5, 0, 3000, 16384]
src=rng.random((sum(sizes),3)).astype(np.float32); pt=np.concatenate([[0],np.cumsum(sizes)]).astype(np.int64)
got=ops.fps(torch.from_numpy(src).cuda(), ratio=0.1, random_start=False, ptr=torch.from_numpy(pt).cuda()).cpu().numpy()
print("exact vs oracle:", bool(np.array_equal(got, O.fps(src, pt, 0.1))))
