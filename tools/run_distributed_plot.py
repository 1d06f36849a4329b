"""torchrun driver: ONE synthetic plot sharded over all ranks (distributed.classify_plot).  With --check every
rank also classifies the WHOLE plot alone (the plain single-GPU pipeline) and compares its chunk of the sharded
result with it: labels and pwood must be identical.  Prints one JSON line (rank 0): device time (max over ranks),
points/s, agreement, bytes per collective.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/run_distributed_plot.py 4000000 --check
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointstowood_b200 import model as M  # noqa: E402
from pointstowood_b200 import ops  # noqa: E402
from pointstowood_b200.distributed import classify_plot  # noqa: E402
from pointstowood_b200.predicter import classify_tiles  # noqa: E402
from pointstowood_b200.preprocessing import Voxelise  # noqa: E402
from pointstowood_b200.synthetic import tls_plot_blocks  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("points", type=int, nargs="?", default=4_000_000)
ap.add_argument("--check", action="store_true")
ap.add_argument("--halo", type=float, default=1.0)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--timeline", action="store_true", help="rank 0: CUPTI kernel list of one sharded pass (torch.profiler)")
args = ap.parse_args()

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
n = args.points
lo, hi = rank * n // world, (rank + 1) * n // world
chunk = torch.from_numpy(tls_plot_blocks(n, 1, rows=(lo, hi))[0]).cuda()
torch.manual_seed(141190)
net = M.randomise_bn_(M.Net(num_classes=1), 5).cuda().eval().set_precision("bf16")
for _ in range(2):
    label, pwood, plot = classify_plot(net, chunk, halo=args.halo, return_plot=True)
torch.cuda.synchronize()
dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    label, pwood, plot = classify_plot(net, chunk, halo=args.halo, return_plot=True)
e1.record()
torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1) / args.steps], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MAX)
label, pwood, plot = classify_plot(net, chunk, halo=args.halo, return_plot=True, timing=True)
phases = plot.phase_ms()
res = dict(phases_ms_rank0={k: round(v, 3) for k, v in phases.items()}, points=n, world=world, ms=t.item(), points_per_s=n / t.item() * 1e3, tiles=int(plot.num_tiles), halo=plot.halo,
           vote_rounds=plot.vote_rounds, collective_bytes_rank0=dict(plot.traffic))
if args.check:
    whole = torch.from_numpy(tls_plot_blocks(n, 1)[0]).cuda()
    store = Voxelise(whole, minpoints=128, maxpoints=16384, gridsize=(2.0, 4.0)).write_voxels()
    prob, pred, xyz, _ = classify_tiles(net, store, 8, 0.5, want_xyz=True)
    ref_label, ref_pwood = ops.spatial_vote(xyz, prob, pred, whole[:, :3].contiguous(), 64, 1.0)
    same = torch.tensor([float((label == ref_label[lo:hi]).sum()), float((pwood == ref_pwood[lo:hi]).sum()),
                         float(plot.num_tiles == store.num_tiles)], device="cuda", dtype=torch.float64)
    dist.all_reduce(same, op=dist.ReduceOp.SUM)
    res.update(label_agreement=same[0].item() / n, pwood_identical=same[1].item() / n, same_tile_count=same[2].item() == world)
if args.timeline:
    import collections
    import re
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        classify_plot(net, chunk, halo=args.halo)
        torch.cuda.synchronize()
    if rank == 0:
        ev = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA), key=lambda e: e.time_range.start)
        agg = collections.defaultdict(lambda: [0, 0.0])
        for e in ev:
            m = re.search(r"(\w+)(<[^(]*>)?\(", e.name)
            name = (m.group(1) + (m.group(2) or "")) if m else e.name
            agg[name[:80]][0] += 1
            agg[name[:80]][1] += e.time_range.end - e.time_range.start
        span = ev[-1].time_range.end - ev[0].time_range.start
        print(f"# rank 0: span {span / 1e3:.2f} ms, kernel time sum {sum(v[1] for v in agg.values()) / 1e3:.2f} ms, {len(ev)} device activities")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:70]:
            print(f"{v[1]:10.1f} us {v[0]:5d} x  {k}")
if rank == 0:
    print(json.dumps(res))
dist.destroy_process_group()
