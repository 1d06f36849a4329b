"""BASELINE.json configs[4]: train.py forward + backward + optimiser step on synthetic labelled tiles
(batch 32 x max_pts 16384 per GPU), data-parallel with one bucketed NCCL gradient all-reduce per step.
    python tools/bench_train.py [--tiles 32] [--steps 5]            (1 GPU)
    torchrun --nproc-per-node N tools/bench_train.py ...            (N GPUs, weak scaling)
Prints one JSON line on rank 0: points/s (whole job), ms per step, all-reduce ms, bytes."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointstowood_b200 import model as M  # noqa: E402
from pointstowood_b200 import trainer as TR  # noqa: E402
from pointstowood_b200.preprocessing import Voxelise  # noqa: E402
from pointstowood_b200.synthetic import tls_plot  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--tiles", type=int, default=32)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--warmup", type=int, default=2)
ap.add_argument("--fp32", action="store_true")
args = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
# labelled tiles cut from a dense synthetic plot (seed 4 + rank), the largest ones (max_pts members)
cloud, label = tls_plot(3_000_000, 4 + rank, side=20.0)
dev = torch.from_numpy(cloud).cuda()
store = Voxelise(dev, minpoints=8192, maxpoints=16384, gridsize=(2.0, 4.0)).write_voxels()     # train.py: --min_pts 8192
order = np.argsort(-store.sizes, kind="stable")[: args.tiles]
data = TR.make_training_batch(dev, torch.from_numpy(label).cuda(), store, sorted(order.tolist()))
torch.manual_seed(141190)
net = TR.freeze_constant_gate(M.Net(num_classes=1).cuda())
crit = TR.Poly1FocalLoss(reduction="mean", gamma=2.0, alpha=None, label_smoothing=0.1)
opt = torch.optim.AdamW([p for p in net.parameters() if p.requires_grad], lr=1e-4, weight_decay=1e-2)
allreduce = TR.GradientAllReduce(list(net.parameters()))
gen = torch.Generator(device="cuda").manual_seed(rank)
for sa in (net.sa1_module, net.sa2_module, net.sa3_module):
    sa.generator = gen
for _ in range(args.warmup):
    out = TR.train_step(net, opt, crit, data, allreduce, autocast_bf16=not args.fp32)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    out = TR.train_step(net, opt, crit, data, allreduce, autocast_bf16=not args.fp32)
e1.record()
torch.cuda.synchronize()
# the exchange alone
a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a0.record()
for _ in range(5):
    allreduce()
a1.record()
torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1) / args.steps, a0.elapsed_time(a1) / 5], device="cuda")
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    n = int(data.pos.size(0))
    nbytes = 4 * sum(p.numel() for p in net.parameters() if p.requires_grad)
    print(json.dumps(dict(op="train_step", n_gpus=world, tiles_per_gpu=args.tiles, points_per_gpu=n, ms_per_step=t[0].item(),
                          points_per_s=world * n / t[0].item() * 1e3, allreduce_ms=t[1].item(), allreduce_bytes=nbytes,
                          loss=float(out["loss"]), precision="fp32" if args.fp32 else "bf16 autocast",
                          mem_gb=torch.cuda.max_memory_allocated() / 1e9)))
if world > 1:
    dist.destroy_process_group()
