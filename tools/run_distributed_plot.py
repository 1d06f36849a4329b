"""torchrun driver: one synthetic plot classified by all ranks (distributed.classify_plot) and, on rank 0,
also by a single GPU; prints the agreement and the device time (max over ranks)."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointstowood_b200 import model as M  # noqa: E402
from pointstowood_b200.distributed import classify_plot  # noqa: E402
from pointstowood_b200.synthetic import tls_plot  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
side = 20.0 * (n / 1e6) ** 0.5
cloud = torch.from_numpy(tls_plot(n, 1, side=side)[0]).cuda()
torch.manual_seed(141190)
net = M.randomise_bn_(M.Net(num_classes=1), 5).cuda().eval().set_precision("bf16")
for _ in range(2):
    label, pwood = classify_plot(net, cloud)
torch.cuda.synchronize()
dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
label, pwood = classify_plot(net, cloud)
e1.record()
torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1)], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    ref_label, ref_pwood = classify_plot(net, cloud, rank=0, world_size=1)
    agree = (label == ref_label).float().mean().item()
    same_p = ((pwood - ref_pwood).abs() <= 1e-6).float().mean().item()
    print(json.dumps(dict(points=n, world=world, ms=t.item(), points_per_s=n / t.item() * 1e3, label_agreement=agree,
                          pwood_identical=same_p, max_dpwood=(pwood - ref_pwood).abs().max().item())))
dist.destroy_process_group()
