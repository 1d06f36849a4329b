"""Warm per-kernel time of ONE bench step (CUPTI through torch.profiler, no replay): kernel time sums, the
share of the step each kernel takes, and how much of the step the GPU sat idle (host-bound gaps, syncs).
Usage: python tools/step_timeline.py [n_points] [top_n]   (not a bench value: the profiler adds host overhead)"""
import collections
import os
import re
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointstowood_b200 import model as M  # noqa: E402
from pointstowood_b200 import ops  # noqa: E402
from pointstowood_b200.predicter import classify_tiles  # noqa: E402
from pointstowood_b200.preprocessing import Voxelise  # noqa: E402
from pointstowood_b200.synthetic import tls_plot  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
top = int(sys.argv[2]) if len(sys.argv) > 2 else 45
cloud, _ = tls_plot(n, 1)
dev = torch.from_numpy(cloud).cuda()
torch.manual_seed(141190)
net = M.randomise_bn_(M.Net(num_classes=1), 5).cuda().eval().set_precision("bf16")


def step():
    store = Voxelise(dev, minpoints=128, maxpoints=16384, gridsize=(2.0, 4.0)).write_voxels()
    prob, pred, xyz, _ = classify_tiles(net, store, 8, 0.5, want_xyz=True)
    return ops.spatial_vote(xyz, prob, pred, dev[:, :3].contiguous(), 64, 1.0)


for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
step()
e1.record()
torch.cuda.synchronize()
print(f"# unprofiled step: {e0.elapsed_time(e1):.2f} ms")

from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
agg = collections.defaultdict(lambda: [0, 0.0])
busy, last_end, first = 0.0, None, None
for e in ev:
    s, t = e.time_range.start, e.time_range.end
    name = e.name
    m = re.search(r"(\w+)(<[^(]*>)?\(", name)
    name = (m.group(1) + (m.group(2) or "")) if m else name
    agg[name[:90]][0] += 1
    agg[name[:90]][1] += t - s
    if first is None:
        first = s
    if last_end is None or s >= last_end:
        busy += t - s
        last_end = t
    elif t > last_end:
        busy += t - last_end
        last_end = t
span = last_end - first
tot = sum(v[1] for v in agg.values())
print(f"# profiled step: span {span / 1e3:.2f} ms, GPU busy {busy / 1e3:.2f} ms ({100 * busy / span:.1f} %), "
      f"{len(ev)} device activities, kernel time sum {tot / 1e3:.2f} ms")
print(f"{'us':>10s} {'share':>7s} {'n':>5s} {'avg us':>9s}  kernel")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{v[1]:10.1f} {100 * v[1] / span:6.1f}% {v[0]:5d} {v[1] / v[0]:9.1f}  {k}")

# idle gaps between consecutive device activities, largest first: what ran before / after
gaps = []
end = None
prev = None
for e in ev:
    s, t = e.time_range.start, e.time_range.end
    if end is not None and s - end > 15:
        gaps.append((s - end, prev.name[:60], e.name[:60], (s - first) / 1e3))
    if end is None or t > end:
        end, prev = t, e
print(f"# {len(gaps)} idle gaps > 15 us, total {sum(g[0] for g in gaps) / 1e3:.2f} ms")
for g in sorted(gaps, reverse=True)[:40]:
    print(f"{g[0]:8.1f} us at {g[3]:7.2f} ms  after {g[1]}  before {g[2]}")
