"""bench.py -- points classified / second for PointsToWood's inference hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--precision bf16|fp32]

Workload (BASELINE.json configs[1]): the predict.py path on a synthetic 1 M-point TLS plot
(xyz + reflectance), grid_size 2.0 4.0, min_pts 128, max_pts 16384, batch_size 8, seeded random
weights (the checkpoint is not shipped).  One STEP = one complete pass over the plot, from the
in-memory [N,4] cloud to one (label, pwood) per ORIGINAL point: height / reflectance normalisation
and 5-D voxel tiling (K6), then for every batch of 8 tiles packing (K7), the network (voxel
sampling K4, radius / kNN K1-K2, fused PointNetConv K5, kNN interpolation, cuBLASLt dense blocks)
and write-back (K8) -- the `classified_pc` rows of src/predicter.py:217 -- and the spatial vote
over them (src/predicter.py:107-142; 64-NN of every original point among all classified points,
median pwood, class votes).  File I/O (src/io.py) is outside the region on both arms.

`value`   : N_points * steps / device time, cloud resident in HBM when the clock starts.
`e2e`     : the same through the host-facing API: pinned host cloud -> H2D -> pipeline -> D2H of the
            per-point probability / label, copies inside the timed region.
`roofline`: the dominant libp2w kernel of the step, timed live with CUDA events.
`cpu_baseline` / `--impl reference`: the reference's host + model code cannot be imported on the
            GPU box (torch_geometric, torch_cluster, torch_scatter absent), so the CPU arm is the
            oracle's restatement (oracle/ref_pipeline.py + ref_model.py, kind "port") on all host
            threads: full tiling, a bounded sample of the batches extrapolated by tile points to the
            whole plot, and the full spatial vote (scipy cKDTree on all cores) on a plot-sized stand-in
            for the classified rows.
N > 1 (torchrun): every rank classifies its own 1 M-point plot (seed 1 + rank): weak scaling, no
collective on the data path; time = max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_POINTS = 1_000_000
CFG = dict(grid_size=(2.0, 4.0), min_pts=128, max_pts=16384, batch_size=8, is_wood=0.5)


def workload(n_points: int) -> str:
    """The same description on both arms (BASELINE.json configs[1])."""
    return (f"predict {n_points}-point synthetic TLS plot per GPU, grid 2/4 m, min_pts 128, max_pts 16384, batch_size 8, "
            "seeded weights; cloud -> tiles -> network -> spatial vote -> (label, pwood) per point")


def load_traffic():
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the roofline kernels, from the
    committed `ncu --set full` capture of this workload (profiles/traffic.json; null if absent)."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f)
    return {}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        if "hbm_gbs" in p and "bf16_tflops" in p:           # a driver file without the sustained figure: burst for both
            return dict(hbm=p["hbm_gbs"], bf16=p["bf16_tflops"],
                        bf16_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]), src="measured")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms during the timed region."""

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "50"], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i] == "Active" for r in self.rows)]
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=reasons, samples=len(sm))


# --------------------------------------------------------------------------------- CPU arm (oracle port)
def cpu_arm(n_points: int, seed: int, budget_s: float = 20.0):
    """Times the oracle restatement of the same pipeline on the host cores: full tiling, then as
    many batches as fit the budget, extrapolated by tile points to the whole plot."""
    import torch
    from oracle import oracle as O
    from oracle import ref_model, ref_pipeline
    from pointstowood_b200.synthetic import tls_plot
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)                       # predict.py:79-84
    O.lib().orc_set_threads(cores)
    cloud, _ = tls_plot(n_points, seed)
    sd = ref_model.seeded_state_dict()
    t0 = time.perf_counter()
    feat5, tiles, _ = ref_pipeline.preprocess(cloud, CFG["grid_size"], CFG["min_pts"], CFG["max_pts"])
    t_pre = time.perf_counter() - t0
    total_pts = sum(len(t) for t in tiles)
    # sample batches evenly across the tile list (2 m tiles are small, 4 m tiles large)
    nb = (len(tiles) + CFG["batch_size"] - 1) // CFG["batch_size"]
    order = np.linspace(0, nb - 1, num=min(nb, 64)).round().astype(int)
    done_pts, t_cls, used = 0, 0.0, 0
    for b in dict.fromkeys(order.tolist()):
        group = tiles[b * CFG["batch_size"]:(b + 1) * CFG["batch_size"]]
        t1 = time.perf_counter()
        ref_pipeline.classify(sd, feat5, group, CFG["batch_size"], CFG["is_wood"])
        t_cls += time.perf_counter() - t1
        done_pts += sum(len(t) for t in group)
        used += 1
        if t_cls > budget_s:
            break
    # spatial vote on a plot-sized stand-in for the classified rows: every tile point at its own
    # coordinates with a synthetic probability (the KD-tree cost does not depend on the values)
    members = np.concatenate(tiles)
    rng = np.random.default_rng(0)
    prob = rng.random(len(members))
    rows = np.concatenate([feat5[members, :3].astype(np.float64), (prob >= 0.5)[:, None].astype(np.float64),
                           prob[:, None]], axis=1)
    t2 = time.perf_counter()
    ref_pipeline.collect_predictions(rows, cloud[:, :3], 1, workers=cores)
    t_vote = time.perf_counter() - t2
    est = t_pre + t_cls * total_pts / max(done_pts, 1) + t_vote
    return dict(value=n_points / est, unit="points/s", cores=cores, kind="port",
                sample=f"full tiling ({t_pre:.2f} s) + {used} of {nb} batches ({done_pts} of {total_pts} tile points, "
                       f"{t_cls:.1f} s), extrapolated by tile points + full spatial vote ({t_vote:.1f} s)"), est


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals = []
    base = None
    for i in range(args.warmup + args.steps):
        base, est = cpu_arm(N_POINTS, 1, budget_s=12.0)
        if i >= args.warmup:
            vals.append(est)
    ms = float(np.mean(vals)) * 1e3
    value = N_POINTS / (ms / 1e3)
    base["value"] = value
    line = dict(impl="reference", metric="points classified/sec", value=value, unit="points/s", n_gpus=args.gpus,
                steps=args.steps, warmup=args.warmup, ms_per_step=ms, higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="f32", data="synthetic",
                config=dict(workload=workload(N_POINTS),
                            arm="reference pipeline restated on oracle CPU ops (kind: port), bounded sample per step"),
                cpu_baseline=base,
                e2e=dict(value=value, unit="points/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


# --------------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="p2w", choices=["p2w", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--points", type=int, default=N_POINTS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--launch-points", type=int, default=1 << 21,
                    help="points of consecutive reference batches that share one launch set")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from pointstowood_b200 import _lib, ops
    from pointstowood_b200 import model as M
    from pointstowood_b200.predicter import classify_tiles
    from pointstowood_b200.preprocessing import Voxelise
    from pointstowood_b200.synthetic import tls_plot

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    assert args.warmup >= 3, "timing rules: at least 3 warm-up steps"
    n_points = args.points
    peaks = load_peaks()
    L = _lib.lib()

    cloud_np, _ = tls_plot(n_points, 1 + rank)
    host = torch.from_numpy(cloud_np).pin_memory()
    dev_cloud = host.cuda()
    bf16 = args.precision == "bf16"
    torch.manual_seed(141190)
    net = M.Net(num_classes=1)
    M.randomise_bn_(net, 5)
    net = net.cuda().eval().set_precision("bf16" if bf16 else "fp32")
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False

    def step(cloud):
        store = Voxelise(cloud, minpoints=CFG["min_pts"], maxpoints=CFG["max_pts"], gridsize=CFG["grid_size"]).write_voxels()
        prob, pred, xyz, _ = classify_tiles(net, store, CFG["batch_size"], CFG["is_wood"],
                                            max_points_per_launch=args.launch_points, want_xyz=True)
        label, pwood = ops.spatial_vote(xyz, prob, pred, cloud[:, :3].contiguous(), 64, 1.0)
        return store, label, pwood

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        store, label, pwood = step(dev_cloud)
    tile_points = int(store.ptr[-1])
    out_label = torch.empty(n_points, dtype=torch.uint8).pin_memory()
    out_pwood = torch.empty(n_points, dtype=torch.float64).pin_memory()

    # ---- device-resident steps, dominant kernel timed live with CUDA events
    ops.KERNEL_TIMER.reset("p2w_pointnet_conv_max", "p2w_knn")
    barrier()
    launches0 = L.p2w_launch_count()
    with ClockSampler(local) as clk:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(args.steps):
            step(dev_cloud)
        ev1.record()
        barrier()
    launches = L.p2w_launch_count() - launches0
    ms = ev0.elapsed_time(ev1) / args.steps
    kt_conv, kt_knn = ops.KERNEL_TIMER.summary("p2w_pointnet_conv_max"), ops.KERNEL_TIMER.summary("p2w_knn")
    ops.KERNEL_TIMER.reset()

    # ---- end to end through the host-facing API (pinned host in, host out)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        cloud = host.cuda(non_blocking=True)
        _, label, pwood = step(cloud)
        out_label.copy_(label, non_blocking=True)
        out_pwood.copy_(pwood, non_blocking=True)
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1) / args.steps

    t = torch.tensor([ms, ms_e2e], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = t.tolist()

    if rank == 0:
        traffic = load_traffic()
        roof_conv = roof_knn = None
        if kt_conv["launches"]:
            per_launch_ms = kt_conv["ms"] / kt_conv["launches"]
            achieved = kt_conv["work"] / kt_conv["launches"] / per_launch_ms / 1e9              # TFLOP/s
            peak = peaks["bf16_sustained"] if bf16 else None
            roof_conv = dict(kernel="conv_tc_kernel (fused gather-MLP-max, tcgen05)" if bf16 else
                             "conv_simt_kernel (fused gather-MLP-max, FP32 parity mode)",
                             bound="tensor", achieved=achieved, peak=peak, unit="TFLOP/s",
                             frac=achieved / peak if peak else None,
                             peak_source=f"{peaks['src']} sustained bf16 (kernel timed inside a long step)",
                             traffic=traffic.get("conv_tc_kernel") if bf16 else None,
                             work="FLOPs = 32 n_tgt (2 (C+4) H + 2 H C') per launch, unpadded (SURVEY.md 8d)",
                             launches=kt_conv["launches"], avg_launch_ms=per_launch_ms)
        if kt_knn["launches"]:
            per_launch_ms = kt_knn["ms"] / kt_knn["launches"]
            achieved = kt_knn["work"] / kt_knn["launches"] / per_launch_ms / 1e6                # GB/s
            roof_knn = dict(kernel="p2w_knn (cell-list / sweep kNN, build + query)", bound="hbm", achieved=achieved,
                            peak=peaks["hbm"], unit="GB/s", frac=achieved / peaks["hbm"], peak_source=peaks["src"],
                            traffic=traffic.get("grid_query_kernel"),
                            work="bytes = 12 (Nx + Ny) + 16 Ny k + 16 (B+1) per call (SURVEY.md 8d)",
                            launches=kt_knn["launches"], avg_launch_ms=per_launch_ms)
        roof = roof_conv if bf16 else roof_knn
        line = dict(metric="points classified/sec", value=world * n_points / (ms / 1e3), unit="points/s", n_gpus=world,
                    steps=args.steps, warmup=args.warmup, ms_per_step=ms, higher_is_better=True, scaling="weak",
                    vs_baseline=None, dtype="bf16" if bf16 else "f32", data="synthetic",
                    config=dict(workload=workload(n_points),
                                tile_points=tile_points, tiles=int(store.num_tiles),
                                launch_points=args.launch_points,
                                l2="no flush needed: a step streams > 4 GB of activations (e.g. the [N0, 544] and "
                                   "[N0, 512] FP buffers) between two launches of any kernel, 30x the 126 MB L2"),
                    e2e=dict(value=world * n_points / (ms_e2e / 1e3), unit="points/s", h2d_bytes_per_step=int(host.numel() * 4),
                             d2h_bytes_per_step=int(n_points * 9)),
                    gpu_launches=int(launches), clocks=clk.summary(), roofline=roof,
                    roofline_knn=roof_knn if bf16 else roof_conv)
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"], _ = cpu_arm(n_points, 1, budget_s=15.0)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
