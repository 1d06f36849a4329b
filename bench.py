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
            threads, with one KD-tree per tile for knn / radius (torch_cluster's CPU back end).  A step is
            the WHOLE pipeline on a bounded sample -- every point of the 5 x 5 m corner of the plot -- and
            the value is those points over that time: nothing extrapolated.
N > 1 (torchrun): ONE plot sharded over the ranks (pointstowood_b200/distributed.py): rank r holds a chunk of
the rows, whole batches of tiles are dealt round to the ranks, the vote runs on x-slabs with a halo; the
result equals the single-GPU one.  `--scaling weak` (default): the plot has N x --points points (per-GPU work
fixed); `--scaling strong`: --points points in total (BASELINE.json configs[3]: --points 100000000 --gpus 8).
Time = max over ranks.
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


def workload_sharded(total_points: int, world: int) -> str:
    return (f"predict ONE {total_points}-point synthetic TLS plot ({total_points // 1_000_000 or 1} blocks of 1 M points, 2500 points/m^2) "
            f"sharded over {world} GPU(s), grid 2/4 m, min_pts 128, max_pts 16384, batch_size 8, seeded weights; "
            "cloud -> tiles -> network -> spatial vote -> (label, pwood) per point")


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
            try:                                  # gone before the next timed region starts (it holds driver locks while it queries)
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
                self.proc.wait()
            self.thread.join(timeout=2)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i] == "Active" for r in self.rows)]
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=reasons, samples=len(sm))


# --------------------------------------------------------------------------------- CPU arm (oracle port)
CPU_SAMPLE_SIDE = 5.0        # metres: the [0, 5) x [0, 5) m corner of the plot, every point in it (1/16 of the 1 M plot)


def cpu_sample(n_points: int, seed: int, side: float = CPU_SAMPLE_SIDE):
    from pointstowood_b200.synthetic import tls_plot
    cloud, _ = tls_plot(n_points, seed)
    return np.ascontiguousarray(cloud[(cloud[:, 0] < side) & (cloud[:, 1] < side)])


def cpu_arm(cloud: np.ndarray):
    """One pass of the oracle restatement of the SAME pipeline over `cloud` on the host cores, nothing skipped and
    nothing extrapolated: tiling, every batch through the network (neighbour searches on one KD-tree per tile,
    single-threaded per call like torch_cluster's CPU back end; dense layers on torch with all threads,
    predict.py:79-84), write-back, spatial vote (KD-tree over the classified rows just produced).
    Returns (seconds, points)."""
    import torch
    from oracle import oracle as O
    from oracle import ref_model, ref_pipeline
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)                       # predict.py:79-84
    O.SEARCH = "kdtree"
    sd = ref_model.seeded_state_dict()
    t0 = time.perf_counter()
    feat5, tiles, _ = ref_pipeline.preprocess(cloud, CFG["grid_size"], CFG["min_pts"], CFG["max_pts"])
    rows = ref_pipeline.classify(sd, feat5, tiles, CFG["batch_size"], CFG["is_wood"])
    ref_pipeline.collect_predictions(rows, cloud[:, :3], 1, workers=cores)
    dt = time.perf_counter() - t0
    O.SEARCH = "brute"
    return dt, len(cloud), len(tiles), len(rows)


def cpu_baseline_dict(value, cores, n_sample, n_tiles, n_rows, secs):
    return dict(value=value, unit="points/s", cores=cores, kind="port",
                sample=f"every point of the [0,{CPU_SAMPLE_SIDE:g}) x [0,{CPU_SAMPLE_SIDE:g}) m corner of the plot ({n_sample} points, "
                       f"{n_tiles} tiles, {n_rows} classified rows) through the whole pipeline in {secs:.1f} s; "
                       "value = those points / that time, nothing extrapolated; oracle port of the reference host + model "
                       "code (the reference itself needs torch_geometric / torch_cluster / torch_scatter, absent here), "
                       "KD-tree per tile for knn / radius, torch CPU for the dense layers")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cloud = cpu_sample(N_POINTS, 1)
    secs = []
    for i in range(args.warmup + args.steps):
        dt, n, n_tiles, n_rows = cpu_arm(cloud)
        if i >= args.warmup:
            secs.append(dt)
    ms = float(np.mean(secs)) * 1e3
    value = len(cloud) / (ms / 1e3)
    cores = os.cpu_count() or 1
    line = dict(impl="reference", metric="points classified/sec", value=value, unit="points/s", n_gpus=args.gpus,
                steps=args.steps, warmup=args.warmup, ms_per_step=ms, higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="f32", data="synthetic",
                config=dict(workload=workload(N_POINTS) if args.gpus == 1 else workload_sharded(N_POINTS * args.gpus, args.gpus),
                            arm="oracle port of the reference pipeline on the host cores; each step = the whole pipeline "
                                f"on a bounded sample of the workload ({len(cloud)} points, see cpu_baseline.sample)"),
                cpu_baseline=cpu_baseline_dict(value, cores, len(cloud), n_tiles, n_rows, ms / 1e3),
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
    ap.add_argument("--points", type=int, default=N_POINTS,
                    help="points per GPU (--scaling weak) or of the whole plot (--scaling strong)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N > 1 always shards ONE plot over the ranks: of N x points (weak) or of points (strong)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--launch-points", type=int, default=1 << 21,
                    help="points of consecutive reference batches that share one launch set")
    ap.add_argument("--halo", type=float, default=1.0, help="metres of classified rows shared across vote slabs")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from pointstowood_b200 import _lib, ops
    from pointstowood_b200 import model as M
    from pointstowood_b200.distributed import Comm, classify_plot
    from pointstowood_b200.predicter import classify_tiles
    from pointstowood_b200.preprocessing import Voxelise
    from pointstowood_b200.synthetic import tls_plot, tls_plot_blocks

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    assert args.warmup >= 3, "timing rules: at least 3 warm-up steps"
    total_points = args.points * world if args.scaling == "weak" else args.points
    peaks = load_peaks()
    L = _lib.lib()

    # rank r holds rows lo..hi of the plot (the whole plot at N = 1).  1 M points: tls_plot(seed 1), configs[1];
    # larger plots: 1 M-point blocks, block b = tls_plot(seed 1 + b) (configs[3])
    lo, hi = rank * total_points // world, (rank + 1) * total_points // world
    if total_points <= N_POINTS:
        cloud_np = tls_plot(total_points, 1)[0][lo:hi]
    else:
        cloud_np = tls_plot_blocks(total_points, 1, rows=(lo, hi))[0]
    n_local = hi - lo
    host = torch.from_numpy(np.ascontiguousarray(cloud_np)).pin_memory()
    dev_cloud = host.cuda()
    bf16 = args.precision == "bf16"
    torch.manual_seed(141190)
    net = M.Net(num_classes=1)
    M.randomise_bn_(net, 5)
    net = net.cuda().eval().set_precision("bf16" if bf16 else "fp32")
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    info = {}

    def step(cloud):
        if world == 1:
            store = Voxelise(cloud, minpoints=CFG["min_pts"], maxpoints=CFG["max_pts"], gridsize=CFG["grid_size"]).write_voxels()
            prob, pred, xyz, _ = classify_tiles(net, store, CFG["batch_size"], CFG["is_wood"],
                                                max_points_per_launch=args.launch_points, want_xyz=True)
            label, pwood = ops.spatial_vote(xyz, prob, pred, cloud[:, :3].contiguous(), 64, 1.0)
            info.update(tile_points=int(store.ptr[-1]), tiles=int(store.num_tiles))
            return label, pwood
        label, pwood, plot = classify_plot(net, cloud, CFG["min_pts"], CFG["max_pts"], CFG["grid_size"], CFG["batch_size"],
                                           CFG["is_wood"], 1, args.launch_points, args.halo, return_plot=True)
        info.update(tile_points=plot.tile_points, tiles=int(plot.num_tiles), traffic=dict(plot.traffic), halo=plot.halo,
                    vote_rounds=plot.vote_rounds)
        return label, pwood

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        label, pwood = step(dev_cloud)
    out_label = torch.empty(n_local, dtype=torch.uint8).pin_memory()
    out_pwood = torch.empty(n_local, dtype=torch.float64).pin_memory()

    # ---- device-resident steps, dominant kernel timed live with CUDA events
    ops.KERNEL_TIMER.reset("p2w_pointnet_conv_max", "p2w_knn")
    ops.PAIR_EVALS.clear()
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
    pair_evals = sum(int(v.item()) for v in ops.PAIR_EVALS.values())
    ops.KERNEL_TIMER.reset()

    # ---- end to end through the host-facing API (pinned host in, host out); one untimed pass first: the fresh device copy of
    # the cloud and the pinned read-back are new allocations / first touches for the caching allocator and the driver
    out_label.copy_(step(host.cuda(non_blocking=True))[0], non_blocking=True)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    e0.record()
    for i in range(args.steps):
        cloud = host.cuda(non_blocking=True)
        label, pwood = step(cloud)
        out_label.copy_(label, non_blocking=True)
        out_pwood.copy_(pwood, non_blocking=True)
        marks[i].record()
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1) / args.steps
    if rank == 0:
        ends = [e0.elapsed_time(m) for m in marks]
        print("bench: host-to-host ms per step: " + " ".join(f"{b - a:.2f}" for a, b in zip([0.0] + ends[:-1], ends)),
              file=sys.stderr, flush=True)

    t = torch.tensor([ms, ms_e2e], device="cuda", dtype=torch.float64)
    tp = torch.tensor([info.get("tile_points", 0), launches], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tp, op=dist.ReduceOp.SUM)
    ms, ms_e2e = t.tolist()
    tile_points, launches = (int(v) for v in tp.tolist())

    if rank == 0:
        traffic = load_traffic()
        clocks = clk.summary()
        at_max = bool(clocks["sm_mhz"] and clocks["sm_max_mhz"] and clocks["sm_mhz"] >= 0.97 * clocks["sm_max_mhz"]
                      and "sw_power_cap" not in clocks["reasons"])
        roof_conv = roof_knn = None
        if kt_conv["launches"]:
            per_launch_ms = kt_conv["ms"] / kt_conv["launches"]
            achieved = kt_conv["work"] / kt_conv["launches"] / per_launch_ms / 1e9              # TFLOP/s
            peak = (peaks["bf16"] if at_max else peaks["bf16_sustained"]) if bf16 else None
            roof_conv = dict(kernel="conv_tc_kernel (fused gather-MLP-max, tcgen05)" if bf16 else
                             "conv_simt_kernel (fused gather-MLP-max, FP32 parity mode)",
                             bound="tensor", achieved=achieved, peak=peak, unit="TFLOP/s",
                             frac=achieved / peak if peak else None,
                             peak_source=f"{peaks['src']} " + ("burst bf16: the sampled SM clock stayed at its maximum with no power cap"
                                                               if at_max else "sustained bf16: the SM clock dropped under the power cap"),
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
            # the bound that actually applies: a distance is 6 FP32 instructions (3 FSUB, FMUL,
            # 2 FFMA); 148 SMs x 128 lanes x the sampled SM clock
            issue_bound = 148 * 128 * (clocks["sm_mhz"] or 1965.0) * 1e6 / 6.0
            roof_knn.update(pair_evals_per_s=pair_evals / (kt_knn["ms"] / 1e3), fp32_issue_bound_pair_evals_per_s=issue_bound,
                            frac_fp32_issue=pair_evals / (kt_knn["ms"] / 1e3) / issue_bound,
                            pair_evals_per_query="all cell-list searches of the step: SA1 radius, SA2 / SA3 k = 32, four k = 2 "
                                                 "interpolation searches, the vote's k = 64")
        roof = roof_conv if bf16 else roof_knn
        config = dict(workload=workload(N_POINTS) if (world == 1 and total_points == N_POINTS) else
                      workload_sharded(total_points, world),
                      tile_points=tile_points, tiles=info.get("tiles"), launch_points=args.launch_points,
                      l2="no flush needed: a step streams > 4 GB of activations (e.g. the [N0, 544] and "
                         "[N0, 512] FP buffers) between two launches of any kernel, 30x the 126 MB L2")
        if world > 1:
            config.update(partition="rows chunked over ranks; whole batches of tiles dealt round to the ranks; vote by x-slabs "
                                    f"with a {info.get('halo', args.halo):g} m halo ({info.get('vote_rounds')} round(s)); results "
                                    "identical to one GPU",
                          collective_bytes_rank0_per_step={k: v for k, v in sorted(info.get("traffic", {}).items())})
        line = dict(metric="points classified/sec", value=total_points / (ms / 1e3), unit="points/s", n_gpus=world,
                    steps=args.steps, warmup=args.warmup, ms_per_step=ms, higher_is_better=True, scaling=args.scaling,
                    vs_baseline=None, dtype="bf16" if bf16 else "f32", data="synthetic", config=config,
                    e2e=dict(value=total_points / (ms_e2e / 1e3), unit="points/s", h2d_bytes_per_step=int(total_points * 16),
                             d2h_bytes_per_step=int(total_points * 9)),
                    gpu_launches=int(launches), clocks=clocks, roofline=roof,
                    roofline_knn=roof_knn if bf16 else roof_conv)
        if world == 1 and not args.no_cpu_baseline:
            sample = cpu_sample(N_POINTS, 1)
            dt, n, n_tiles, n_rows = cpu_arm(sample)
            line["cpu_baseline"] = cpu_baseline_dict(n / dt, os.cpu_count() or 1, n, n_tiles, n_rows, dt)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
