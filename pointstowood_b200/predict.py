"""predict.py of the reference (pointstowood/predict.py:58-180) on the B200 path: same flags, same column handling,
same outputs (`<input>_ours.<fmt>` next to the input with n_z / label / pwood appended), no voxel files on disk.

    python -m pointstowood_b200.predict --point-cloud plot.ply --model model/global.pth [--is-wood 0.5 ...]
"""
from __future__ import annotations

import argparse
import datetime
import os
import os.path as OP

import numpy as np

from .io import load_file, save_file

__all__ = ["preprocess_point_cloud_data", "build_parser", "predict_file", "main"]


_RESULT_COLUMNS = ("label", "pwood", "pleaf")          # outputs of an earlier run: never inputs
_REFLECTANCE_ALIASES = ("refl", "intensity")


def preprocess_point_cloud_data(df):
    """Column handling of predict.py:36-53: names lower-cased, a `scalar_` prefix (CloudCompare) removed, earlier
    results dropped, the reflectance column (alias `refl` / `intensity`; zeros when the file has none) moved
    to position 3.  Returns (frame, names of the extra columns in file order, True)."""
    names = []
    for col in df.columns:
        low = str(col).lower()
        if low in _RESULT_COLUMNS:
            names.append(None)
            continue
        low = low.replace("scalar_", "")
        names.append("reflectance" if low in _REFLECTANCE_ALIASES else low)
    keep = [i for i, n in enumerate(names) if n is not None]
    df = df.iloc[:, keep].copy()
    df.columns = [names[i] for i in keep]
    extras = [c for c in df.columns[3:] if c not in _RESULT_COLUMNS]
    if "reflectance" in df.columns:
        print("Reflectance detected")
    else:
        df["reflectance"] = np.zeros(len(df))
        print("No reflectance detected, column added with zeros.")
    order = list(df.columns[:3]) + ["reflectance"] + [c for c in df.columns[3:] if c != "reflectance"]
    return df[order], extras, True


# (flag spellings, argparse keywords): the reference's command line (predict.py:62-75) plus --precision / --wdir
_FLAGS = [
    (("--point-cloud", "-p"), dict(default=[], nargs="+", type=str, help="point cloud files (.ply, .pcd, .las)")),
    (("--odir",), dict(type=str, default=".", help="output directory")),
    (("--batch_size",), dict(default=8, type=int, help="tiles per reference batch (it fixes the voxel-grid origin)")),
    (("--num_procs",), dict(default=-1, type=int, help="host threads, -1 = all")),
    (("--resolution",), dict(type=float, default=0.01, help="accepted for compatibility")),
    (("--grid_size",), dict(type=float, nargs="+", default=[2.0, 4.0], help="tile sizes in metres")),
    (("--min_pts",), dict(type=int, default=128, help="smallest tile kept")),
    (("--max_pts",), dict(type=int, default=16384, help="largest tile (bigger ones are thinned)")),
    (("--model",), dict(type=str, default="model.pth", help="checkpoint: a path, or a name under <wdir>/model/")),
    (("--is-wood",), dict(default=0.5, type=float, help="probability from which a point counts as wood")),
    (("--any-wood",), dict(default=1, type=float, help="probability from which ANY neighbour makes a point wood")),
    (("--output_fmt",), dict(default="ply", help="ply, pcd, csv or las")),
    (("--verbose",), dict(action="store_true", help="progress messages")),
    (("--precision",), dict(default="bf16", choices=["bf16", "fp32"],
                            help="bf16: tensor-core PointNetConv (|dp| <= 1e-2); fp32: parity mode (|dp| <= 1e-3)")),
    (("--wdir",), dict(type=str, default=".", help="directory holding model/ (the reference derives it from the cwd)")),
]


def build_parser() -> argparse.ArgumentParser:
    parser = argparse.ArgumentParser(description=__doc__)
    for names, kw in _FLAGS:
        parser.add_argument(*names, **kw)
    return parser


def predict_file(args, point_cloud_file: str, net=None) -> str:
    """One input file through load -> tiles -> network -> vote -> save; returns the output path."""
    import torch
    from . import model as M
    from .predicter import SemanticSegmentation
    from .preprocessing import preprocess
    start = datetime.datetime.now()
    stem = OP.splitext(OP.basename(point_cloud_file))[0]
    out = OP.join(OP.dirname(point_cloud_file), f"{stem}_ours.{args.output_fmt}")          # predict.py:123-125
    args.pc, args.headers = load_file(filename=point_cloud_file, additional_headers=True, verbose=False)
    args.pc, args.headers, args.reflectance = preprocess_point_cloud_data(args.pc)
    if args.verbose:
        print(f"Voxelising to {args.grid_size} grid sizes")
    preprocess(args)                                   # args.tiles (on-device TileStore), args.pc['n_z']
    if net is None:
        net = M.Net(num_classes=1).cuda()
        path = args.model if OP.isfile(args.model) else OP.join(args.wdir, "model", args.model)
        try:
            M.load_model(path, net, torch.device("cuda"))
        except (KeyError, FileNotFoundError):
            raise Exception(f"No model loaded at {path}")
        net = net.eval().set_precision(args.precision)
    args.net = net
    SemanticSegmentation(args)                         # args.pc gains 'label' / 'pwood' (src/predicter.py:226-227)
    headers = list(dict.fromkeys(list(args.headers) + ["n_z", "label", "pwood"]))             # src/predicter.py:233
    save_file(out, args.pc.copy(), additional_fields=headers, verbose=False)
    if args.verbose:
        print(f"{point_cloud_file}: {len(args.pc)} points in {(datetime.datetime.now() - start).total_seconds():.1f} s -> {out}")
    return out


def predict_file_sharded(args, point_cloud_file: str, net=None):
    """The same flow with ONE plot sharded over the ranks of a torchrun launch (BASELINE.json configs[3]): every rank
    reads the file, keeps its contiguous chunk of rows on its GPU, `distributed.classify_plot` tiles / classifies / votes
    with the exchanges of DESIGN.md section 8, the per-point columns are gathered and rank 0 writes the output -- the
    file a single GPU writes.  Returns the output path on rank 0, None elsewhere."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from . import model as M
    from .distributed import Comm, classify_plot
    rank, world = dist.get_rank(), dist.get_world_size()
    stem = OP.splitext(OP.basename(point_cloud_file))[0]
    out = OP.join(OP.dirname(point_cloud_file), f"{stem}_ours.{args.output_fmt}")
    pc, headers = load_file(filename=point_cloud_file, additional_headers=True, verbose=False)
    pc, headers, _ = preprocess_point_cloud_data(pc)
    if pc.shape[1] != 4:
        raise Exception("sharded prediction takes x, y, z, reflectance only: the extra scalar columns of this file would "
                        "take part in the voxel grid (run it on one GPU)")
    n = len(pc)
    lo, hi = rank * n // world, (rank + 1) * n // world
    chunk = torch.from_numpy(np.ascontiguousarray(pc.values[lo:hi], dtype=np.float32)).cuda()
    if net is None:
        net = M.Net(num_classes=1).cuda()
        path = args.model if OP.isfile(args.model) else OP.join(args.wdir, "model", args.model)
        try:
            M.load_model(path, net, torch.device("cuda"))
        except (KeyError, FileNotFoundError):
            raise Exception(f"No model loaded at {path}")
        net = net.eval().set_precision(args.precision)
    args.net = net
    label, pwood, plot = classify_plot(net, chunk, args.min_pts, args.max_pts, args.grid_size, args.batch_size, args.is_wood,
                                       args.any_wood, return_plot=True)
    comm = Comm()
    counts = [(r + 1) * n // world - r * n // world for r in range(world)]
    cols = [comm.all_gather_v(t, counts, "all-gather: results") for t in (plot.n_z.double(), label.double(), pwood)]
    if rank != 0:
        return None
    pc = pc.copy()
    for name, col in zip(("n_z", "label", "pwood"), cols):
        pc[name] = col.cpu().numpy()
    save_file(out, pc, additional_fields=list(dict.fromkeys(list(headers) + ["n_z", "label", "pwood"])), verbose=False)
    return out


def main(argv=None):
    """python -m pointstowood_b200.predict ...  classifies every file on one GPU;
    torchrun --nproc-per-node N -m pointstowood_b200.predict ...  shards every file's plot over N GPUs (same output)."""
    args = build_parser().parse_args(argv)
    import torch
    torch.set_num_threads(os.cpu_count() if args.num_procs == -1 else args.num_procs)
    if not args.point_cloud:
        raise Exception("no input specified, please specify --point-cloud")
    for f in args.point_cloud:
        if not OP.isfile(f):
            raise FileNotFoundError(f"Point cloud file not found: {f}")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        import torch.distributed as dist
        own_group = not dist.is_initialized()
        if own_group:
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")) % max(torch.cuda.device_count(), 1))
            dist.init_process_group(os.environ.get("P2W_DIST_BACKEND", "nccl"))
        net, outs = None, []
        for f in args.point_cloud:
            outs.append(predict_file_sharded(args, f, net))
            net = args.net
        if own_group:
            dist.barrier()
            dist.destroy_process_group()
        return outs
    net = None
    outs = []
    for f in args.point_cloud:
        outs.append(predict_file(args, f, net))
        net = args.net
    return outs


if __name__ == "__main__":
    main()
