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


def preprocess_point_cloud_data(df):
    """predict.py:36-53: lower-case names, drop earlier results, reflectance as the 4th column (zeros if absent).
    Returns (frame, extra headers, has_reflectance)."""
    df.columns = df.columns.str.lower()
    dropped = ["label", "pwood", "pleaf"]
    df = df.drop(columns=dropped, errors="ignore")
    df = df.rename(columns=lambda c: c.replace("scalar_", "") if "scalar_" in c else c)
    df = df.rename(columns={"refl": "reflectance", "intensity": "reflectance"})
    headers = [h for h in df.columns[3:] if h not in dropped]
    if "reflectance" not in df.columns:
        df["reflectance"] = np.zeros(len(df))
        print("No reflectance detected, column added with zeros.")
    else:
        print("Reflectance detected")
    cols = list(df.columns)
    cols.insert(3, cols.pop(cols.index("reflectance")))
    return df[cols], headers, True


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description=__doc__)
    p.add_argument("--point-cloud", "-p", default=[], nargs="+", type=str, help="list of point cloud files")
    p.add_argument("--odir", type=str, default=".", help="output directory")
    p.add_argument("--batch_size", default=8, type=int, help="tiles per reference batch (fixes the voxel-grid origin)")
    p.add_argument("--num_procs", default=-1, type=int, help="host threads (torch.set_num_threads)")
    p.add_argument("--resolution", type=float, default=0.01, help="kept for compatibility")
    p.add_argument("--grid_size", type=float, nargs="+", default=[2.0, 4.0], help="grid sizes for voxelization")
    p.add_argument("--min_pts", type=int, default=128, help="minimum number of points in a voxel")
    p.add_argument("--max_pts", type=int, default=16384, help="maximum number of points in a voxel")
    p.add_argument("--model", type=str, default="model.pth", help="checkpoint: a path, or a name under <wdir>/model/")
    p.add_argument("--is-wood", default=0.5, type=float, help="probability above which a point is wood")
    p.add_argument("--any-wood", default=1, type=float, help="probability above which ANY neighbour makes a point wood")
    p.add_argument("--output_fmt", default="ply", help="file type of the output")
    p.add_argument("--precision", default="bf16", choices=["bf16", "fp32"],
                   help="bf16: tensor-core PointNetConv (|dp| <= 1e-2); fp32: parity mode (|dp| <= 1e-3)")
    p.add_argument("--wdir", type=str, default=".", help="directory that holds model/ (the reference derives it from the cwd)")
    p.add_argument("--verbose", action="store_true", help="print stuff")
    return p


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


def main(argv=None):
    args = build_parser().parse_args(argv)
    import torch
    torch.set_num_threads(os.cpu_count() if args.num_procs == -1 else args.num_procs)
    if not args.point_cloud:
        raise Exception("no input specified, please specify --point-cloud")
    for f in args.point_cloud:
        if not OP.isfile(f):
            raise FileNotFoundError(f"Point cloud file not found: {f}")
    net = None
    outs = []
    for f in args.point_cloud:
        outs.append(predict_file(args, f, net))
        net = args.net
    return outs


if __name__ == "__main__":
    main()
