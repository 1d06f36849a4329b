"""One plot sharded over ranks (pointstowood_b200/distributed.py) against the single-GPU pipeline: the tiles, the
classified rows and the voted (label, pwood) must be IDENTICAL.  World size 1 runs in process; world size 2 runs
as two processes that share cuda:0 and exchange through gloo (host-staged), so the whole plan -- uneven chunks,
voxel-table merge, member all-to-all, thinning at the owner, slab vote with halo -- is exercised on a one-GPU box.
(tools/run_distributed_plot.py is the NCCL run on real multi-GPU boxes.)"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

KW = dict(min_pts=128, max_pts=4096, grid_size=(2.0, 4.0), batch_size=8)


def _cloud():
    from pointstowood_b200.synthetic import tls_plot
    return tls_plot(150_000, 41, side=8.0)[0]


def _net():
    from pointstowood_b200 import model as M
    torch.manual_seed(141190)
    return M.randomise_bn_(M.Net(num_classes=1), 5).cuda().eval().set_precision("bf16")


def _single(cloud, net):
    from pointstowood_b200 import ops
    from pointstowood_b200.predicter import classify_tiles
    from pointstowood_b200.preprocessing import Voxelise
    dev = torch.from_numpy(cloud).cuda()
    store = Voxelise(dev, minpoints=KW["min_pts"], maxpoints=KW["max_pts"], gridsize=KW["grid_size"]).write_voxels()
    prob, pred, xyz, _ = classify_tiles(net, store, KW["batch_size"], 0.5, want_xyz=True)
    label, pwood = ops.spatial_vote(xyz, prob, pred, dev[:, :3].contiguous(), 64, 1.0)
    return store, prob.cpu().numpy(), label.cpu().numpy(), pwood.cpu().numpy()


def test_world_size_1_equals_plain_pipeline():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from pointstowood_b200.distributed import classify_plot
    cloud, net = _cloud(), _net()
    store, prob, label, pwood = _single(cloud, net)
    assert (np.diff(store.ptr) == KW["max_pts"]).any(), "the fixture must hold an oversized tile"
    got_l, got_p, plot = classify_plot(net, torch.from_numpy(cloud).cuda(), **KW, return_plot=True)
    assert plot.num_tiles == store.num_tiles and plot.tile_points == int(store.ptr[-1])
    assert np.array_equal(got_l.cpu().numpy(), label)
    assert np.array_equal(got_p.cpu().numpy(), pwood)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out, halo):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.cuda.set_device(0)
    from pointstowood_b200.distributed import classify_plot
    cloud = _cloud()
    n = len(cloud)
    cuts = [0, int(0.37 * n), n]
    chunk = torch.from_numpy(cloud[cuts[rank]:cuts[rank + 1]].copy()).cuda()
    label, pwood, plot = classify_plot(_net(), chunk, **KW, halo=halo, return_plot=True)
    out.put(dict(rank=rank, lo=cuts[rank], hi=cuts[rank + 1], label=label.cpu().numpy(), pwood=pwood.cpu().numpy(),
                 rounds=plot.vote_rounds, tiles=plot.num_tiles, own=len(plot.local_tiles),
                 traffic=dict(plot.traffic)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("halo", [0.5, 0.01])
def test_world_size_2_equals_single_gpu(halo):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    cloud, net = _cloud(), _net()
    store, prob, label, pwood = _single(cloud, net)
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out, halo)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted([out.get(timeout=600) for _ in range(2)], key=lambda d: d["rank"])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert all(g["tiles"] == store.num_tiles for g in got) and sum(g["own"] for g in got) == store.num_tiles
    assert min(g["own"] for g in got) >= 8
    for g in got:
        assert np.array_equal(g["label"], label[g["lo"]:g["hi"]])
        assert np.array_equal(g["pwood"], pwood[g["lo"]:g["hi"]])
    if halo < 0.1:
        assert got[0]["rounds"] > 1            # a 1 cm halo fails the bound: widened, still exact
    else:
        assert got[0]["rounds"] == 1


def _predict_worker(rank, world, port, src, ckpt):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK="0",
                      P2W_DIST_BACKEND="gloo")
    from pointstowood_b200 import predict
    predict.main(["--point-cloud", src, "--model", ckpt, "--is-wood", "0.5", "--max_pts", "4096"])


def test_predict_cli_sharded_writes_the_single_gpu_file(tmp_path):
    """`torchrun -m pointstowood_b200.predict` (two ranks sharing cuda:0 over gloo here): the output file equals the one
    a single process writes."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import pandas as pd
    from oracle import ref_model
    from pointstowood_b200 import io as pio
    from pointstowood_b200 import predict
    cloud = _cloud()[:90_000]
    ckpt = str(tmp_path / "seeded.pth")
    torch.save({"model_state_dict": ref_model.seeded_state_dict()}, ckpt)
    outs = {}
    for mode in ("single", "sharded"):
        d = tmp_path / mode
        d.mkdir()
        src = str(d / "plot.ply")
        pio.write_ply(src, pd.DataFrame(cloud, columns=["x", "y", "z", "scalar_Reflectance"]))
        if mode == "single":
            predict.main(["--point-cloud", src, "--model", ckpt, "--is-wood", "0.5", "--max_pts", "4096"])
        else:
            ctx = mp.get_context("spawn")
            port = _free_port()
            procs = [ctx.Process(target=_predict_worker, args=(r, 2, port, src, ckpt)) for r in range(2)]
            for p in procs:
                p.start()
            for p in procs:
                p.join(timeout=600)
                assert p.exitcode == 0
        outs[mode] = pio.read_ply(str(d / "plot_ours.ply"))
    a, b = outs["single"], outs["sharded"]
    assert list(a.columns) == list(b.columns) == ["x", "y", "z", "reflectance", "n_z", "label", "pwood"]
    for col in a.columns:
        assert np.array_equal(a[col].to_numpy(), b[col].to_numpy()), col
