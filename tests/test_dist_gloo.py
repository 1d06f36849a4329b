"""World-size-2 gloo test (CPU) of the only exchange on the multi-GPU inference path: batches are
dealt to ranks by shard_batches, every rank 'classifies' its share, rank 0 gathers the rows."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pointstowood_b200.predicter import gather_rows, plan_batches, shard_batches
    rng = np.random.default_rng(0)
    ptr = np.concatenate([[0], np.cumsum(rng.integers(10, 200, 37))])
    batches = plan_batches(37, 8)
    mine = shard_batches(batches, ptr, world, rank)
    # stand-in for classify_tiles: row value = global member position
    rows = torch.cat([torch.arange(ptr[batches[b][0]], ptr[batches[b][1]], dtype=torch.float64).view(-1, 1)
                      for b in mine]) if mine else torch.empty((0, 1), dtype=torch.float64)
    got = gather_rows(rows, mine, dst=0)
    if rank == 0:
        pieces = {}
        for ids, r in got:
            o = 0
            for b in ids:
                n = int(ptr[batches[b][1]] - ptr[batches[b][0]])
                pieces[b] = r[o:o + n]
                o += n
        full = torch.cat([pieces[b] for b in range(len(batches))]).view(-1)
        out.put(bool(torch.equal(full, torch.arange(ptr[-1], dtype=torch.float64))))
    else:
        assert got is None
    dist.destroy_process_group()


def test_shard_and_gather_world_size_2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    ok = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok


def _worker_allgather(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pointstowood_b200.distributed import all_gather_rows, shard_contiguous
    from pointstowood_b200.predicter import plan_batches
    rng = np.random.default_rng(1)
    ptr = np.concatenate([[0], np.cumsum(rng.integers(10, 200, 53))])
    batches = plan_batches(53, 8)
    mine = shard_contiguous(batches, ptr, world, rank)
    rows = torch.cat([torch.arange(ptr[batches[b][0]], ptr[batches[b][1]], dtype=torch.float32).view(-1, 1).repeat(1, 5)
                      for b in mine]) if mine else torch.empty((0, 5), dtype=torch.float32)
    everything = all_gather_rows(rows)                      # every rank: all rows, rank-major = batch order
    ok = bool(torch.equal(everything[:, 0], torch.arange(ptr[-1], dtype=torch.float32)))
    out.put((rank, ok, len(mine)))
    dist.destroy_process_group()


def test_contiguous_shards_and_uneven_all_gather_world_size_2():
    """distributed.classify_plot's exchange: contiguous batch ranges, one all-gather of uneven row blocks."""
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_allgather, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    got = [out.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in got)
    assert sum(n for _, _, n in got) == 7 and min(n for _, _, n in got) >= 2


def test_shard_contiguous_partitions_and_balances():
    from pointstowood_b200.distributed import shard_contiguous
    from pointstowood_b200.predicter import plan_batches
    rng = np.random.default_rng(2)
    sizes = np.concatenate([rng.integers(128, 2000, 900), rng.integers(2000, 16384, 265)])      # 2 m then 4 m tiles
    ptr = np.concatenate([[0], np.cumsum(sizes)])
    batches = plan_batches(len(sizes), 8)
    for world in (1, 2, 4, 8):
        shards = [shard_contiguous(batches, ptr, world, r) for r in range(world)]
        assert sum(shards, []) == list(range(len(batches)))                                     # a partition, in order
        load = [sum(int(ptr[batches[b][1]] - ptr[batches[b][0]]) for b in s) for s in shards]
        assert max(load) <= 1.15 * (ptr[-1] / world) + 8 * 16384


def _worker_grads(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pointstowood_b200.trainer import GradientAllReduce
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.randn(n)) for n in (1000, 37, 5000, 1)]
    params[1].requires_grad_(False)                                     # frozen parameters stay out of the exchange
    for i, p in enumerate(params):
        if p.requires_grad:
            p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
    GradientAllReduce(params, bucket_bytes=8000)()                      # several buckets
    mean = (1 + 2) / 2
    ok = all(torch.allclose(p.grad, torch.full_like(p, mean * (i + 1))) for i, p in enumerate(params) if p.requires_grad)
    out.put((rank, bool(ok and params[1].grad is None)))
    dist.destroy_process_group()


def test_gradient_all_reduce_world_size_2():
    """train.py's only collective: bucketed gradient averaging."""
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_grads, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    got = [out.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in got)


def _worker_broadcast(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pointstowood_b200.trainer import broadcast_module_
    torch.manual_seed(100 + rank)                                       # replicas that start DIFFERENT
    net = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.BatchNorm1d(5), torch.nn.Linear(5, 1))
    net[1].running_mean.uniform_(-1, 1)
    net[1].num_batches_tracked.fill_(3 + rank)
    broadcast_module_(net, src=0)
    flat = torch.cat([t.detach().double().reshape(-1) for t in list(net.parameters()) + list(net.buffers())])
    got = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(got, flat)
    out.put((rank, bool(all(torch.equal(g, got[0]) for g in got)), int(net[1].num_batches_tracked)))
    dist.destroy_process_group()


def test_replicas_start_from_rank_0_weights_and_buffers():
    """SemanticTraining's first collective: parameters and BatchNorm buffers of rank 0 on every rank."""
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_broadcast, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    got = [out.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in got) and all(nbt == 3 for _, _, nbt in got)
