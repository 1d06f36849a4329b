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
