"""CPU, world_size 2, gloo: the N>1 plumbing (shards + one counter all-reduce) gives the single-process counts."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gen_adversarial_b200 import dist as gdist


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, out):
    os.environ.update({"MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port), "RANK": str(rank), "WORLD_SIZE": str(world)})
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    logits = torch.randn(n, 10, generator=g)
    labels = torch.randint(0, 10, (n,), generator=g)
    robust = torch.randn(n, 10, generator=g)
    lo, hi = gdist.shard_bounds(n, rank, world)
    xs, off = gdist.shard(logits, rank, world)
    assert off == lo and xs.shape[0] == hi - lo
    counters = torch.tensor([hi - lo, gdist.count_correct(xs, labels[lo:hi]), gdist.count_correct(robust[lo:hi], labels[lo:hi])],
                            dtype=torch.int64)
    gdist.reduce_counters(counters)
    if rank == 0:
        out.put(counters.tolist())
    dist.destroy_process_group()


def test_sharded_counters_equal_single_process_counts():
    n, world = 37, 2
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, out)) for r in range(world)]
    [p.start() for p in procs]
    [p.join(120) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    got = out.get()
    g = torch.Generator().manual_seed(0)
    logits = torch.randn(n, 10, generator=g)
    labels = torch.randint(0, 10, (n,), generator=g)
    robust = torch.randn(n, 10, generator=g)
    assert got == [n, gdist.count_correct(logits, labels), gdist.count_correct(robust, labels)]


def test_shard_bounds_cover_the_batch_exactly():
    for n in (0, 1, 7, 512, 513):
        for world in (1, 2, 4, 8):
            spans = [gdist.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
