"""CPU, world_size 2, gloo: the host-side logic of the pair-sharded multi-GPU path."""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from glue_factory_colon_b200.shard import gather_to_rank0, pair_cost, shard_bounds  # noqa: E402


def test_shard_bounds_cover_every_pair_once():
    for n, world in ((64, 8), (7, 2), (3, 4), (32, 1), (10, 3)):
        costs = [pair_cost(1024 + 37 * i, 4096 - 29 * i) for i in range(n)]
        b = shard_bounds(costs, world)
        assert len(b) == world and b[0][0] == 0 and b[-1][1] == n
        assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
        if n >= world:
            assert all(e > s for s, e in b)
    even = shard_bounds([1.0] * 64, 8)
    assert all(e - s == 8 for s, e in even)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    costs = [pair_cost(100 + i, 90) for i in range(7)]
    bounds = shard_bounds(costs, world)
    s, e = bounds[rank]
    # stand-in for the per-rank matcher output: matches0 of the rank's pairs
    local = torch.arange(s, e).view(-1, 1).repeat(1, 5)
    out = gather_to_rank0(local, [b[1] - b[0] for b in bounds])
    if rank == 0:
        q.put(out.tolist())
    dist.barrier()
    dist.destroy_process_group()


def test_gather_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert got == [[i] * 5 for i in range(7)]
