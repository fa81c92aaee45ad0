"""CPU, world_size 2, gloo: the N>1 host logic (page sharding, max-over-ranks timing) used by bench.py."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from smart_image_processing_b200 import sharding


def test_shard_bounds_cover_and_balance():
    for n in (0, 1, 7, 256, 4096, 4097):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                lo, hi = sharding.shard_bounds(n, r, world)
                assert 0 <= lo <= hi <= n
                seen.extend(range(lo, hi))
            assert seen == list(range(n))
            sizes = [len(sharding.page_ids(n, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_bounds(4, 4, 4)
    assert list(sharding.weak_batch_seeds(4, 2)) == [8, 9, 10, 11]


def _worker(rank, world, port, n_pages, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ids = list(sharding.page_ids(n_pages, rank, world))
    # every rank processes only its own pages (a checksum stands in for the pixel work)
    local = torch.tensor([sum(i * i + 1 for i in ids), len(ids)], dtype=torch.int64)
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    slowest = sharding.max_over_ranks(10.0 + rank)
    dist.barrier()
    if rank == 0:
        out.put((int(sum(g[0] for g in gathered)), int(sum(g[1] for g in gathered)), slowest))
    dist.destroy_process_group()


def test_two_ranks_partition_pages_and_reduce_time():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    n_pages, world = 37, 2
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_pages, out)) for r in range(world)]
    for p in procs:
        p.start()
    checksum, count, slowest = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert count == n_pages and checksum == sum(i * i + 1 for i in range(n_pages))
    assert slowest == 11.0
