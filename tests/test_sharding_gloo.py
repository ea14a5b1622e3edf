"""Host-side multi-rank logic on CPU: world_size 2, gloo backend (the N>1 path of bench.py minus the kernels)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ratrack_b200 import sharding


def test_shard_range_is_a_partition():
    for total in (0, 1, 7, 32, 1024, 1027):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_range(total, r, world) for r in range(world)]
            assert sum(c for _, c in spans) == total
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
            pos = 0
            for s, c in spans:
                assert s == pos
                pos += c
    with pytest.raises(ValueError):
        sharding.shard_range(8, 2, 2)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    start, count = sharding.shard_range(1027, rank, world)
    pairs, ms = sharding.job_throughput(count, 10.0 + 5.0 * rank)
    counts = [sharding.shard_range(1027, r, world)[1] for r in range(world)]
    flow = torch.full((count, 3, 4), float(rank))
    parts = sharding.gather_flow(flow, counts)
    ok = pairs == 1027 and ms == 10.0 + 5.0 * (world - 1) and all(
        p.shape[0] == c and bool((p == float(r)).all()) for r, (p, c) in enumerate(zip(parts, counts)))
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, ok))


def test_two_rank_gloo_throughput_and_gather():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
