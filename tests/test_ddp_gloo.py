"""N>1 host logic on CPU: two gloo ranks run the bucketed gradient all-reduce of myolo/ddp.py on a
flat buffer with the engine's tail/head split and agree on the mean."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from myolo import ddp
    ar = ddp.BucketedAllReduce()
    n, tail = 1000, 640
    flat = torch.arange(n, dtype=torch.float32) * (rank + 1)
    s1 = ar(flat, tail, n)          # tail bucket first (from inside the backward pass) ...
    s2 = ar(flat, 0, tail)          # ... head bucket last: drains both
    assert s1 == s2 == 1.0 / world and ar._pending == []
    expect = torch.arange(n, dtype=torch.float32) * sum(r + 1 for r in range(world))
    ok = torch.equal(flat, expect)
    idx = list(ddp.shard_indices(10, rank, world))
    q.put((rank, ok, idx))
    dist.destroy_process_group()


def test_bucketed_allreduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    try:
        res = sorted(q.get(timeout=120) for _ in procs)
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
    finally:                                   # never leave a rank behind
        for p in procs:
            if p.is_alive():
                p.kill()
                p.join(timeout=10)
    assert res[0][1] and res[1][1]
    assert res[0][2] == [0, 1, 2, 3, 4] and res[1][2] == [5, 6, 7, 8, 9]
