"""Two-GPU data parallelism on real devices (skipped on a single-GPU box; run with `gpurun --gpus 2`):
after one step with the bucketed NCCL all-reduce, both replicas hold identical weights, and their update
equals Adam applied to the MEAN of the two single-replica gradients."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q, transport="torch"):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "mask-yolo_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    from myolo import ddp
    from myolo.engine import Engine, init_params
    from tests import helpers as Hh
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    c = Hh.engine_cfg(S=96)
    P = init_params(c["NB"], c["NC"], 5 + rank, "trained_like")         # different weights per rank: broadcast must fix it
    img = torch.rand(2, 96, 96, 3, generator=torch.Generator().manual_seed(40 + rank))
    inputs = Hh.to_device(Hh.batch_from_boxes(c, 2, img, Hh.random_boxes(2, 2, 50 + rank), 60 + rank), f"cuda:{rank}")
    eng = Engine(c, 2, "training", "h16", device=rank, params=P)

    class M:        # the attribute surface ddp.attach needs
        engine = eng
        allreduce = None
    ddp.attach(M, transport=transport)
    assert type(M.allreduce).__name__ == {"torch": "BucketedAllReduce", "cabi": "CabiAllReduce"}[transport]
    p0 = eng.params.clone()
    # single-replica gradient of this rank (no exchange), then the data-parallel step
    eng.forward_training(inputs)
    eng.backward()
    g_local = eng.grads.clone()
    g_all = [torch.empty_like(g_local) for _ in range(world)]
    dist.all_gather(g_all, g_local)
    g_mean = sum(g_all) / world
    eng.t, eng.seen = 0, 0
    eng.params.copy_(p0); eng.adam_m.zero_(); eng.adam_v.zero_(); eng.refresh_weights()
    eng.train_step(inputs, 1e-3, M.allreduce)
    torch.cuda.synchronize()
    expect = p0 - 1e-3 * g_mean / (g_mean.abs() + 1e-8)                # Keras Adam, first step
    big = g_mean.abs() > 1e-3 * g_mean.abs().max()
    err = (eng.params - expect)[big].abs().max().item()
    ps = [torch.empty_like(eng.params) for _ in range(world)]
    dist.all_gather(ps, eng.params)
    same = all(torch.equal(ps[0], t) for t in ps)
    q.put((rank, err, same))
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("transport", ["torch", "cabi"])
def test_two_gpu_step_is_mean_of_replica_gradients(transport):
    """transport 'torch': torch.distributed all_reduce; 'cabi': the C ABI's own NCCL communicator (myolo_allreduce_*)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, transport)) for r in range(2)]
    for p in procs:
        p.start()
    try:
        res = sorted(q.get(timeout=180) for _ in procs)
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
    finally:                                   # never leave a rank behind (a deadlocked collective would hang pytest's exit)
        for p in procs:
            if p.is_alive():
                p.kill()
                p.join(timeout=10)
    for rank, err, same in res:
        assert same, "replicas must hold identical weights after the step"
        assert err <= 2e-5, (rank, err)


def _cabi_world1_worker(port, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "mask-yolo_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    try:
        import torch.distributed as dist
        from myolo import ddp
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        torch.cuda.set_device(0)
        dist.init_process_group("gloo", rank=0, world_size=1)
        ar = ddp.CabiAllReduce(device="cuda:0")
        flat = torch.arange(100000, dtype=torch.float32, device="cuda") * 0.5
        ref = flat.clone()
        s1 = ar(flat, 60000, 100000)          # tail bucket first, as Engine.train_step issues them
        s2 = ar(flat, 0, 60000)
        torch.cuda.synchronize()
        ok = bool(torch.equal(flat, ref)) and s1 == 1.0 and s2 == 1.0 and ar.comm
        ar.close()
        ok = ok and ar.comm is None
        dist.destroy_process_group()
        q.put(("ok" if ok else "mismatch", ""))
    except Exception as e:                     # report instead of hanging the parent
        import traceback
        q.put(("error", traceback.format_exc()[-1500:]))


def test_cabi_allreduce_single_rank_communicator():
    """myolo_allreduce_unique_id / _init / _run / _destroy on one GPU: a one-rank NCCL communicator created through the
    C ABI sums a buffer with itself (identity), on the side stream, in the two-bucket order of a training step."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    p = ctx.Process(target=_cabi_world1_worker, args=(_free_port(), q))
    p.start()
    try:
        status, detail = q.get(timeout=120)
        p.join(timeout=60)
    finally:
        if p.is_alive():
            p.kill()
            p.join(timeout=10)
    assert status == "ok", detail
    assert p.exitcode == 0
