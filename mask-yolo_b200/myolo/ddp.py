"""Data parallelism for Mask-YOLO training: one process per GPU, batch sharded across ranks, ONE
gradient exchange per step.  The flat gradient buffer is all-reduced (sum) in two
buckets -- the feature_map + mask-head tail as soon as the mask-branch backward has produced it
(asynchronously; under the backbone backward when every filter gradient is issued in line, at the join of the
filter-gradient stream in the default h16 schedule, Engine.backward), the backbone/yolo head at the end -- and the
1/world factor is folded into the fused Adam kernel.  BatchNorm statistics and the loss normalisers
stay per replica, so an N-rank step is the mean of N single-replica reference steps.
The reference has no distributed code.  Two transports carry the exchange: `torch.distributed` (NCCL over NVLink on
GPUs, gloo in the CPU tests; the default) and the C ABI's own NCCL communicator (`myolo_allreduce_*`,
transport="cabi": the rendezvous still rides on torch.distributed, the data path does not)."""
from __future__ import annotations

import ctypes
import glob
import os

import torch
import torch.distributed as dist


class BucketedAllReduce(object):
    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self._pending = []

    def __call__(self, flat: torch.Tensor, lo: int, hi: int):
        """Called by Engine.train_step: first for the tail bucket [tail_off:n) from inside the backward
        pass, then for the head bucket [0:tail_off) after it.  Returns the gradient scale 1/world."""
        if hi > lo:
            self._pending.append(dist.all_reduce(flat[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        if lo == 0:                                   # head bucket = last call of the step: drain
            for w in self._pending:
                w.wait()
            self._pending = []
        return 1.0 / self.world


def load_nccl_global():
    """Make the symbols of the libnccl.so.2 PyTorch ships (and has loaded) visible to libmyolo_sm100.so's dlsym."""
    roots = [os.path.dirname(os.path.dirname(torch.__file__))]
    for r in roots:
        for path in glob.glob(os.path.join(r, "nvidia", "nccl", "lib", "libnccl.so.2")):
            ctypes.CDLL(path, mode=ctypes.RTLD_GLOBAL)
            return path
    ctypes.CDLL("libnccl.so.2", mode=ctypes.RTLD_GLOBAL)        # system library on the loader path
    return "libnccl.so.2"


class CabiAllReduce(object):
    """The same two-bucket exchange through `myolo_allreduce_run` on a side stream: the tail bucket is issued from inside
    the backward pass (ordered after the mask-branch backward by an event), the head bucket at the end, and the step's
    stream waits for both before Adam.  The calls bypass the engine's launch recording on purpose: the hook that issues
    them is itself replayed as a host action."""

    def __init__(self, group=None, device=None):
        from . import _cabi as C
        self.C = C
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        load_nccl_global()
        idbuf = ctypes.create_string_buffer(128)
        if self.rank == 0:
            C.call("myolo_allreduce_unique_id", idbuf)
        box = [idbuf.raw]
        dist.broadcast_object_list(box, src=0, group=group)          # the side channel for the 128 id bytes
        self._id = ctypes.create_string_buffer(box[0], 128)
        handle = ctypes.c_void_p()
        with torch.cuda.device(self.dev):
            C.call("myolo_allreduce_init", self._id, self.rank, self.world, ctypes.addressof(handle))
        self.comm = handle.value
        self.stream = torch.cuda.Stream(device=self.dev)
        self._run = C.lib().myolo_allreduce_run

    def __call__(self, flat: torch.Tensor, lo: int, hi: int):
        main = torch.cuda.current_stream(self.dev)
        if hi > lo:
            self.stream.wait_stream(main)                             # the slice is final on the step's stream
            part = flat[lo:hi]
            rc = self._run(self.comm, part.data_ptr(), hi - lo, self.stream.cuda_stream)
            if rc != 0:
                raise self.C.MyoloError("myolo_allreduce_run failed (%d): %s" % (rc, self.C.lib().myolo_last_error().decode()))
            self.C.launch_count += 1
        if lo == 0:                                                   # head bucket = last call of the step
            main.wait_stream(self.stream)
        return 1.0 / self.world

    def close(self):
        if self.comm:
            torch.cuda.synchronize(self.dev)
            self.C.call("myolo_allreduce_destroy", self.comm)
            self.comm = None


def broadcast_parameters(engine, src: int = 0, group=None):
    """Identical weights on every replica (rank `src` wins), then re-stage the GEMM weight copies."""
    dist.broadcast(engine.params, src=src, group=group)
    for t in engine.stats.values():
        dist.broadcast(t, src=src, group=group)
    engine.refresh_weights()


def attach(model, group=None, transport=None):
    """Make `model` (a MaskYOLO in training/yolo mode) data-parallel over the default process group.
    transport: "torch" (default; torch.distributed all_reduce) or "cabi" (myolo_allreduce_* of the C ABI); the
    environment variable MYOLO_DDP_TRANSPORT overrides the default."""
    transport = transport or os.environ.get("MYOLO_DDP_TRANSPORT", "torch")
    if transport not in ("torch", "cabi"):
        raise ValueError("transport must be 'torch' or 'cabi', got %r" % (transport,))
    broadcast_parameters(model.engine, 0, group)
    model.allreduce = CabiAllReduce(group, device=model.engine.dev) if transport == "cabi" else BucketedAllReduce(group)
    return model


def shard_indices(n_items: int, rank: int, world: int):
    """Contiguous slice of the BatchGenerator index space owned by `rank`."""
    per = n_items // world
    return range(rank * per, (rank + 1) * per)
