"""Data parallelism for Mask-YOLO training: one process per GPU, batch sharded across ranks, ONE
gradient exchange per step.  The flat gradient buffer is all-reduced (sum) in two
buckets -- the feature_map + mask-head tail as soon as the mask-branch backward has produced it
(asynchronously, overlapping the backbone backward), the backbone/yolo head at the end -- and the
1/world factor is folded into the fused Adam kernel.  BatchNorm statistics and the loss normalisers
stay per replica, so an N-rank step is the mean of N single-replica reference steps.
The reference has no distributed code; `torch.distributed` (NCCL over NVLink on GPUs, gloo in the CPU
tests) is the transport."""
from __future__ import annotations

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


def broadcast_parameters(engine, src: int = 0, group=None):
    """Identical weights on every replica (rank `src` wins), then re-stage the GEMM weight copies."""
    dist.broadcast(engine.params, src=src, group=group)
    for t in engine.stats.values():
        dist.broadcast(t, src=src, group=group)
    engine.refresh_weights()


def attach(model, group=None):
    """Make `model` (a MaskYOLO in training/yolo mode) data-parallel over the default process group."""
    broadcast_parameters(model.engine, 0, group)
    model.allreduce = BucketedAllReduce(group)
    return model


def shard_indices(n_items: int, rank: int, world: int):
    """Contiguous slice of the BatchGenerator index space owned by `rank`."""
    per = n_items // world
    return range(rank * per, (rank + 1) * per)
