"""Batch-sharded data parallelism: one process per GPU, one NCCL all-reduce (sum) over the flat gradient
bucket per optimizer step, 1/world folded into the fused Adam launch.  Nothing else shards
(SURVEY.md 8e): BatchNorm statistics and the cross-entropy class weights stay per-shard, exactly what
Keras' fit() would do with 32-sample mini-batches."""
import os

import torch.distributed as dist

from .models.trainers import Trainer


class _Dist(object):
    def __init__(self):
        self.world_size = dist.get_world_size()
        self.rank = dist.get_rank()

    # all-reduce pieces of the gradient arena while the backward pass is still running (Trainer.forward_backward);
    # DAFK_AR_OVERLAP=0 falls back to one all-reduce per bucket after the backward pass
    overlap = os.environ.get("DAFK_AR_OVERLAP", "1") != "0"

    def allreduce_(self, buckets):
        for b in buckets:
            dist.all_reduce(b, op=dist.ReduceOp.SUM)

    def allreduce_async(self, t):
        """enqueue on NCCL's own stream (it first waits for the work queued so far on the current stream); returns the handle"""
        return dist.all_reduce(t, op=dist.ReduceOp.SUM, async_op=True)

    def wait(self, works):
        for w in works:
            w.wait()          # the current stream waits for the collective (no host synchronisation)


def broadcast_weights(net):
    """every rank starts from rank 0's weights (arenas are flat, so this is one broadcast per arena)"""
    seen = set()
    models = list(net.Encoders_Anatomy) + [net.Enc_Modality, net.Anatomy_Fuser, net.Segmentor, net.Decoder, net.D_Mask]
    for m in models + [getattr(net, "D_Image1", None), getattr(net, "D_Image2", None), getattr(net, "Balancer", None)]:
        if m is None:
            continue
        for arena in (m._scope.arena, m._scope.state):
            if id(arena) in seen or arena.flat is None:
                continue
            seen.add(id(arena))
            dist.broadcast(arena.flat, src=0)
            arena.version += 1


def enable_data_parallel(net):
    assert dist.is_initialized()
    broadcast_weights(net)
    Trainer.dist = _Dist()
    return Trainer.dist
