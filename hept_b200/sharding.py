"""Event sharding for multi-GPU runs.

Every stage of the hot path (hashing, min/max, sort, block attention, combine, backward) is confined
to one event, so events shard across ranks with no data-path collective.  The only exchange in the
training configuration is the sum-all-reduce of parameter gradients; it is done on one flat bucket
(the attention module's trainable parameters are ~57 KB, the whole tracking model 1.3 MB — latency
bound, one NCCL launch).  Parameters that never receive a gradient in the reference (``w_rpe.bias``,
``e2lsh.alpha``, ``regions``) are skipped, which is why plain DDP would need
``find_unused_parameters=True``.
"""
from __future__ import annotations

from typing import Iterable, List, Sequence

import torch
import torch.distributed as dist


def events_of_rank(num_events: int, rank: int, world: int) -> List[int]:
    """Round-robin assignment: rank r owns events {e : e mod world == r}."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    return list(range(rank, num_events, world))


def allreduce_gradients(params: Iterable[torch.nn.Parameter], group=None, average: bool = True) -> int:
    """Sum (or average) the .grad of every parameter that has one, through ONE flat buffer.

    Returns the number of bytes reduced.  Every rank must hold gradients for the same parameters
    (true for this path: which parameters get gradients does not depend on the data).
    """
    grads = [p.grad for p in params if p is not None and p.grad is not None]
    if not grads or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return 0
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat /= dist.get_world_size(group)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off : off + n].view_as(g))
        off += n
    return flat.numel() * flat.element_size()
