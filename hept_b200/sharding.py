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


class GradBucket:
    """One persistent flat fp32 buffer whose slices ARE the parameters' ``.grad`` tensors.

    Autograd accumulates into the views in place, so after ``backward()`` the buffer already holds every gradient:
    the all-reduce runs on the buffer itself — one NCCL launch per step, no ``cat``, no divide kernel (``ReduceOp.AVG``),
    no copy-back.  Zero the gradients with ``bucket.zero()`` (one fill kernel), never with ``set_to_none=True``, which
    would detach the views.  Parameters that never receive a gradient in the reference (``w_rpe.bias``)
    simply stay zero: an optimiser step with a zero gradient leaves them unchanged, exactly like the reference's ``None``.
    """

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params = [p for p in params if p is not None and p.requires_grad]
        if not self.params:
            raise ValueError("GradBucket: no trainable parameters")
        dev = self.params[0].device
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            if p.device != dev or p.dtype != torch.float32:
                raise ValueError("GradBucket: parameters must be fp32 on one device")
            p.grad = self.flat[off: off + p.numel()].view_as(p)
            off += p.numel()

    @property
    def nbytes(self) -> int:
        return self.flat.numel() * 4

    def zero(self) -> None:
        self.flat.zero_()

    def attached(self) -> bool:
        """True while every parameter's .grad still is its view of the buffer."""
        off = 0
        for p in self.params:
            if p.grad is None or p.grad.data_ptr() != self.flat.data_ptr() + 4 * off:
                return False
            off += p.numel()
        return True

    def allreduce(self, group=None, average: bool = True, async_op: bool = False):
        """Sum (average) the bucket over the ranks in place.  Returns the work handle when ``async_op``."""
        if not dist.is_initialized() or dist.get_world_size(group) == 1:
            return None
        if average and dist.get_backend(group) == "nccl":
            return dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=group, async_op=async_op)
        work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        if average:
            if async_op:
                work.wait()
                work = None
            self.flat /= dist.get_world_size(group)
        return work
