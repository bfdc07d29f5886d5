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
    the all-reduce runs on the buffer itself — one launch per step, no ``cat``, no divide kernel, no copy-back.  Zero the
    gradients with ``bucket.zero()`` (one fill kernel), never with ``set_to_none=True``, which would detach the views.
    Parameters that never receive a gradient in the reference (``w_rpe.bias``) simply stay zero: an
    optimiser step with a zero gradient leaves them unchanged, exactly like the reference's ``None``.

    ``p2p=True`` (CUDA, one node): the buffer is allocated as torch symmetric memory, mapped into every peer over NVLink, and
    ``allreduce()`` is ONE kernel of the library (``hept_p2p_allreduce``, csrc/p2p.cu): every rank reads every peer's bucket
    directly and sums in rank order — two flag round trips instead of a ring, the same bits on every rank.  torch only
    provides the mapping.  If the mapping cannot be set up (no peer access, no symmetric-memory support) the bucket says so in
    ``p2p_error`` and uses NCCL (``ReduceOp.AVG``), which is also what ``p2p=False`` does.
    """

    P2P_MAX_BYTES = 256 * 1024

    def __init__(self, params: Iterable[torch.nn.Parameter], p2p: bool = False, group=None):
        self.params = [p for p in params if p is not None and p.requires_grad]
        if not self.params:
            raise ValueError("GradBucket: no trainable parameters")
        dev = self.params[0].device
        total = sum(p.numel() for p in self.params)
        self._p2p, self.p2p_error, self._seq = None, None, 0
        self.flat = None
        # one-shot pays W - 1 remote reads of the whole bucket: it wins while the bucket is latency-bound (measured on 8 B200:
        # 57 KB 22.7 us against NCCL's 32.5; 1.3 MB 45 us against 32), so larger buckets stay on NCCL
        if (p2p and total * 4 <= self.P2P_MAX_BYTES and dev.type == "cuda" and dist.is_initialized()
                and dist.get_world_size(group) > 1):
            try:
                self._setup_p2p(total, dev, group)
            except Exception as e:           # noqa: BLE001  (any failure of the mapping means: use NCCL, and say why)
                self._p2p, self.p2p_error = None, f"{type(e).__name__}: {e}"
        if self.flat is None:
            self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            if p.device != dev or p.dtype != torch.float32:
                raise ValueError("GradBucket: parameters must be fp32 on one device")
            p.grad = self.flat[off: off + p.numel()].view_as(p)
            off += p.numel()

    def _setup_p2p(self, total: int, dev, group) -> None:
        import torch.distributed._symmetric_memory as symm

        from . import _lib

        lib = _lib.load()
        flag_floats = lib.hept_p2p_flag_bytes() // 4
        padded = (total + 3) // 4 * 4
        pg = group if group is not None else dist.group.WORLD
        name = pg.group_name
        with torch.cuda.device(dev):
            buf = symm.empty(flag_floats + padded, dtype=torch.float32, device=dev)
            buf.zero_()
            hdl = symm.rendezvous(buf, name)
            torch.cuda.synchronize(dev)
        dist.barrier(group)                  # every rank's flags are zero before anybody signals
        ptrs = hdl.buffer_ptrs_dev
        self._p2p = {"hdl": hdl, "buf": buf, "ptrs": int(ptrs), "padded": padded, "rank": dist.get_rank(group),
                     "world": dist.get_world_size(group), "scratch": torch.empty(padded, dtype=torch.float32, device=dev),
                     "err": torch.zeros(1, dtype=torch.int32, device=dev)}
        self.flat = buf[flag_floats: flag_floats + total]

    @property
    def nbytes(self) -> int:
        return self.flat.numel() * 4

    @property
    def uses_p2p(self) -> bool:
        return self._p2p is not None

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

    def p2p_ok(self) -> bool:
        """False if a peer ever failed to arrive within the kernel's wait bound (reads the device flag: synchronises)."""
        return self._p2p is None or int(self._p2p["err"].item()) == 0

    def allreduce(self, group=None, average: bool = True, async_op: bool = False):
        """Sum (average) the bucket over the ranks in place.  Returns the work handle when ``async_op`` (NCCL path only)."""
        if not dist.is_initialized() or dist.get_world_size(group) == 1:
            return None
        if self._p2p is not None:
            import ctypes as C

            from . import _lib

            st = self._p2p
            self._seq += 1
            with torch.cuda.device(self.flat.device):
                _lib.check(_lib.load().hept_p2p_allreduce(
                    C.c_void_p(st["ptrs"]), st["rank"], st["world"], st["padded"], self._seq, 1.0 / st["world"] if average else 1.0,
                    C.c_void_p(st["scratch"].data_ptr()), C.c_void_p(st["err"].data_ptr()),
                    C.c_void_p(torch.cuda.current_stream(self.flat.device).cuda_stream)), "hept_p2p_allreduce")
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
