"""Multi-GPU plumbing: utterances are independent (reference core.pyx:44-45 -- per-item calls share
nothing), so a batch is SHARDED across ranks and every rank aligns its own utterances with no
data-path collective.  The only exchange is the optional all-gather of the per-token durations
([B_local, Tx] int32) for loss bookkeeping / logging -- what the reference's DDP ranks never exchange
(each aligns its own per_gpu_batchsize shard, config.py:145) but a caller that wants global duration
statistics needs.

`all_gather_durations*` are backend-agnostic (`nccl` on GPUs, `gloo` in the CPU tests): only torch.distributed calls, no
kernels.  `OneSidedDurationGather` is the GPU form the bench uses: symmetric-memory buffers and stores over NVLink
(mas_b200_put_durations, or the fused kernel's own output stage) instead of a collective -- a torch.distributed call per
step costs more host time than the alignment step takes on a B200.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of rank `rank`; the remainder goes to the first ranks (sizes differ by <= 1)."""
    if world <= 0 or not (0 <= rank < world) or n_items < 0:
        raise ValueError("bad shard arguments")
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def balanced_shards(t_x: Sequence[int], t_y: Sequence[int], world: int) -> List[List[int]]:
    """Utterance indices per rank, balancing the DP work sum(t_x * t_y): longest-first greedy onto the
    least-loaded rank (ties -> lowest rank), each rank's list returned in ascending index order.  Every
    index appears exactly once."""
    if world <= 0:
        raise ValueError("world must be positive")
    cost = [int(a) * int(b) for a, b in zip(t_x, t_y)]
    order = sorted(range(len(cost)), key=lambda i: (-cost[i], i))
    loads = [0] * world
    out: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (loads[k], k))
        out[r].append(i)
        loads[r] += cost[i]
    return [sorted(v) for v in out]


def all_gather_durations(dur_local: torch.Tensor, group: Optional[dist.ProcessGroup] = None,
                         out: Optional[torch.Tensor] = None, async_op: bool = False):
    """durations [B_local, Tx] int32 of every rank -> [sum B_local, Tx] on every rank (rank order).
    Equal shard sizes take one all_gather_into_tensor; uneven shards are padded to the largest and trimmed.
    Returns the gathered tensor (and the work handle when async_op)."""
    if not dist.is_initialized():
        return (dur_local, None) if async_op else dur_local
    world = dist.get_world_size(group)
    n_local = torch.tensor([dur_local.shape[0]], dtype=torch.int64, device=dur_local.device)
    sizes = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(sizes, n_local, group=group)
    sizes = [int(s.item()) for s in sizes]
    Tx = dur_local.shape[1]
    if len(set(sizes)) == 1:
        if out is None:
            out = torch.empty((sum(sizes), Tx), dtype=dur_local.dtype, device=dur_local.device)
        work = dist.all_gather_into_tensor(out, dur_local.contiguous(), group=group, async_op=async_op)
        return (out, work) if async_op else out
    m = max(sizes)
    padded = torch.zeros((m, Tx), dtype=dur_local.dtype, device=dur_local.device)
    padded[: dur_local.shape[0]] = dur_local
    buf = torch.empty((world * m, Tx), dtype=dur_local.dtype, device=dur_local.device)
    dist.all_gather_into_tensor(buf, padded, group=group)
    res = torch.cat([buf[r * m: r * m + sizes[r]] for r in range(world)], dim=0)
    return (res, None) if async_op else res


def all_gather_durations_into(out: torch.Tensor, dur_local: torch.Tensor, group: Optional[dist.ProcessGroup] = None):
    """Equal-shard fast path with a caller-owned output [world * B_local, Tx]: one collective, no size exchange
    (what bench.py issues every step on its side stream)."""
    return dist.all_gather_into_tensor(out, dur_local.contiguous(), group=group)


class OneSidedDurationGather:
    """The duration all-gather without a collective: every rank's fused kernel stores its [B_local, Tx] durations
    straight into rows [rank * B_local, +B_local) of EVERY rank's gather buffer (peer-to-peer stores over NVLink from the
    kernel's output stage; `MasParams::peer_dur` in csrc/mas_forward.cuh).  No NCCL kernel runs beside the next step and
    nothing sits on the host path of a step -- torch.distributed costs ~25 us of host time per call, more than half of
    the 37 us step it would accompany.

    The buffers are symmetric memory (torch.distributed._symmetric_memory: one allocation per rank, mapped into every
    peer's address space).  `gathered` holds the durations of the LAST step every rank has completed; a reader
    synchronises with the writers the way it would for any one-sided put: stream-sync + `barrier()` (or any later
    collective) before it reads, and before the step after next overwrites the rows.

    GPU only (NVLink / PCIe peer access between the ranks' devices); raises if symmetric memory cannot be set up --
    callers fall back to `all_gather_durations_into`."""

    def __init__(self, b_local: int, tx: int, device, group: Optional[dist.ProcessGroup] = None):
        import torch.distributed._symmetric_memory as symm

        group = group or dist.group.WORLD
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.gathered = symm.empty((self.world * b_local, tx), dtype=torch.int32, device=device)
        self.gathered.zero_()
        self.handle = symm.rendezvous(self.gathered, group)
        self._ptrs_dev = int(self.handle.buffer_ptrs_dev)
        self._on = False

    def put(self, dur_local: torch.Tensor, stream: Optional[torch.cuda.Stream] = None) -> None:
        """Stream-ordered put of this rank's durations [B_local, Tx] into every rank's buffer by a small copy kernel
        (mas_b200_put_durations) -- for callers that keep the gather off the step's critical path on a side stream."""
        from . import _lib

        st = stream if stream is not None else torch.cuda.current_stream(dur_local.device)
        _lib.check(_lib.lib().mas_b200_put_durations(dur_local.data_ptr(), dur_local.shape[0], dur_local.shape[1], self._ptrs_dev,
                                                     self.world, self.rank, st.cuda_stream), "mas_b200_put_durations")

    def enable(self) -> None:
        """From now on the fused kernel of this process stores its durations itself (no extra launch; the remote stores
        add their NVLink round trip to the end of the kernel: +4 us per step measured at N = 2)."""
        from . import _lib

        _lib.check(_lib.lib().mas_b200_set_pointer_option(b"peer_dur_ptrs", self._ptrs_dev), "mas_b200_set_pointer_option")
        _lib.set_option("peer_rank", self.rank)
        _lib.set_option("peer_world", self.world)
        self._on = True

    def disable(self) -> None:
        from . import _lib

        if self._on:
            _lib.set_option("peer_world", 0)
            _lib.check(_lib.lib().mas_b200_set_pointer_option(b"peer_dur_ptrs", None), "mas_b200_set_pointer_option")
            self._on = False

    def barrier(self) -> None:
        """All ranks' puts issued before this point (on their current streams) are complete and visible."""
        torch.cuda.current_stream(self.gathered.device).synchronize()
        dist.barrier()
