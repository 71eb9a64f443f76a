"""Multi-GPU plumbing: utterances are independent (reference core.pyx:44-45 -- per-item calls share
nothing), so a batch is SHARDED across ranks and every rank aligns its own utterances with no
data-path collective.  The only exchange is the optional all-gather of the per-token durations
([B_local, Tx] int32) for loss bookkeeping / logging -- what the reference's DDP ranks never exchange
(each aligns its own per_gpu_batchsize shard, config.py:145) but a caller that wants global duration
statistics needs.

Backend-agnostic (`nccl` on GPUs, `gloo` in the CPU tests): only torch.distributed calls, no kernels.
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
