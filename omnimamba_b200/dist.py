"""Data-parallel plumbing for the scan path: the scan is sequence-local and independent per (batch, head),
so ranks hold replicas and their own batch shard; the only exchange is the all-reduce of PARAMETER gradients after backward
(DDP semantics of train_stage2.py:38 [R]).  Backend-agnostic (nccl on the B200 box, gloo in the CPU tests)."""
from __future__ import annotations

from typing import Iterable, List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """[begin, end) of rank's contiguous share of n units; shares differ by at most one unit, earlier ranks take the extra."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(n, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def shard_batch(tensors: Iterable[Optional[torch.Tensor]], rank: int, world: int) -> List[Optional[torch.Tensor]]:
    """Slice dim 0 (batch) of every tensor to this rank's share (views, no copies)."""
    out = []
    for t in tensors:
        if t is None:
            out.append(None)
            continue
        b, e = shard_range(t.shape[0], rank, world)
        out.append(t[b:e])
    return out


def allreduce_param_grads(grads: List[Optional[torch.Tensor]], average: bool = False, group=None) -> List[Optional[torch.Tensor]]:
    """One flat all-reduce (SUM, or mean when `average`) over the given parameter gradients, in place.  The flat bucket is
    fp32: these gradients (dA, dD, ddt_bias, conv / norm weights) are tiny next to the activations, so a single bucket sized
    for launch latency - not link count - is the right shape on NVSwitch."""
    live = [g for g in grads if g is not None]
    if not live or not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return grads
    flat = torch.cat([g.reshape(-1).float() for g in live])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat /= dist.get_world_size(group)
    off = 0
    for g in live:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n
    return grads


def max_over_ranks(value: float, device=None, group=None) -> float:
    """Timing rule: a multi-GPU step takes as long as its slowest rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
