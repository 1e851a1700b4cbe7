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


class BucketedGradReducer:
    """DDP's gradient exchange for the stage-1 / stage-2 training step (train_stage2.py:38 [R]: DistributedDataParallel,
    `ddp_find_unused_parameters=False`), built for NVSwitch: the trainable parameters' gradients live in a few large flat
    buckets (sized for launch latency, not link count), every bucket is all-reduced on a side stream as soon as autograd has
    produced its last gradient, and the optimizer waits on the side stream - so the exchange overlaps the rest of backward.

    `params` in FORWARD order; buckets are filled in reverse (the order backward produces gradients).  `p.grad` of every
    parameter is a view into its bucket, so nothing is copied in or out.  Averaging (DDP semantics) is folded into the
    all-reduce as a pre-scale of 1/world.  Works with nccl (device streams) and gloo (CPU tests: synchronous)."""

    def __init__(self, params, bucket_bytes: int = 256 << 20, group=None, average: bool = True):
        self.group, self.average = group, average
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.params = [p for p in params if p.requires_grad]
        self.buckets = []        # (flat tensor, [params])
        self._pending = {}
        cur, cur_bytes = [], 0
        for p in reversed(self.params):
            nb = p.numel() * 4
            if cur and cur_bytes + nb > bucket_bytes:
                self._close(cur)
                cur, cur_bytes = [], 0
            cur.append(p)
            cur_bytes += nb
        if cur:
            self._close(cur)
        self.cuda = bool(self.params) and self.params[0].is_cuda
        self.stream = torch.cuda.Stream() if self.cuda else None
        self.exposed_ms = 0.0
        self._ev_bwd = self._ev_comm = None
        for bi, (_, ps) in enumerate(self.buckets):
            for p in ps:
                p.register_post_accumulate_grad_hook(self._make_hook(bi))

    def _close(self, ps):
        flat = torch.zeros(sum(p.numel() for p in ps), device=ps[0].device, dtype=torch.float32)
        off = 0
        for p in ps:
            assert p.dtype == torch.float32, "master parameters are fp32 (bf16 autocast training)"
            p.grad = flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        self.buckets.append((flat, ps))

    def _make_hook(self, bi):
        def hook(p):
            left = self._pending.get(bi, len(self.buckets[bi][1])) - 1
            self._pending[bi] = left
            if left == 0:
                self._reduce(bi)
        return hook

    def _reduce(self, bi):
        flat = self.buckets[bi][0]
        if self.world == 1:
            return
        if self.cuda:
            self.stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.stream):
                if self.average:
                    flat.mul_(1.0 / self.world)
                dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
        else:
            if self.average:
                flat.mul_(1.0 / self.world)
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)

    def zero_grad(self):
        """Zero the buckets in place (the grads stay views of them) and re-arm the per-bucket counters."""
        for flat, ps in self.buckets:
            flat.zero_()
            off = 0
            for p in ps:  # an optimizer / user may have replaced .grad: re-attach the views
                if p.grad is None or p.grad.data_ptr() != flat.data_ptr() + 4 * off:
                    p.grad = flat[off:off + p.numel()].view_as(p)
                off += p.numel()
        self._pending = {}

    def finish(self, timed: bool = False):
        """Call after backward(): the current stream waits for the outstanding all-reduces.  With `timed`, records the events
        that `exposed()` turns into the non-overlapped part of the exchange."""
        if not self.cuda or self.world == 1:
            return
        if timed:
            self._ev_bwd = torch.cuda.Event(enable_timing=True)
            self._ev_comm = torch.cuda.Event(enable_timing=True)
            self._ev_bwd.record(torch.cuda.current_stream())
            self._ev_comm.record(self.stream)
        torch.cuda.current_stream().wait_stream(self.stream)

    def exposed(self) -> float:
        """ms between the end of backward on the compute stream and the end of the last all-reduce (after a synchronize)."""
        if self._ev_bwd is None:
            return 0.0
        return max(0.0, self._ev_bwd.elapsed_time(self._ev_comm))

    def total_bytes(self) -> int:
        return sum(f.numel() * 4 for f, _ in self.buckets)
