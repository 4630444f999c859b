"""Data-parallel plumbing: clips are independent units, so ranks take contiguous shards of
the batch, weights are replicated, and the only collectives are the logit gather in evaluation and the
gradient all-reduce in training.  torch.distributed (NCCL on GPUs, gloo in CPU tests) is the transport."""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of rank `rank`; the first n_items % world ranks take one extra item."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(x: Dict[str, torch.Tensor], rank: int, world: int) -> Dict[str, torch.Tensor]:
    """Slice every tensor of a reference-style batch dict (train.py:207-218) along dim 0."""
    n = next(iter(x.values())).shape[0]
    lo, hi = shard_bounds(n, rank, world)
    return {k: v[lo:hi] for k, v in x.items()}


def gather_logits(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """All-gather per-rank [b_r, 21] outputs into the full [n_total, 21] batch (ragged shards allowed)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local
    rank = dist.get_rank(group)
    sizes = [shard_bounds(n_total, r, world) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world * width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad, group=group) if local.is_cuda else dist.all_gather(list(out.chunk(world)), pad, group=group)
    return torch.cat([out[r * width: r * width + (hi - lo)] for r, (lo, hi) in enumerate(sizes)], dim=0)


def allreduce_mean_(flat_grad: torch.Tensor, group=None) -> torch.Tensor:
    """Sum-all-reduce a flat gradient bucket and divide by the world size (equal valid-row counts per rank
    make this the global AULoss gradient, SURVEY.md §8e caveat 2)."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
        flat_grad.div_(dist.get_world_size(group))
    return flat_grad


class PipelinedLogitGather:
    """Evaluation logit gather that does not stall the next batch: submit(out21) copies the rank's [b, 21] logits into one of
    two staging buffers and starts the all-gather asynchronously (NCCL runs it on its own stream, ordered behind the copy);
    the gathered [world*b, 21] tensor of a submit is valid after wait() or after the submit two calls later.  The transformer
    hot path of batch i+1 thus overlaps the collective of batch i — clips are independent, nothing in the compute waits for
    the gather.  With one rank it degenerates to returning the input."""

    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.stage = [None, None]
        self.out = [None, None]
        self.work = [None, None]
        self.k = 0

    def submit(self, out21: torch.Tensor) -> torch.Tensor:
        if self.world == 1:
            return out21
        i = self.k & 1
        self.k += 1
        if self.work[i] is not None:
            self.work[i].wait()                       # the gather that last used this pair of buffers (two submits ago)
        if self.stage[i] is None or self.stage[i].shape != out21.shape:
            self.stage[i] = torch.empty_like(out21)
            self.out[i] = torch.empty((self.world * out21.shape[0],) + tuple(out21.shape[1:]), dtype=out21.dtype, device=out21.device)
        self.stage[i].copy_(out21, non_blocking=True)
        if out21.is_cuda:
            self.work[i] = dist.all_gather_into_tensor(self.out[i], self.stage[i], group=self.group, async_op=True)
        else:
            self.work[i] = dist.all_gather(list(self.out[i].chunk(self.world)), self.stage[i], group=self.group, async_op=True)
        return self.out[i]

    def wait(self) -> None:
        for i in (0, 1):
            if self.work[i] is not None:
                self.work[i].wait()
                self.work[i] = None


class SegmentReducer:
    """Gradient all-reduce overlapped with the backward pass.  The flat gradient bucket is laid out in contiguous segments, one
    per stack, in the order the stacks finish their backward (optim.FusedAdam).  ``ready(params)`` is called whenever the kernels
    producing some parameters' gradients have been enqueued; when the last parameter of a segment is reported its slice is
    all-reduced asynchronously (NCCL runs it on its own stream, ordered behind the stream that called), so the reduction of the
    fusion head's gradients travels over NVLink while the TFormer and SFormer are still in their backward.  ``finish()`` reduces
    whatever was never reported and joins everything into the calling stream.  Sums only: the 1 / world_size of the mean is
    folded into the optimiser's update kernel."""

    def __init__(self, flat_grad: torch.Tensor, bounds, counts, seg_of: Dict[int, int], group=None):
        self.g, self.bounds, self.counts, self.seg_of, self.group = flat_grad, list(bounds), list(counts), seg_of, group
        self.armed = False
        self.pending, self.works, self.launched = [], [], []
        self.launch_order = []                       # segment indices in the order their reductions were started (tests)

    def arm(self) -> None:
        self.pending = list(self.counts)
        self.works = [None] * len(self.bounds)
        self.launched = [False] * len(self.bounds)
        self.launch_order = []
        self.armed = True

    def disarm(self) -> None:
        self.armed = False

    def _launch(self, si: int) -> None:
        lo, hi = self.bounds[si]
        self.works[si] = dist.all_reduce(self.g[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        self.launched[si] = True
        self.launch_order.append(si)

    def ready(self, params) -> None:
        if not self.armed:
            return
        for prm in params:
            si = self.seg_of.get(id(prm))
            if si is None or self.launched[si]:
                continue
            self.pending[si] -= 1
            if self.pending[si] == 0:
                self._launch(si)

    def finish(self) -> None:
        if not self.armed:
            return
        for si in range(len(self.bounds)):
            if not self.launched[si]:
                self._launch(si)
        for i, w in enumerate(self.works):
            if w is not None:
                w.wait()
                self.works[i] = None
