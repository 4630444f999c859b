"""Data-parallel plumbing: clips are independent units, so ranks take contiguous shards of
the batch, weights are replicated, and the only collectives are the logit gather in evaluation and the
gradient all-reduce in training.  torch.distributed (NCCL on GPUs, gloo in CPU tests) is the transport."""
from __future__ import annotations

from typing import Dict, Optional, Tuple

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


class PeerLogitGather:
    """The evaluation-time logit gather as PUSHES over NVLink peer memory (csrc/avf_peer.cu, include/avformer_b200.h
    avf_logits_push / avf_logits_wait) instead of a collective: ``push(out21)`` stores the rank's [b, 21] logits straight into every
    peer's table and returns at once; ``wait()`` — issued at the end of the step, behind whatever independent work follows the
    fusion head — blocks the STREAM (not the host) until the blocks of all ranks for that step have arrived, which they usually
    have long before.  No per-step rendezvous: a rank may run ahead of the slowest peer by most of a step.

    The peer-mapped block comes from torch symmetric memory (torch.distributed._symmetric_memory: allocation + exchange of the
    handles inside one node); the kernels are ours.  Every rank must issue the same sequence of push / wait pairs.  Both calls are
    graph-capturable: the step counter lives on the device; ``table()`` follows it on the host (``note_replay()`` after each replay
    of a graph that contains a captured pair)."""

    def __init__(self, rows_per_rank: int, cols: int = 21, group=None, timeout_s: float = 5.0):
        import ctypes
        import torch.distributed._symmetric_memory as symm
        from . import _lib
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("PeerLogitGather needs an initialised torch.distributed process group (NCCL, one node)")
        self._ct = ctypes
        self._L = _lib
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.rows, self.cols, self.n = int(rows_per_rank), int(cols), int(rows_per_rank) * int(cols)
        self.timeout_ns = int(timeout_s * 1e9)
        dev = torch.device("cuda", torch.cuda.current_device())
        nbytes = int(_lib.lib().avf_peer_gather_bytes(self.world, self.n))
        self.block = symm.empty((nbytes + 3) // 4, dtype=torch.float32, device=dev)
        self.block.zero_()
        self.handle = symm.rendezvous(self.block, self.group)
        ptrs = [int(q) for q in self.handle.buffer_ptrs]
        if len(ptrs) != self.world or ptrs[self.rank] != self.block.data_ptr():
            raise RuntimeError("PeerLogitGather: symmetric-memory rendezvous returned an unexpected pointer table")
        self.peer_base = torch.tensor(ptrs, dtype=torch.int64, device=dev)
        self.state = torch.zeros(2, dtype=torch.int32, device=dev)          # [0] step counter, [1] error word
        self.done = 0                                                        # host mirror of the step counter
        torch.cuda.synchronize()
        dist.barrier(self.group)                                             # every table is zeroed before anybody pushes

    def push(self, out21: torch.Tensor) -> None:
        if not (out21.is_cuda and out21.dtype == torch.float32 and out21.is_contiguous() and out21.numel() == self.n):
            raise ValueError(f"PeerLogitGather.push: expected a contiguous fp32 CUDA tensor of {self.rows} x {self.cols}, got {tuple(out21.shape)} {out21.dtype}")
        c = self._ct.c_void_p
        self._L.check(self._L.lib().avf_logits_push(c(out21.data_ptr()), self.n, c(self.peer_base.data_ptr()), self.world, self.rank,
                                                    c(self.state.data_ptr()), c(torch.cuda.current_stream().cuda_stream)), "avf_logits_push")

    def wait(self) -> None:
        c = self._ct.c_void_p
        self._L.check(self._L.lib().avf_logits_wait(c(self.block.data_ptr()), self.n, self.world, c(self.state.data_ptr()), self.timeout_ns,
                                                    c(torch.cuda.current_stream().cuda_stream)), "avf_logits_wait")
        if not torch.cuda.is_current_stream_capturing():
            self.done += 1

    def note_replay(self, pairs: int = 1) -> None:
        """A captured graph holding `pairs` push / wait pairs has been replayed."""
        self.done += pairs

    def table(self) -> torch.Tensor:
        """[world * rows, cols] logits of the last completed step (stream-ordered behind its wait; overwritten two steps later)."""
        if self.done == 0:
            raise RuntimeError("PeerLogitGather.table() before the first push / wait pair")
        slot = (self.done - 1) & 1
        return self.block[slot * self.world * self.n:(slot + 1) * self.world * self.n].view(self.world * self.rows, self.cols)

    def check(self) -> None:
        """Synchronise and raise if a wait ever timed out (a rank that stopped pushing) or the host mirror lost count."""
        torch.cuda.synchronize()
        step, err = (int(v) for v in self.state.tolist())
        if err != 0:
            raise RuntimeError(f"PeerLogitGather: the block of rank {err - 1} did not arrive within {self.timeout_ns / 1e9:.1f} s")
        if step != self.done:
            raise RuntimeError(f"PeerLogitGather: device step counter {step} != host count {self.done} (missing note_replay()?)")


class PeerAllReduce:
    """The gradient all-reduce of data-parallel training as ONE kernel over NVLink peer memory (csrc/avf_peer.cu,
    avf_grad_allreduce): the flat fp32 gradient bucket lives in a peer-mapped block (``.grad``, n floats); ``reduce_()`` sums the
    buckets of all ranks in place — rank r adds up slice r straight from the peers' memory, in rank order, and stores it into every
    bucket, so all ranks end with identical bits and the result does not depend on timing.  Asynchronous on the current stream,
    graph-capturable, issued once per step on every rank.  Sums only; the optimiser folds 1 / world into its update."""

    def __init__(self, n_floats: int, device=None, group=None, timeout_s: float = 30.0):
        import ctypes
        import torch.distributed._symmetric_memory as symm
        from . import _lib
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("PeerAllReduce needs an initialised torch.distributed process group (NCCL, one node)")
        self._ct, self._L = ctypes, _lib
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        if self.world > 16:
            raise RuntimeError("PeerAllReduce: at most 16 ranks (one NVLink domain)")
        self.n = int(n_floats)
        self.timeout_ns = int(timeout_s * 1e9)
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        nbytes = int(_lib.lib().avf_peer_allreduce_bytes(self.world, self.n))
        self.block = symm.empty(nbytes // 4, dtype=torch.float32, device=dev)
        self.block.zero_()
        self.handle = symm.rendezvous(self.block, self.group)
        ptrs = [int(q) for q in self.handle.buffer_ptrs]
        if len(ptrs) != self.world or ptrs[self.rank] != self.block.data_ptr():
            raise RuntimeError("PeerAllReduce: symmetric-memory rendezvous returned an unexpected pointer table")
        self.peer_base = torch.tensor(ptrs, dtype=torch.int64, device=dev)
        self.state = torch.zeros(4, dtype=torch.int32, device=dev)
        self.grad = self.block[:self.n]
        torch.cuda.synchronize()
        dist.barrier(self.group)

    def reduce_(self) -> torch.Tensor:
        c = self._ct.c_void_p
        self._L.check(self._L.lib().avf_grad_allreduce(c(self.peer_base.data_ptr()), self.n, self.world, self.rank, c(self.state.data_ptr()),
                                                       self.timeout_ns, c(torch.cuda.current_stream().cuda_stream)), "avf_grad_allreduce")
        return self.grad

    def reduce_adam_(self, p: torch.Tensor, m: torch.Tensor, v: torch.Tensor, step: int, lr: float, beta1: float, beta2: float, eps: float,
                     weight_decay: float, decoupled: bool = False, grad_scale: Optional[float] = None, shadow: Optional[torch.Tensor] = None) -> None:
        """reduce_() and the Adam / AdamW step on the replicated flat parameter bucket as ONE kernel (avf_adam_allreduce_step):
        bit-identical to reduce_() followed by functional.adam_step_ with grad_scale = 1 / world."""
        if not (p.numel() == m.numel() == v.numel() == self.n and self.n % 4 == 0):
            raise ValueError("PeerAllReduce.reduce_adam_: parameter / moment buckets must have the gradient bucket's length (a multiple of 4)")
        c = self._ct.c_void_p
        gs = (1.0 / self.world) if grad_scale is None else float(grad_scale)
        self._L.check(self._L.lib().avf_adam_allreduce_step(
            c(self.peer_base.data_ptr()), self.n, self.world, self.rank, c(self.state.data_ptr()), self.timeout_ns, c(p.data_ptr()), c(m.data_ptr()),
            c(v.data_ptr()), c(shadow.data_ptr()) if shadow is not None else None, lr, beta1, beta2, eps, weight_decay, int(step), int(decoupled), gs,
            c(torch.cuda.current_stream().cuda_stream)), "avf_adam_allreduce_step")

    def check(self) -> int:
        """Synchronise and return the number of completed reductions.  (A peer that never shows up makes the kernel trap after its
        timeout, which surfaces here or at any earlier synchronisation as a CUDA error; the error word is checked as well.)"""
        torch.cuda.synchronize()
        done, err = (int(v) for v in self.state[:2].tolist())
        if err != 0:
            who, where = ((err - 1), "entry") if err < 100 else ((err - 101), "exit")
            raise RuntimeError(f"PeerAllReduce: rank {who} did not reach the {where} of the reduction within {self.timeout_ns / 1e9:.1f} s")
        return done


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
