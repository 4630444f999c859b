"""Fused Adam over one flat parameter bucket — the optimiser of train.py:334
(``torch.optim.Adam(params=model.parameters(), lr=..., weight_decay=...)``: L2 coupled into the gradient), plus the
decoupled AdamW variant BASELINE config 4 names.

Layout: on the first ``step()`` every parameter that received a gradient is moved into one contiguous fp32 buffer
(``p.data`` becomes a view of it), with matching flat buffers for the gradient (``p.grad`` becomes a view, so later
backward passes accumulate in place), ``exp_avg`` and ``exp_avg_sq``.  A step is then: one sum-all-reduce of the flat
gradient over the data-parallel group (NCCL over NVLink; skipped for a single process) and ONE kernel launch
(``avf_adam_step``) that folds the 1/world_size of the gradient mean into the update.  Parameters that never receive a
gradient (frozen sub-models, the per-modality ``AU_linear_last*`` whose result the AVFormer discards) are left
untouched, exactly like torch.optim.Adam skips ``grad is None``.

Overlap with the backward pass (train.py:235-236 is "backward, then step"; data parallel adds the reduction in between):
``FusedAdam(..., segments=[[params of stack A], [params of stack B], ...])`` lays the bucket out segment by segment, in the
order the stacks FINISH their backward (fusion head first, SFormer last: ``segments_of(model)``).  The autograd bridges
report finished parameter gradients (``autograd.add_grad_listener``); as soon as a segment is complete its slice of the bucket
is all-reduced asynchronously, NCCL running it on its own stream next to the backward kernels of the stacks below, and
``step()`` applies the update segment by segment as the reductions land.  ``finish_reductions()`` joins the outstanding
ones (a CUDA-graph capture of forward + backward calls it before the capture ends, so the collectives are part of the graph).
"""
from __future__ import annotations

import os
import weakref
from typing import Iterable, List, Optional, Sequence

import torch
import torch.distributed as dist

from . import functional as AF
from .dp import SegmentReducer


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params: Iterable, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                 decoupled: bool = False, process_group=None, segments: Optional[Sequence[Sequence[torch.nn.Parameter]]] = None,
                 overlap: Optional[bool] = None, reduce: Optional[str] = None):
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1) or not (0 <= betas[1] < 1) or weight_decay < 0:
            raise ValueError("FusedAdam: invalid hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, decoupled=decoupled))
        self.process_group = process_group
        self._buckets: List[Optional[dict]] = [None] * len(self.param_groups)
        # reduction segments (single parameter group only): parameter id -> segment index, in backward-completion order
        self._seg_of = {}
        if segments is not None:
            if len(self.param_groups) != 1:
                raise ValueError("FusedAdam: reduction segments need a single parameter group")
            for si, seg in enumerate(segments):
                for prm in seg:
                    self._seg_of[id(prm)] = si
        self._n_seg = (max(self._seg_of.values()) + 1) if self._seg_of else 1
        # Off by default: measured on 2 and 8 B200s (bench.py --mode train, 64 clips per GPU) the overlapped reduction is SLOWER than one
        # all-reduce behind the backward (2.42 vs 2.28 ms, 2.59 vs 2.46 ms; one GPU 2.17 ms) — the NCCL CTAs take SMs away from the
        # persistent, statically striding backward GEMMs, whose displaced CTAs finish their tile share late.  AVF_OVERLAP_REDUCE=1 or
        # overlap=True turns it on (the data-parallel parity check, bench.py --check, always exercises it).
        self.overlap = (os.environ.get("AVF_OVERLAP_REDUCE", "0") == "1") if overlap is None else bool(overlap)
        self._reducer: Optional[SegmentReducer] = None
        # how the flat bucket is summed across ranks when the reduction is not overlapped: "peer" = one kernel over NVLink peer memory
        # (dp.PeerAllReduce; the bucket is then allocated from torch symmetric memory), "nccl" = dist.all_reduce.  AVF_GRAD_REDUCE
        # overrides; "peer" falls back to "nccl" (on all ranks together) where no peer mapping is available, see reduce_note.
        self.reduce_mode = os.environ.get("AVF_GRAD_REDUCE", "peer") if reduce is None else str(reduce)
        if self.reduce_mode not in ("peer", "nccl"):
            raise ValueError("FusedAdam: reduce must be 'peer' or 'nccl'")
        self.reduce_note = None
        self._listening = False
        weakself = weakref.ref(self)

        def _on_ready(ps):
            me = weakself()
            if me is not None:
                me._grads_ready(ps)
        self._listener = _on_ready

    # -- bucket construction -------------------------------------------------------------------
    def _flatten(self, params: List[torch.nn.Parameter]) -> dict:
        # segment by segment (stable inside a segment); parameters outside every segment form the last one
        params = sorted(params, key=lambda q: self._seg_of.get(id(q), self._n_seg))
        dev = params[0].device
        offs, n = [], 0
        for p in params:
            if p.dtype != torch.float32 or not p.is_cuda:
                raise TypeError("FusedAdam: parameters must be float32 CUDA tensors (the B200 path has no CPU fallback)")
            offs.append(n)
            n += (p.numel() + 7) // 8 * 8            # 32-byte aligned in the fp32 bucket, 16-byte (TMA) aligned in the bf16 shadow
        flat_p = torch.zeros(n, dtype=torch.float32, device=dev)
        peer = self._peer_bucket(n, dev)
        flat_g = peer.grad if peer is not None else torch.zeros(n, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for p, o in zip(params, offs):
                flat_p[o:o + p.numel()].copy_(p.detach().reshape(-1))
                flat_g[o:o + p.numel()].copy_(p.grad.detach().reshape(-1))
                p.data = flat_p[o:o + p.numel()].view(p.shape)
                p.grad = flat_g[o:o + p.numel()].view(p.shape)
                p._avf_direct_grad = True           # autograd.direct_grad_ok: the backward kernels may accumulate into this view
        shadow = torch.empty(n, dtype=torch.bfloat16, device=dev)       # bf16 copy of the bucket, refreshed by the update kernel
        shadow_views = {id(p): shadow[o:o + p.numel()].view(p.shape) for p, o in zip(params, offs) if p.dim() >= 2}
        bounds, counts = [], []                                          # [lo, hi) of each non-empty segment in the bucket
        for si in range(self._n_seg + 1):
            idx = [i for i, q in enumerate(params) if self._seg_of.get(id(q), self._n_seg) == si]
            if idx:
                lo = offs[idx[0]]
                hi = offs[idx[-1] + 1] if idx[-1] + 1 < len(params) else n
                bounds.append((lo, hi))
                counts.append(len(idx))
        seg_of = {}
        for si, (lo, hi) in enumerate(bounds):                          # re-index to the non-empty segments, bucket members only
            for q, o in zip(params, offs):
                if lo <= o < hi:
                    seg_of[id(q)] = si
        self._seg_of = seg_of
        self._n_seg = len(bounds)
        # host-side caches for step(): per-parameter gradient views / addresses and the bf16 shadow views of the weight matrices
        g_views = [flat_g[o:o + p.numel()].view(p.shape) for p, o in zip(params, offs)]
        b_extra = dict(ids=frozenset(id(p) for p in params), g_views=g_views, g_ptrs=[v.data_ptr() for v in g_views],
                       mats=[(p, shadow_views[id(p)]) for p in params if p.dim() >= 2])
        b = dict(params=params, offs=offs, p=flat_p, g=flat_g, peer=peer, **b_extra, m=torch.zeros_like(flat_p), v=torch.zeros_like(flat_p), shadow=shadow, step=0,
                 bounds=bounds, counts=counts)
        # moments restored by load_state_dict before the bucket existed (torch keeps them in self.state until then)
        with torch.no_grad():
            for p, o in zip(params, offs):
                st = self.state.get(p)
                if st:
                    b["m"][o:o + p.numel()].copy_(st["exp_avg"].reshape(-1))
                    b["v"][o:o + p.numel()].copy_(st["exp_avg_sq"].reshape(-1))
                    b["step"] = max(b["step"], int(st["step"]))
                    self.state.pop(p)
        return b

    def _peer_bucket(self, n: int, dev):
        """The gradient bucket as a peer-mapped block (collective: every rank builds its bucket in its first step())."""
        if self._world() == 1 or self.reduce_mode != "peer":
            return None
        from .dp import PeerAllReduce
        peer = None
        try:
            peer = PeerAllReduce(n, device=dev, group=self.process_group)
        except Exception as ex:
            self.reduce_note = f"peer-memory all-reduce unavailable ({type(ex).__name__}: {str(ex)[:160]}); NCCL all-reduce instead"
        agree = torch.tensor([1 if peer is not None else 0], device=dev)
        dist.all_reduce(agree, op=dist.ReduceOp.MIN, group=self.process_group)
        if int(agree.item()) == 0:
            self.reduce_mode = "nccl"
            return None
        return peer

    # -- checkpointing: torch.optim.Adam's state layout (step / exp_avg / exp_avg_sq per parameter) ------------------------------
    def state_dict(self):
        """The moments live in the flat buckets; expose them per parameter in torch.optim.Adam's format so that a resumed run
        continues with the same exp_avg / exp_avg_sq / bias-correction step (train.py saves only the model, but a runner that
        checkpoints the optimiser must not silently restart the moments)."""
        for b in self._buckets:
            if b is None:
                continue
            for p, o in zip(b["params"], b["offs"]):
                n = p.numel()
                self.state[p] = {"step": torch.tensor(float(b["step"])), "exp_avg": b["m"][o:o + n].view(p.shape).clone(),
                                 "exp_avg_sq": b["v"][o:o + n].view(p.shape).clone()}
        try:
            return super().state_dict()
        finally:
            for b in self._buckets:
                if b is not None:
                    for p in b["params"]:
                        self.state.pop(p, None)

    def load_state_dict(self, state_dict) -> None:
        super().load_state_dict(state_dict)          # fills self.state[param] with tensors on the parameter's device
        with torch.no_grad():
            for b in self._buckets:
                if b is None:
                    continue
                for p, o in zip(b["params"], b["offs"]):
                    st = self.state.pop(p, None)
                    if st:
                        n = p.numel()
                        b["m"][o:o + n].copy_(st["exp_avg"].reshape(-1))
                        b["v"][o:o + n].copy_(st["exp_avg_sq"].reshape(-1))
                        b["step"] = int(st["step"])

    # -- overlapped gradient reduction (dp.SegmentReducer does the bookkeeping) ----------------------
    def _world(self) -> int:
        return dist.get_world_size(self.process_group) if dist.is_initialized() else 1

    def _arm(self) -> None:
        """Start of a backward pass: every segment waits for all of its parameters again."""
        b = self._buckets[0] if len(self._buckets) == 1 else None
        if b is None or not self.overlap or self._world() == 1:
            self._reducer = None
            return
        if b.get("reducer") is None:
            b["reducer"] = SegmentReducer(b["g"], b["bounds"], b["counts"], self._seg_of, self.process_group)
        self._reducer = b["reducer"]
        self._reducer.arm()
        if not self._listening:
            from . import autograd as AG
            AG.add_grad_listener(self._listener)
            self._listening = True

    def _grads_ready(self, params) -> None:
        """Called by the autograd bridges once the kernels that accumulate these parameters' gradients are enqueued."""
        if self._reducer is not None:
            self._reducer.ready(params)

    def finish_reductions(self) -> None:
        """Join every outstanding segment reduction into the current stream (segments whose completion was never reported —
        a stack whose inputs need no gradient reports nothing — are reduced here)."""
        if self._reducer is not None:
            self._reducer.finish()

    def zero_grad(self, set_to_none: bool = True) -> None:
        """Bucketed parameters keep their flat gradient views (zeroed with one memset); others follow torch's semantics."""
        self._zero_grad(set_to_none)
        self._arm()

    def _zero_grad(self, set_to_none: bool = True) -> None:
        for gi, group in enumerate(self.param_groups):
            b = self._buckets[gi]
            bucketed = frozenset()
            if b is not None:
                b["g"].zero_()
                for p, view, ptr in zip(b["params"], b["g_views"], b["g_ptrs"]):
                    g = p.grad
                    if g is None or g.data_ptr() != ptr:
                        p.grad = view
                bucketed = b["ids"]
            for p in group["params"]:
                if id(p) in bucketed or p.grad is None:
                    continue
                if set_to_none:
                    p.grad = None
                else:
                    p.grad.detach_().zero_()

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        world = self._world()
        for gi, group in enumerate(self.param_groups):
            b = self._buckets[gi]
            if b is None:
                with_grad = [p for p in group["params"] if p.grad is not None]
                if not with_grad:
                    continue
                b = self._buckets[gi] = self._flatten(with_grad)
            else:
                # steady state: this runs on the host in front of every update, so no tensor is created here (cached views / addresses)
                ids, n_with = b["ids"], 0
                for p in group["params"]:
                    if p.grad is not None:
                        n_with += 1
                        if id(p) not in ids:
                            raise RuntimeError("FusedAdam: the set of parameters receiving gradients changed after the first step")
                if n_with == 0:
                    continue
                if n_with != len(ids):
                    raise RuntimeError("FusedAdam: the set of parameters receiving gradients changed after the first step")
                for p, view, ptr in zip(b["params"], b["g_views"], b["g_ptrs"]):   # a gradient that autograd re-allocated: pull it into the bucket
                    if p.grad.data_ptr() != ptr:
                        view.copy_(p.grad)
                        p.grad = view
            if world > 1 and self._reducer is not None and self._reducer.armed:
                self._reducer.finish()                            # segments were armed for this backward pass: reduce / join what is left
                self._reducer.disarm()
                fused = False
            elif world > 1 and b.get("peer") is not None:
                fused = True                                      # the sum over ranks AND the update: one kernel over NVLink peer memory
            elif world > 1:
                dist.all_reduce(b["g"], op=dist.ReduceOp.SUM, group=self.process_group)
                fused = False
            else:
                fused = False
            b["step"] += 1
            if fused:
                b["peer"].reduce_adam_(b["p"], b["m"], b["v"], b["step"], group["lr"], group["betas"][0], group["betas"][1], group["eps"],
                                       group["weight_decay"], decoupled=group["decoupled"], grad_scale=1.0 / world, shadow=b["shadow"])
            else:
                AF.adam_step_(b["p"], b["g"], b["m"], b["v"], b["step"], group["lr"], group["betas"][0], group["betas"][1], group["eps"],
                              group["weight_decay"], decoupled=group["decoupled"], grad_scale=1.0 / world, shadow=b["shadow"])
            # GEMM operand copies for free: weight matrices point at their bf16 image in the shadow bucket, valid for exactly this
            # parameter version (load_state_dict / manual edits bump the version and fall back to the cast kernel)
            for p, sview in b["mats"]:
                p._avf_bf16 = (sview, p._version)
            for p in b["params"]:                               # per-parameter staleness: packed copies of OTHER (frozen) parameters stay valid
                p._avf_wver = getattr(p, "_avf_wver", 0) + 1
        return loss


def segments_of(model) -> List[List[torch.nn.Parameter]]:
    """The AVFormer's trainable hot-path parameters grouped by stack, in the order the stacks finish their backward pass
    (models/avformer.py:93-106 read backwards): fusion head, video AU_former, audio AU_former, TFormer, SFormer.  Everything else
    (conv backbones) forms a last segment."""
    groups = [model.au_head, model.video_model.au_head, model.audio_model.au_head, model.video_model.video_model.t_former]
    segs = [[q for q in g.parameters() if q.requires_grad] for g in groups]
    sf = model.video_model.video_model.s_former
    segs.append([q for n, q in sf.named_parameters() if q.requires_grad and (n.startswith("spatial_transformer.") or n == "pos_embedding")])
    return segs
