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
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch
import torch.distributed as dist

from . import functional as AF


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params: Iterable, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                 decoupled: bool = False, process_group=None):
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1) or not (0 <= betas[1] < 1) or weight_decay < 0:
            raise ValueError("FusedAdam: invalid hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, decoupled=decoupled))
        self.process_group = process_group
        self._buckets: List[Optional[dict]] = [None] * len(self.param_groups)

    # -- bucket construction -------------------------------------------------------------------
    @staticmethod
    def _flatten(params: List[torch.nn.Parameter]) -> dict:
        dev = params[0].device
        offs, n = [], 0
        for p in params:
            if p.dtype != torch.float32 or not p.is_cuda:
                raise TypeError("FusedAdam: parameters must be float32 CUDA tensors (the B200 path has no CPU fallback)")
            offs.append(n)
            n += (p.numel() + 7) // 8 * 8            # 32-byte aligned in the fp32 bucket, 16-byte (TMA) aligned in the bf16 shadow
        flat_p = torch.zeros(n, dtype=torch.float32, device=dev)
        flat_g = torch.zeros(n, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for p, o in zip(params, offs):
                flat_p[o:o + p.numel()].copy_(p.detach().reshape(-1))
                flat_g[o:o + p.numel()].copy_(p.grad.detach().reshape(-1))
                p.data = flat_p[o:o + p.numel()].view(p.shape)
                p.grad = flat_g[o:o + p.numel()].view(p.shape)
        shadow = torch.empty(n, dtype=torch.bfloat16, device=dev)       # bf16 copy of the bucket, refreshed by the update kernel
        return dict(params=params, offs=offs, p=flat_p, g=flat_g, m=torch.zeros_like(flat_p), v=torch.zeros_like(flat_p), shadow=shadow, step=0)

    def zero_grad(self, set_to_none: bool = True) -> None:
        """Bucketed parameters keep their flat gradient views (zeroed with one memset); others follow torch's semantics."""
        for gi, group in enumerate(self.param_groups):
            b = self._buckets[gi]
            bucketed = set()
            if b is not None:
                b["g"].zero_()
                for p, o in zip(b["params"], b["offs"]):
                    view = b["g"][o:o + p.numel()].view(p.shape)
                    if p.grad is None or p.grad.data_ptr() != view.data_ptr():
                        p.grad = view
                    bucketed.add(id(p))
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
        world = dist.get_world_size(self.process_group) if dist.is_initialized() else 1
        for gi, group in enumerate(self.param_groups):
            with_grad = [p for p in group["params"] if p.grad is not None]
            if not with_grad:
                continue
            b = self._buckets[gi]
            if b is None or [id(p) for p in b["params"]] != [id(p) for p in with_grad]:
                if b is not None:
                    raise RuntimeError("FusedAdam: the set of parameters receiving gradients changed after the first step")
                b = self._buckets[gi] = self._flatten(with_grad)
            else:
                for p, o in zip(b["params"], b["offs"]):          # a gradient that autograd re-allocated: pull it into the bucket
                    view = b["g"][o:o + p.numel()].view(p.shape)
                    if p.grad.data_ptr() != view.data_ptr():
                        view.copy_(p.grad)
                        p.grad = view
            if world > 1:
                dist.all_reduce(b["g"], op=dist.ReduceOp.SUM, group=self.process_group)
            b["step"] += 1
            AF.adam_step_(b["p"], b["g"], b["m"], b["v"], b["step"], group["lr"], group["betas"][0], group["betas"][1], group["eps"],
                          group["weight_decay"], decoupled=group["decoupled"], grad_scale=1.0 / world, shadow=b["shadow"])
            # GEMM operand copies for free: weight matrices point at their bf16 image in the shadow bucket, valid for exactly this
            # parameter version (load_state_dict / manual edits bump the version and fall back to the cast kernel)
            for p, o in zip(b["params"], b["offs"]):
                if p.dim() >= 2:
                    p._avf_bf16 = (b["shadow"][o:o + p.numel()].view(p.shape), p._version)
        AF.bump_weights_epoch()
        return loss
