"""AULoss — positive-weighted multi-label BCE-with-logits (models/loss.py:63-103) as one CUDA kernel
(valid-row selection, stable softplus form, mean) with its closed-form gradient."""
from __future__ import annotations

import torch
from torch import nn

from . import functional as AF

AU_POS_WEIGHT = (1., 1., 1., 1., 1., 1., 1., 3., 3., 3., 1., 2.)


class _AUBCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y_pred, y_true, pos_weight):
        loss, _, grad = AF.au_bce_loss(y_pred, y_true, pos_weight, want_grad=y_pred.requires_grad)
        ctx.save_for_backward(grad)
        ctx.shape = y_pred.shape
        return loss.clone()

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return (grad * g).to(grad.dtype).view(ctx.shape), None, None


class _LossFn(nn.Module):
    """Holds the ``pos_weight`` buffer under the reference's key ``loss_AU.loss_fn.pos_weight``."""

    def __init__(self):
        super().__init__()
        self.register_buffer("pos_weight", torch.tensor(AU_POS_WEIGHT))


class AULoss(nn.Module):
    def __init__(self, ignore=-1):
        super().__init__()
        if ignore != -1:
            raise NotImplementedError("AULoss: the kernel implements the reference's ignore value -1")
        self.ignore = ignore
        self.loss_fn = _LossFn()

    def forward(self, y_pred, y_true):
        """y_pred [N,12] logits, y_true [N,12] in {0,1} (or -1 in column 0 to drop the row) -> scalar."""
        return _AUBCE.apply(y_pred, y_true, self.loss_fn.pos_weight)
