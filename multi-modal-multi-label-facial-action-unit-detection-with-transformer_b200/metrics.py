"""MultiLabelAccF1 (metrics/accf1.py:45-77) with the counting on the GPU.

The reference gathers every prediction on the host and calls sklearn per AU.  Accuracy and binary F1 only need four counters
per AU (TP, FP, FN, TN over the labelled entries), so ``update`` adds them up on the device (``avf_au_confusion_update``) and
``get`` reads 48 integers — after a sum-all-reduce when several ranks evaluate shards of the data (no logit
gather needed).  Same constructor / update / clear / get as the reference class.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from . import functional as AF


class MultiLabelAccF1:
    def __init__(self, ignore_index=-1, average="binary", device=None):
        if average != "binary":
            raise NotImplementedError("MultiLabelAccF1: the AU metric of the reference uses average='binary'")
        self.ignore_index = ignore_index
        self.average = average
        self.device = torch.device(device) if device is not None else None
        self.counts = None

    def _counts(self, device):
        if self.counts is None:
            self.device = self.device or device
            self.counts = torch.zeros(48, dtype=torch.int64, device=self.device)
        return self.counts

    def _to_dev(self, a):
        if isinstance(a, np.ndarray):
            a = torch.from_numpy(np.ascontiguousarray(a))
        dev = self.device or (a.device if a.is_cuda else torch.device("cuda", torch.cuda.current_device()))
        a = a.to(dev)
        return a if (a.dtype == torch.float32 and a.stride(-1) == 1) else a.float().contiguous()

    def _update(self, pred, y_true, threshold):
        pred, y_true = self._to_dev(pred), self._to_dev(y_true)
        AF._cuda(pred, "y_pred")
        counts = self._counts(pred.device)
        ign = float("nan") if self.ignore_index is None else float(self.ignore_index)
        _lib.check(_lib.lib().avf_au_confusion_update(AF._ptr(pred), pred.stride(0), threshold, AF._ptr(y_true), y_true.stride(0), ign,
                                                      AF._ptr(counts), pred.shape[0], AF._stream()), "au_confusion_update")

    def update(self, y_pred, y_true):
        """y_pred: 0/1 predictions [n,12] as the reference passes them (np.round(sigmoid(logits)), train.py:155) — numpy or tensor."""
        self._update(y_pred, y_true, 0.5)

    def update_from_logits(self, logits, y_true):
        """logits [n, >=12] straight from the model (the [B,21] output works as is): decision = logit > 0."""
        self._update(logits, y_true, 0.0)

    def clear(self):
        if self.counts is not None:
            self.counts.zero_()

    def confusion(self, group=None) -> np.ndarray:
        """[12,4] int64 {TP, FP, FN, TN}, summed over the ranks of ``group`` when torch.distributed is initialised."""
        if self.counts is None:
            return np.zeros((12, 4), dtype=np.int64)
        c = self.counts.clone()
        if dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(c, op=dist.ReduceOp.SUM, group=group)
        return c.cpu().numpy().reshape(12, 4)

    def get(self, group=None):
        c = self.confusion(group).astype(np.float64)
        tp, fp, fn, tn = c[:, 0], c[:, 1], c[:, 2], c[:, 3]
        labeled = c.sum()
        acc = float((tp + tn).sum() / labeled) if labeled > 0 else float("nan")
        denom = 2 * tp + fp + fn
        f1 = np.where(denom > 0, 2 * tp / np.maximum(denom, 1), 0.0)        # sklearn: F1 = 0 when there is nothing to find
        return acc, float(f1.mean())
