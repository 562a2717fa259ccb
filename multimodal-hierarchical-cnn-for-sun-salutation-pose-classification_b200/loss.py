"""`nn.CrossEntropyLoss()` of the training scripts (Quadtree_from scratch/Quadtree_train.py:44,64) on one libqtcnn kernel:
forward = mean over the batch of logsumexp(logits) - logits[label]; the backward kernel forms (softmax - onehot)/B times the
incoming gradient from the stored logits. fp32 logits [B, C] with C <= 32 and int64 class labels (default arguments of
the reference: no class weights, no label smoothing, mean reduction)."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import capi, ops
from .capi import check, ptr, stream

_counters = {}


def _counter(device):
    key = device.index if device.index is not None else torch.cuda.current_device()
    t = _counters.get(key)
    if t is None:
        t = _counters[key] = torch.zeros(1, device=device, dtype=torch.int32)
    return t


class _CrossEntropyFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels):
        if not logits.is_cuda:
            raise RuntimeError("CrossEntropyLoss: the B200 path has no CPU implementation")
        lg = logits.detach().float().contiguous()
        lab = labels.detach().to(torch.int64).contiguous()
        b, nc = lg.shape
        out = torch.empty(b + 1, device=lg.device, dtype=torch.float32)  # [loss rows | mean]
        check(capi.lib().qt_cross_entropy(ptr(lg), nc, ptr(lab), b, nc, 0.0, None, out.data_ptr(), out.data_ptr() + 4 * b, None,
                                          ptr(_counter(lg.device)), stream()), "cross_entropy")
        ops._count()
        ctx.save_for_backward(lg, lab)
        return out[b]

    @staticmethod
    def backward(ctx, gout):
        lg, lab = ctx.saved_tensors
        b, nc = lg.shape
        dl = torch.empty_like(lg)
        scratch = torch.empty(b + 1, device=lg.device, dtype=torch.float32)
        # gradient of the mean: (softmax - onehot) / B, scaled by the incoming gradient on the device (no host read)
        up = gout.detach().reshape(1).float().contiguous()
        check(capi.lib().qt_cross_entropy(ptr(lg), nc, ptr(lab), b, nc, 1.0 / b, ptr(up), scratch.data_ptr(), scratch.data_ptr() + 4 * b,
                                          ptr(dl), ptr(_counter(lg.device)), stream()), "cross_entropy_bwd")
        ops._count()
        return dl, None


def cross_entropy(logits: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
    return _CrossEntropyFn.apply(logits, labels)


class CrossEntropyLoss(nn.Module):
    """Drop-in for `nn.CrossEntropyLoss()` with the reference's (default) arguments."""

    def forward(self, input, target):  # noqa: A002 - torch's argument names
        return cross_entropy(input, target)
