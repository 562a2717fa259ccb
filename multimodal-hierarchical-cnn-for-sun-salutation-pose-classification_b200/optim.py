"""Optimizer side of the training loops the reference scripts run, on the kernels of libqtcnn:

  Adam             drop-in for `optim.Adam(model.parameters(), lr=..., weight_decay=...)`
                   (Quadtree_from scratch/Quadtree_train.py:45, 3dcnn/train_3D_Quadtree_cnn_model.py:88): same constructor
                   arguments, `param_groups` (LR schedulers such as ReduceLROnPlateau keep working) and `state_dict` layout
                   (`step`, `exp_avg`, `exp_avg_sq`) as torch.optim.Adam, L2 weight decay folded into the gradient. Every
                   parameter of every group is updated by ONE `qt_adam_multi` launch, which also rewrites the bf16 GEMM
                   operand copies of the conv / linear weights (no separate repack pass after the step).
  clip_grad_norm_  `torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm)` (3dcnn/train_...:123): global L2 norm by a
                   fixed-order two-launch reduction; with `optimizer=` one of the Adam above the clip coefficient is consumed
                   inside the update kernel instead of a separate pass over the gradients.
"""
from __future__ import annotations

import ctypes
import math
from typing import Iterable, Optional

import torch

from . import capi, ops
from .capi import check, ptr, stream


def _pinned_bytes(n):
    return torch.empty(max(n, 1), dtype=torch.uint8, pin_memory=torch.cuda.is_available())


class _DeviceTable:
    """A ctypes array mirrored into device memory; re-uploaded (pinned staging, async) only when its contents change."""

    def __init__(self):
        self.key = None
        self.dev = None
        self.host = None

    def update(self, key, arr, device):
        if key == self.key and self.dev is not None and self.dev.device == device:
            return self.dev
        raw = bytes(arr)
        if self.host is None or self.host.numel() < len(raw):
            self.host = _pinned_bytes(len(raw))
            self.dev = torch.empty(max(len(raw), 1), dtype=torch.uint8, device=device)
        self.host[:len(raw)].copy_(torch.frombuffer(bytearray(raw), dtype=torch.uint8))
        self.dev[:len(raw)].copy_(self.host[:len(raw)], non_blocking=True)
        self.key = key
        return self.dev


class Adam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1) or not (0 <= betas[1] < 1) or weight_decay < 0:
            raise ValueError("Adam: invalid hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay))
        self._table = _DeviceTable()
        self._pending_clip = None  # device scalar set by clip_grad_norm_(..., optimizer=self)
        self.grad_scale = 1.0      # multiplied into every gradient inside the kernel (parallel.DataParallelGrads.attach: 1/world)
        self.launches_last_step = 0

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        todo = []
        eff_groups, eff_index = [], {}
        device = None
        for gi, group in enumerate(self.param_groups):
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda or p.dtype != torch.float32 or p.grad.dtype != torch.float32:
                    raise RuntimeError("qtcnn_b200.optim.Adam: fp32 CUDA parameters only (no CPU implementation)")
                if p.grad.is_sparse:
                    raise RuntimeError("Adam does not support sparse gradients")
                if not p.is_contiguous():
                    raise RuntimeError("qtcnn_b200.optim.Adam: contiguous parameters only")
                device = p.device if device is None else device
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0, dtype=torch.float32)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st["step"] += 1
                t = int(st["step"])
                key = (gi, t, group["lr"], b1, b2, group["eps"], group["weight_decay"])
                if key not in eff_index:
                    eff_index[key] = len(eff_groups)
                    eff_groups.append(key)
                todo.append((p, st, eff_index[key]))
        if not todo:
            return loss
        if len(eff_groups) > 8:
            raise RuntimeError("qtcnn_b200.optim.Adam: more than 8 distinct (group, step) combinations in one step")
        L = capi.lib()
        garr = (capi.AdamGroup * len(eff_groups))()
        for i, (_, t, lr, b1, b2, eps, wd) in enumerate(eff_groups):
            g = garr[i]
            g.step_size = lr / (1.0 - b1 ** t)
            g.beta1, g.beta2, g.eps, g.weight_decay = b1, b2, eps, wd
            g.inv_bc2_sqrt = 1.0 / math.sqrt(1.0 - b2 ** t)
            g.omb1, g.omb2 = 1.0 - b1, 1.0 - b2
        arr = (capi.AdamItem * len(todo))()
        first, max_taps, packed_entries, key = 0, 1, [], []
        for i, (p, st, gidx) in enumerate(todo):
            grad = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
            it = arr[i]
            it.p, it.g, it.m, it.v = p.data_ptr(), grad.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()
            it.n = p.numel()
            it.group = gidx
            e = ops._packed.get(p)
            if e is not None and e.registered and e.wf is not None and e.ptr == p.data_ptr():
                cout, cin, taps = ops._w_dims(p)
                it.wf, it.wd = e.wf.data_ptr(), (e.wd.data_ptr() if e.wd is not None else None)
                it.cout, it.cin, it.taps = cout, cin, taps
                max_taps = max(max_taps, taps)
                packed_entries.append((p, e))
            nb = L.qt_adam_item_plan(ctypes.byref(it))
            if nb < 0:
                check(-1, "adam_item_plan")
            it.first_block = first
            first += nb
            key.append((it.p, it.g, it.m, it.v, it.wf, it.wd, gidx))
        table = self._table.update(tuple(key), arr, device)
        clip = self._pending_clip
        self._pending_clip = None
        check(L.qt_adam_multi(ptr(table), len(todo), first, max_taps, garr, len(eff_groups), ptr(clip), float(self.grad_scale), stream()), "adam_multi")
        ops._count()
        self.launches_last_step = 1
        # the bf16 GEMM copies were rewritten by the same launch: mark them fresh for the generation the optimizer
        # post-step hook is about to start (ops._on_optimizer_step runs after this method returns)
        nxt = ops._generation + 1
        for p, e in packed_entries:
            e.version, e.ptr = (p._version, nxt), p.data_ptr()
        if packed_entries:
            reg = ops._registry(device)
            if all(e.version is not None and e.version[1] == nxt for _, e in reg.items if _() is not None):
                reg.generation = nxt
        return loss


def clip_grad_norm_(parameters: Iterable[torch.Tensor], max_norm: float, norm_type: float = 2.0,
                    optimizer: Optional[Adam] = None) -> torch.Tensor:
    """Global-norm gradient clipping; returns the total norm (a device scalar, like torch's). With `optimizer=` (an
    `Adam` of this module) the gradients are left untouched and the coefficient is applied inside the next
    `optimizer.step()`; otherwise they are scaled in place."""
    if norm_type != 2.0:
        raise ValueError("clip_grad_norm_: only the L2 norm (the reference's default) is implemented")
    if isinstance(parameters, torch.Tensor):
        parameters = [parameters]
    grads = [p.grad for p in parameters if p.grad is not None]
    if not grads:
        return torch.tensor(0.0)
    device = grads[0].device
    L = capi.lib()
    arr = (capi.NormItem * len(grads))()
    first, keep = 0, []
    for i, g in enumerate(grads):
        if not g.is_cuda or g.dtype != torch.float32:
            raise RuntimeError("clip_grad_norm_: fp32 CUDA gradients only")
        gc = g if g.is_contiguous() else g.contiguous()
        keep.append(gc)
        arr[i].g, arr[i].n, arr[i].first_block = gc.data_ptr(), gc.numel(), first
        first += L.qt_grad_norm_blocks(gc.numel())
    tbl = getattr(clip_grad_norm_, "_table", None)
    if tbl is None:
        tbl = clip_grad_norm_._table = _DeviceTable()
    table = tbl.update(tuple((a.g, a.n) for a in arr), arr, device)
    out = torch.empty(2, device=device, dtype=torch.float32)  # [total_norm, coef]
    partial = ops.workspace(4 * first, device, "gradnorm")
    gscale = float(optimizer.grad_scale) if isinstance(optimizer, Adam) else 1.0
    check(L.qt_grad_clip_coef(ptr(table), len(grads), first, float(max_norm), gscale, ptr(partial), out.data_ptr(), out.data_ptr() + 4,
                              stream()), "grad_clip_coef")
    ops._count(2)
    if optimizer is not None:
        if not isinstance(optimizer, Adam):
            raise TypeError("clip_grad_norm_(optimizer=...) expects qtcnn_b200.optim.Adam")
        optimizer._pending_clip = out[1:2]
    else:
        torch._foreach_mul_(grads, out[1])
    return out[0]
