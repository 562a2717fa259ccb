"""Data-parallel gradient synchronisation: one process per GPU, bucketed all-reduce (NCCL over NVLink/NVSwitch)
overlapped with backward.

The reference is single-GPU; QuadtreeCNN training shards by batch with one exchange per step:
the average of the 70 parameter gradients (the never-used `base_cnn.fc` produces none, §0.7). Gradients are
produced directly inside flat per-bucket buffers (ops.grad_out hands the weight-gradient kernels views of
them), buckets follow the order in which backward produces gradients — classifier first (its 58 MB weight
gets a bucket of its own), stem last — and each bucket's all-reduce is issued asynchronously from the
post-accumulate hook of its last member, so NCCL runs beside the remaining backward kernels. `finish()` makes
the compute stream wait for the outstanding collectives before the optimizer reads the gradients.
BatchNorm statistics stay local to each rank (the reference has no SyncBN).
"""
from __future__ import annotations

from typing import Dict, List

import torch
import torch.distributed as dist

from . import ops


class _Bucket:
    def __init__(self, params: List[torch.nn.Parameter]):
        self.params = params
        n = sum(p.numel() for p in params)
        self.flat = torch.zeros(n, device=params[0].device, dtype=params[0].dtype)
        self.views = {}
        off = 0
        for p in params:
            self.views[p] = self.flat[off:off + p.numel()].view(p.shape)
            off += p.numel()
        self.pending = len(params)
        self.work = None


class DataParallelGrads:
    def __init__(self, model: torch.nn.Module, bucket_bytes: int = 25 << 20, broadcast: bool = True, overlap: bool = True):
        if not dist.is_initialized():
            raise RuntimeError("DataParallelGrads needs an initialised process group")
        self.model = model
        self.world = dist.get_world_size()
        self.bucket_bytes = bucket_bytes
        # overlap=True: a bucket's all-reduce starts from the hook of its last gradient and runs beside the rest of backward.
        # overlap=False: all buckets are reduced in finish(), after backward. The persistent conv kernels own every SM (one
        # 223 KB CTA each), so a concurrent NCCL kernel takes SMs away from them and stretches those launches; measured on
        # B200 the exposed all-reduce (NVLink 5: ~0.3 ms for 104 MB) costs less than the interference (profiles/r02_scaling.md).
        self.overlap = overlap
        self.params = []
        seen = set()
        for p in model.parameters():
            if p.requires_grad and id(p) not in seen:
                seen.add(id(p))
                self.params.append(p)
        if broadcast:
            tensors = list({id(t): t for t in list(model.parameters()) + list(model.buffers())}.values())
            for t in tensors:
                dist.broadcast(t.data, src=0)
            # the broadcast wrote through .data (no _version bump): drop bf16 GEMM copies packed before wrapping
            ops.invalidate_packed_weights()
        self.buckets: List[_Bucket] = []
        self.bucket_of: Dict[torch.nn.Parameter, _Bucket] = {}
        self._order: List[torch.nn.Parameter] = []   # production order observed during the first backward
        self._handles = [p.register_post_accumulate_grad_hook(self._hook) for p in self.params]
        self.collectives_last_step = 0
        self.avg = dist.get_backend() == "nccl"  # gloo has no AVG
        self._scaled_by = None  # optimizer that applies the 1/world factor (attach)
        self._sync = True

    # ------------------------------------------------------------------------------------------
    def _build_buckets(self):
        cur, cur_bytes = [], 0
        groups = []
        for p in self._order:
            nbytes = p.numel() * p.element_size()
            if cur and (cur_bytes + nbytes > self.bucket_bytes or cur[0].dtype != p.dtype):
                groups.append(cur)
                cur, cur_bytes = [], 0
            cur.append(p)
            cur_bytes += nbytes
        if cur:
            groups.append(cur)
        self.buckets = [_Bucket(g) for g in groups]
        for b in self.buckets:
            for p in b.params:
                self.bucket_of[p] = b
                ops.register_grad_view(p, b.views[p])

    def attach(self, optimizer):
        """Let `optimizer` (qtcnn_b200.optim.Adam) apply the 1/world factor inside its update kernel: the buckets are then
        all-reduced with SUM — the one fp32 reduction NVSwitch can do in the switch (NVLS; NCCL offers no in-switch AVG) — and
        `p.grad` holds the SUM over ranks between finish() and step()."""
        if not hasattr(optimizer, "grad_scale"):
            raise TypeError("attach() needs an optimizer with a grad_scale attribute (qtcnn_b200.optim.Adam)")
        optimizer.grad_scale = 1.0 / self.world
        self._scaled_by = optimizer
        return self

    def no_sync(self):
        """Context manager for gradient accumulation: backward passes inside it only accumulate into `p.grad` (the
        bucket views); the all-reduce is issued by the first backward outside it, as with DDP.no_sync()."""
        dp = self

        class _NoSync:
            def __enter__(self_inner):
                self_inner.prev, dp._sync = dp._sync, False

            def __exit__(self_inner, *exc):
                dp._sync = self_inner.prev
                return False
        return _NoSync()

    def _hook(self, p):
        b = self.bucket_of.get(p)
        if b is None:  # first step: learn which parameters receive gradients and in which order
            if not any(q is p for q in self._order):
                self._order.append(p)
            return
        v = b.views[p]
        if p.grad.data_ptr() != v.data_ptr():  # autograd cloned instead of adopting our view
            v.copy_(p.grad)
            p.grad = v
        if not self._sync:
            return
        if b.pending == 0:
            raise RuntimeError("DataParallelGrads: a second backward reached an already reduced bucket before finish(); "
                               "wrap all but the last micro-batch backward in `with dp.no_sync():`")
        b.pending -= 1
        if b.pending == 0 and self.overlap:
            b.work = dist.all_reduce(b.flat, op=self._op(), async_op=True)

    def _op(self):
        return dist.ReduceOp.AVG if (self.avg and self._scaled_by is None) else dist.ReduceOp.SUM

    def finish(self):
        """Call after backward(), before optimizer.step()."""
        n = 0
        if not self.buckets:
            # first step: plain per-tensor all-reduce, then freeze the bucket plan
            for p in self._order:
                dist.all_reduce(p.grad, op=dist.ReduceOp.SUM)
                if self._scaled_by is None:
                    p.grad.div_(self.world)
                n += 1
            self._build_buckets()
        else:
            for b in self.buckets:
                if b.pending != 0:
                    raise RuntimeError("DataParallelGrads: a bucket did not fill; the set of parameters receiving "
                                       "gradients changed after the first step")
                if b.work is None:  # overlap=False: every bucket is reduced here, back to back
                    b.work = dist.all_reduce(b.flat, op=self._op(), async_op=True)
            for b in self.buckets:
                b.work.wait()
                if not self.avg and self._scaled_by is None:
                    b.flat.div_(self.world)
                b.work = None
                b.pending = len(b.params)
                n += 1
        self.collectives_last_step = n

    def remove(self):
        for h in self._handles:
            h.remove()
        for p in self.params:
            ops.unregister_grad_view(p)
