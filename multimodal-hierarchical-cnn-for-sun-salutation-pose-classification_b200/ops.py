"""Tensor-level wrappers over the C ABI: allocate outputs with torch, pass raw pointers, return tensors.

Activations are dense channels-last bf16 buffers ([N,H,W,C] / [N,D,H,W,C]); statistics and weight gradients
are fp32. Nothing here computes on the host or falls back to ATen kernels for the hot ops.
"""
from __future__ import annotations

import ctypes
import weakref
from typing import Optional, Tuple

import torch

from . import capi
from .capi import check, ptr, stream

BF16 = torch.bfloat16
_launches = 0  # number of libqtcnn kernel-launching calls (bench.py reports it)


def _count(n=1):
    global _launches
    _launches += n


def launches() -> int:
    return _launches


def L():
    return capi.lib()


# ------------------------------------------------------------------------------------------------- workspace
_ws = {}


def workspace(nbytes: int, device, tag="gemm") -> torch.Tensor:
    key = (device.index if device.index is not None else torch.cuda.current_device(), tag)
    t = _ws.get(key)
    if t is None or t.numel() < nbytes:
        t = torch.empty(max(nbytes, 1 << 20), device=device, dtype=torch.uint8)
        _ws[key] = t
    return t


# ------------------------------------------------------------------------------------------------- weights
# Fused / foreach optimizers update parameters without bumping Tensor._version, so packed bf16 weight copies
# are also invalidated by a global optimizer-step counter (hooked once, covers every torch.optim optimizer).
_generation = 0
_hard_epoch = 0  # explicit invalidations: these also drop the packs of frozen parameters


def _on_optimizer_step(*_args, **_kw):
    global _generation
    _generation += 1


def _gen_of(w) -> int:
    """Cache generation a parameter's packs belong to. A frozen parameter (requires_grad False) is never touched by an
    optimizer step — ours writes through raw pointers only where a gradient exists — so its packs survive steps (19 repack
    launches per step on the frozen backbones otherwise); in-place torch edits still bump `_version`."""
    return _generation if w.requires_grad else -1 - _hard_epoch


def invalidate_packed_weights():
    """Call after modifying parameters through a path that neither bumps `_version` nor is a torch optimizer."""
    global _hard_epoch
    _hard_epoch += 1
    _on_optimizer_step()


try:
    from torch.optim.optimizer import register_optimizer_step_post_hook as _reg_post_hook
    _reg_post_hook(_on_optimizer_step)
except Exception as _e:  # pragma: no cover
    raise RuntimeError("this torch build lacks register_optimizer_step_post_hook; weight caches cannot be kept fresh") from _e


class _Packed:
    __slots__ = ("version", "ptr", "wf", "wd", "w8", "registered", "__weakref__")

    def __init__(self):
        self.version = None
        self.ptr = 0
        self.wf = self.wd = self.w8 = None
        self.registered = False


class _IdMap:
    """Tensor-keyed weak map (WeakKeyDictionary cannot be used: it compares keys with ==, which is elementwise
    for tensors). Entries die with their tensor."""

    def __init__(self):
        self._d = {}

    def get(self, t, default=None):
        item = self._d.get(id(t))
        if item is None or item[0]() is not t:
            return default
        return item[1]

    def set(self, t, value):
        key = id(t)
        self._d[key] = (weakref.ref(t, lambda _r, k=key, d=self._d: d.pop(k, None)), value)

    def pop(self, t, default=None):
        item = self._d.pop(id(t), None)
        return default if item is None else item[1]


_packed = _IdMap()


class _PackRegistry:
    """All GEMM weights of one device that have been packed in both layouts. After an optimizer step the first
    conv that asks for a weight repacks EVERY registered weight with one `qt_wpack_multi` launch (in place: the
    bf16 buffers and therefore the device table stay valid across steps)."""

    def __init__(self, device):
        self.device = device
        self.items = []          # (weakref(param), entry)
        self.table = None        # uint8 device tensor holding qt_wpack_item[]
        self.table_ptrs = None
        self.blocks = 0
        self.max_taps = 1
        self.generation = -1     # generation the registered weights were last packed at

    def add(self, w, e):
        if not any(x[1] is e for x in self.items):  # an entry whose buffers were dropped and re-created is already listed
            self.items.append((weakref.ref(w), e))
        e.registered = True
        self.table = None

    def _rebuild(self):
        live = []
        for ref, e in self.items:
            w = ref()
            if w is not None and e.wf is not None and _packed.get(w) is e:
                live.append((ref, e))
            else:
                e.registered = False
        self.items = live
        arr = (capi.WpackItem * max(len(live), 1))()
        first, max_taps, ptrs = 0, 1, []
        for i, (ref, e) in enumerate(live):
            w = ref()
            cout, cin, taps = _w_dims(w)
            it = arr[i]
            it.w, it.wf, it.wd = w.data_ptr(), e.wf.data_ptr(), (e.wd.data_ptr() if e.wd is not None else None)
            it.cout, it.cin, it.taps = cout, cin, taps
            nb = L().qt_wpack_item_plan(ctypes.byref(it))
            if nb < 0:
                check(-1, "wpack_item_plan")
            it.first_block = first
            first += nb
            max_taps = max(max_taps, taps)
            ptrs.append((w.data_ptr(), e.wf.data_ptr(), e.wd.data_ptr() if e.wd is not None else 0))
        raw = bytes(arr)[: ctypes.sizeof(capi.WpackItem) * len(live)] if live else b""
        self.table = torch.frombuffer(bytearray(raw or b"\0"), dtype=torch.uint8).to(self.device)
        self.table_ptrs, self.blocks, self.max_taps = ptrs, first, max_taps

    def repack_all(self):
        if self.table is not None:
            cur = []
            for ref, e in self.items:
                w = ref()
                if w is None or e.wf is None:
                    cur = None
                    break
                cur.append((w.data_ptr(), e.wf.data_ptr(), e.wd.data_ptr() if e.wd is not None else 0))
            if cur != self.table_ptrs:
                self.table = None
        if self.table is None:
            self._rebuild()
        if self.items:
            check(L().qt_wpack_multi(ptr(self.table), len(self.items), self.blocks, self.max_taps, stream()), "wpack_multi")
            _count()
        self.generation = _generation
        for ref, e in self.items:
            w = ref()
            e.version, e.ptr = (w._version, _gen_of(w)), w.data_ptr()


_registries = {}


def _registry(device) -> _PackRegistry:
    key = device.index if device.index is not None else torch.cuda.current_device()
    r = _registries.get(key)
    if r is None:
        r = _registries[key] = _PackRegistry(device)
    return r


def _entry(w: torch.Tensor) -> _Packed:
    e = _packed.get(w)
    if e is None:
        e = _Packed()
        _packed.set(w, e)
    gen = _gen_of(w)
    if e.registered and e.version is not None and e.version[1] != gen and e.ptr == w.data_ptr():
        _registry(w.device).repack_all()  # optimizer stepped: every registered weight in one launch
    ver = (w._version, gen)
    if e.version != ver or e.ptr != w.data_ptr():
        e.version, e.ptr = ver, w.data_ptr()
        e.wf = e.wd = e.w8 = None
        e.registered = False
    return e


def _w_dims(w: torch.Tensor) -> Tuple[int, int, int]:
    cout, cin = w.shape[0], w.shape[1]
    taps = 1
    for s in w.shape[2:]:
        taps *= s
    return cout, cin, taps


def packed_fprop(w: torch.Tensor) -> torch.Tensor:
    """bf16 [cout][taps][cin] copy of an fp32 parameter, refreshed when the parameter changes. When autograd is
    recording, the data-gradient layout [cin][taps][cout] is produced by the same kernel (one read of w)."""
    e = _entry(w)
    if e.wf is None:
        cout, cin, taps = _w_dims(w)
        e.wf = torch.empty(cout, taps, cin, device=w.device, dtype=BF16)
        wc = w.detach().contiguous()
        if torch.is_grad_enabled() or w.requires_grad:
            e.wd = torch.empty(cin, taps, cout, device=w.device, dtype=BF16)
        check(L().qt_wpack_both(ptr(wc), ptr(e.wf), ptr(e.wd), cout, cin, taps, stream()), "wpack_both")
        _count()
        if e.wd is not None and w.is_contiguous() and w.dtype == torch.float32 and w.is_cuda and taps <= 32:
            _registry(w.device).add(w, e)
    return e.wf


def packed_dgrad(w: torch.Tensor) -> torch.Tensor:
    """bf16 [cin][taps][cout] copy (B operand of the data-gradient GEMM)."""
    e = _entry(w)
    if e.wd is None:
        cout, cin, taps = _w_dims(w)
        e.wd = torch.empty(cin, taps, cout, device=w.device, dtype=BF16)
        wc = w.detach().contiguous()
        check(L().qt_wpack_dgrad(ptr(wc), ptr(e.wd), cout, cin, taps, stream()), "wpack_dgrad")
        _count()
    return e.wd


def packed_pair(w: torch.Tensor) -> torch.Tensor:
    """bf16 [cout][2][9][64] pair layout of a 32-input-channel Conv3d weight (slab kernel, two depth planes per slab row);
    kept in the entry's `w8` slot, re-made lazily after the parameter changes."""
    e = _entry(w)
    if e.w8 is None:
        cout = w.shape[0]
        if tuple(w.shape[1:]) != (32, 3, 3, 3):
            raise RuntimeError("packed_pair: expects a [cout, 32, 3, 3, 3] Conv3d weight")
        e.w8 = torch.empty(cout, 2, 9, 64, device=w.device, dtype=BF16)
        check(L().qt_wpack_conv3d_pair(ptr(w.detach().contiguous()), ptr(e.w8), cout, stream()), "wpack_conv3d_pair")
        _count()
    return e.w8


def packed_stem(w: torch.Tensor) -> torch.Tensor:
    e = _entry(w)
    if e.w8 is None:
        cout, cin, r, s = w.shape
        e.w8 = torch.empty(cout, 8, 32, device=w.device, dtype=BF16)
        wc = w.detach().contiguous()
        check(L().qt_wpack_stem(ptr(wc), ptr(e.w8), cout, cin, r, s, stream()), "wpack_stem")
        _count()
    return e.w8


# ------------------------------------------------------------------------------------------------- conv
_desc_cache = {}


def conv2d_desc(n, h, w, cin, cout, k, stride, pad):
    key = ("2d", n, h, w, cin, cout, k, stride, pad)
    d = _desc_cache.get(key)
    if d is None:
        d = capi.conv_desc(n, (1, h, w), cin, cout, (1, k, k), (1, stride, stride), (0, pad, pad))
        _desc_cache[key] = d
    return d


def quadrant_desc(n, h, w, cin, cout, k, pad):
    """Four quadrant views (TL, TR, BL, BR) of a dense [n,h,w,cin] map as one grouped conv; output is
    quadrant-major [4][n][h/2][w/2][cout]. Even h, w only (the reference's 14x14 / 28x28 maps)."""
    key = ("quad", n, h, w, cin, cout, k, pad)
    d = _desc_cache.get(key)
    if d is None:
        if h % 2 or w % 2:
            raise RuntimeError("quadrant_desc: even feature-map sizes only")
        qh, qw = h // 2, w // 2
        xs = (h * w * cin, 0, w * cin, cin)
        xoff = (0, qw * cin, qh * w * cin, qh * w * cin + qw * cin)
        yoff = tuple(q * n * qh * qw * cout for q in range(4))
        d = capi.conv_desc(n, (1, qh, qw), cin, cout, (1, k, k), (1, 1, 1), (0, pad, pad), x_stride=xs, groups=4,
                           x_group_off=xoff, y_group_off=yoff)
        _desc_cache[key] = d
    return d


def _fam(d, which, name):
    """Profiling family: the kernel the C dispatch will pick for this pass + the pass name."""
    if _prof is None:
        return name
    slab = L().qt_conv_plan(d, which) >= 1
    kern = ("wgrad3x3_kernel" if which == 2 else "conv3x3_kernel") if slab else ("igemm_wgrad_kernel" if which == 2 else "igemm_kmajor_kernel")
    return f"{kern}{'[conv3d]' if d.k_d == 3 else ''}:{name}"


def conv_out_hw(d) -> Tuple[int, int, int]:
    return (capi.out_size(d.in_d, d.k_d, d.stride_d, d.pad_d), capi.out_size(d.in_h, d.k_h, d.stride_h, d.pad_h),
            capi.out_size(d.in_w, d.k_w, d.stride_w, d.pad_w))


def conv_fprop(d, x, wf, y, bias=None, relu=False, want_stats=False):
    """y = conv(x) [+bias][ReLU]; returns the BatchNorm partial-sum buffer when want_stats."""
    stats = None
    flags = (capi.QT_EPI_BIAS if bias is not None else 0) | (capi.QT_EPI_RELU if relu else 0)
    if want_stats:
        rows = L().qt_conv_stat_rows(d)
        stats = torch.empty(rows, 2, d.out_c, device=x.device, dtype=torch.float32)
        flags |= capi.QT_EPI_STATS
    with gemm_scope(_fam(d, 0, "fprop"), conv_flops(d) if _prof is not None else 0.0):
        check(L().qt_conv_fprop(d, ptr(x), ptr(wf), ptr(y), ptr(bias), ptr(stats), flags, None, 0, stream()), "conv_fprop")
    _count()
    return stats


def conv_dgrad(d, dy, wd, dx, accumulate=False):
    with gemm_scope(_fam(d, 1, "dgrad"), conv_flops(d) if _prof is not None else 0.0):
        check(L().qt_conv_dgrad(d, ptr(dy), ptr(wd), ptr(dx), 1 if accumulate else 0, stream()), "conv_dgrad")
    _count(d.stride_h * d.stride_w * d.stride_d)


def conv_wgrad(d, x, dy, dw, accumulate=False):
    nbytes = L().qt_conv_wgrad_workspace_bytes(d)
    ws = workspace(nbytes, x.device)
    with gemm_scope(_fam(d, 2, "wgrad"), conv_flops(d) if _prof is not None else 0.0):
        check(L().qt_conv_wgrad(d, ptr(x), ptr(dy), ptr(dw), 1 if accumulate else 0, ptr(ws), ws.numel(), stream()),
              "conv_wgrad")
    _count(2)


# ------------------------------------------------------------------------------------------------- batch norm
class BNState:
    """Per-call statistics of one train-mode BatchNorm (saved for backward)."""
    __slots__ = ("mean", "invstd", "scale", "shift")

    def __init__(self, c, device):
        buf = torch.empty(4, c, device=device, dtype=torch.float32)
        self.mean, self.invstd, self.scale, self.shift = buf[0], buf[1], buf[2], buf[3]


def bn_finalize(stats, count, bn, c, device, training=True) -> BNState:
    """Train: batch statistics from the conv epilogue's partial sums (+ running-stat update, as
    nn.BatchNorm does). Eval: coefficients from the running statistics."""
    st = BNState(c, device)
    gamma = bn.weight.detach() if bn.weight is not None else None
    beta = bn.bias.detach() if bn.bias is not None else None
    if training:
        ws = workspace(L().qt_bn_workspace_bytes(c), device, "bn")
        track = bn.track_running_stats and bn.running_mean is not None
        momentum = 0.1 if bn.momentum is None else float(bn.momentum)
        nbt = bn.num_batches_tracked if (track and bn.num_batches_tracked is not None and bn.num_batches_tracked.is_cuda
                                         and bn.num_batches_tracked.dtype == torch.int64) else None
        check(L().qt_bn_finalize_tracked(ptr(stats), stats.shape[0], c, float(count), ptr(gamma), ptr(beta), float(bn.eps), momentum,
                                         ptr(bn.running_mean) if track else None, ptr(bn.running_var) if track else None, ptr(nbt),
                                         ptr(st.mean), ptr(st.invstd), ptr(st.scale), ptr(st.shift), ptr(ws), ws.numel(), stream()),
              "bn_finalize")
        _count(2)
        if nbt is None and track and bn.num_batches_tracked is not None:
            bn.num_batches_tracked.add_(1)
    else:
        check(L().qt_bn_eval_coeffs(c, ptr(gamma), ptr(beta), ptr(bn.running_mean), ptr(bn.running_var), float(bn.eps),
                                    ptr(st.mean), ptr(st.invstd), ptr(st.scale), ptr(st.shift), stream()), "bn_eval_coeffs")
        _count()
    return st


def bn_apply(y, st: BNState, out, residual=None, relu=True):
    c = y.shape[-1]
    m = y.numel() // c
    # algorithmic bytes: y in, (residual in,) out — bf16
    with gemm_scope("bn_apply", 0.0, 2.0 * y.numel() * (3 if residual is not None else 2)):
        check(L().qt_bn_apply(ptr(y), ptr(st.scale), ptr(st.shift), ptr(residual), ptr(out), m, c, 1 if relu else 0, stream()),
              "bn_apply")
    _count()
    return out


def bn_backward(dout, act, y, st: BNState, gamma, dgamma, dbeta, dy, dz_out=None, eval_mode=False, mask_from_y=False):
    """act: stored post-ReLU output (mask act > 0), or None. mask_from_y=True (with act=None) recomputes the ReLU mask
    from y with this BatchNorm's own scale/shift — valid when the layer is BN -> ReLU with no residual add."""
    c = y.shape[-1]
    m = y.numel() // c
    ws = workspace(L().qt_bn_workspace_bytes(c), y.device, "bn")
    msc, msh = (ptr(st.scale), ptr(st.shift)) if (mask_from_y and act is None) else (None, None)
    # algorithmic bytes of the two passes: reduce reads dout, y (, act); apply reads them again and writes dy (, dz)
    a = 1 if act is not None else 0
    with gemm_scope("bn_backward", 0.0, 2.0 * y.numel() * ((2 + a) + (2 + a) + 1 + (1 if dz_out is not None else 0))):
        check(L().qt_bn_backward(ptr(dout), ptr(act), ptr(y), ptr(st.mean), ptr(st.invstd), ptr(gamma), msc, msh, m, c, ptr(dgamma),
                                 ptr(dbeta), 0, 1 if eval_mode else 0, ptr(dy), ptr(dz_out), ptr(ws), ws.numel(), stream()),
              "bn_backward")
    _count(4)


def colsum(x2d, out, accumulate=False):
    m, c = x2d.shape
    ws = workspace(L().qt_bn_workspace_bytes(c), x2d.device, "bn")
    check(L().qt_colsum(ptr(x2d), m, c, ptr(out), 1 if accumulate else 0, ptr(ws), ws.numel(), stream()), "colsum")
    _count(3)


# ------------------------------------------------------------------------------------------------- input
_norm_cache = {}


def input_norm(device, mean=None, std=None):
    """Per-channel (scale, shift) device tensors of ToTensor + Normalize for uint8 pixels: v/255 - mean over std."""
    from .data import IMAGENET_MEAN, IMAGENET_STD
    mean = tuple(IMAGENET_MEAN if mean is None else mean)
    std = tuple(IMAGENET_STD if std is None else std)
    key = (device.index if device.index is not None else torch.cuda.current_device(), mean, std)
    t = _norm_cache.get(key)
    if t is None:
        scale = torch.tensor([1.0 / (255.0 * s) for s in std] + [0.0], dtype=torch.float32)
        shift = torch.tensor([-m / s for m, s in zip(mean, std)] + [0.0], dtype=torch.float32)
        t = _norm_cache[key] = (scale.to(device), shift.to(device))
    return t


def stem_source(x: torch.Tensor):
    """(contiguous NCHW tensor, QT_DTYPE_*, scale, shift) for the input-packing kernels: fp32 / bf16 inputs are taken as
    already normalised (what the reference's DataLoader yields), uint8 inputs are decoded pixels normalised on the device."""
    xd = x.detach()
    if xd.dtype == torch.uint8:
        scale, shift = input_norm(xd.device)
        return xd.contiguous(), capi.QT_DTYPE_U8, scale, shift
    if xd.dtype == BF16:
        return xd.contiguous(), capi.QT_DTYPE_BF16, None, None
    return xd.float().contiguous(), capi.QT_DTYPE_F32, None, None


# ------------------------------------------------------------------------------------------------- misc
def as_nhwc(t: torch.Tensor) -> torch.Tensor:
    """Logical NCHW tensor (any dtype / memory format) -> dense [N,H,W,C] bf16 buffer (no copy when it already
    is a channels-last bf16 tensor produced by this package)."""
    if t.dtype != BF16:
        t = t.to(BF16)
    b = t.permute(0, 2, 3, 1)
    return b if b.is_contiguous() else b.contiguous()


def as_channels_last(t: torch.Tensor) -> torch.Tensor:
    """Logical NCHW / NCDHW tensor -> dense channels-last bf16 buffer [N,H,W,C] / [N,D,H,W,C] (no copy when it already is a
    channels-last bf16 tensor produced by this package)."""
    if t.dim() == 4:
        return as_nhwc(t)
    if t.dtype != BF16:
        t = t.to(BF16)
    b = t.permute(0, 2, 3, 4, 1)
    return b if b.is_contiguous() else b.contiguous()


def as_channels_first_view(buf: torch.Tensor) -> torch.Tensor:
    """Dense channels-last buffer -> logical NCHW / NCDHW view."""
    return buf.permute(0, 3, 1, 2) if buf.dim() == 4 else buf.permute(0, 4, 1, 2, 3)


def conv_nd_desc(n, dhw, cin, cout, k, stride, pad):
    key = ("nd", n, tuple(dhw), cin, cout, tuple(k), tuple(stride), tuple(pad))
    d = _desc_cache.get(key)
    if d is None:
        d = capi.conv_desc(n, tuple(dhw), cin, cout, tuple(k), tuple(stride), tuple(pad))
        _desc_cache[key] = d
    return d


def as_nchw_view(buf: torch.Tensor) -> torch.Tensor:
    """Dense [N,H,W,C] buffer -> logical NCHW (channels_last) view."""
    return buf.permute(0, 3, 1, 2)


_rank_salt = [None]


def new_seed() -> int:
    """Dropout seed drawn from torch's CPU generator (so torch.manual_seed makes runs reproducible; a host-side draw,
    no device synchronisation), mixed with the data-parallel rank so that ranks seeded identically still apply
    independent masks to their shards (as per-device Philox streams do under DDP)."""
    base = int(torch.randint(0, 2 ** 62, (1,)).item())
    if _rank_salt[0] is None:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            _rank_salt[0] = (dist.get_rank() * 0x9E3779B97F4A7C15) & ((1 << 62) - 1)
        else:
            return base  # not cached: a process group may still be created later
    return base ^ _rank_salt[0]


# ------------------------------------------------------------------------------------------------- param grads
_grad_views = _IdMap()


def register_grad_view(p, view):
    """parallel.DataParallelGrads: weight-gradient kernels write straight into the flat bucket buffers."""
    _grad_views.set(p, view)


def unregister_grad_view(p):
    _grad_views.pop(p, None)


def grad_out(p: torch.Tensor) -> torch.Tensor:
    """Destination tensor for the gradient of parameter `p`: the bucket view under data parallelism, but only while
    `p.grad` is None (autograd then adopts the view as `p.grad`). When a gradient already exists (zero_grad with
    set_to_none=False, micro-batch accumulation) a fresh tensor is returned, so autograd's `p.grad += new` adds into
    the view instead of a kernel overwriting it and the sum doubling."""
    v = _grad_views.get(p)
    if v is not None and getattr(p, "grad", None) is None:
        # a FRESH tensor object over the bucket storage: autograd's AccumulateGrad only adopts ("steals") a gradient whose
        # TensorImpl nobody else references; handing out the registered view object itself made it clone the gradient and
        # the hook copy it back — two device copies per parameter per step (round 1's hidden data-parallel overhead)
        return v.detach()
    return torch.empty_like(p, memory_format=torch.contiguous_format)


# ------------------------------------------------------------------------------------------------- profiling
_prof = None


def profile_begin():
    """Start recording one CUDA-event pair per GEMM-family launch (bench.py's roofline pass)."""
    global _prof
    _prof = []


def profile_end():
    global _prof
    rec, _prof = _prof, None
    torch.cuda.synchronize()
    out = {}
    for fam, flops, nbytes, e0, e1 in rec:
        d = out.setdefault(fam, {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "n": 0})
        d["ms"] += e0.elapsed_time(e1)
        d["flops"] += flops
        d["bytes"] += nbytes
        d["n"] += 1
    return out


def profiling() -> bool:
    return _prof is not None


class gemm_scope:
    """`with gemm_scope(family, flops[, nbytes]): <launch>` — a no-op unless profiling is on. Tensor-core families carry
    their algorithmic FLOPs, streaming (HBM-bound) families their algorithmic bytes."""
    __slots__ = ("fam", "flops", "nbytes", "e0")

    def __init__(self, fam, flops, nbytes=0.0):
        self.fam, self.flops, self.nbytes, self.e0 = fam, flops, nbytes, None

    def __enter__(self):
        if _prof is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if self.e0 is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            _prof.append((self.fam, self.flops, self.nbytes, self.e0, e1))
        return False


def conv_flops(d) -> float:
    """Algorithmic FLOPs (2*MAC) of one convolution pass described by `d` (same for fprop, dgrad, wgrad)."""
    od, oh, ow = conv_out_hw(d)
    return 2.0 * d.groups * d.n * od * oh * ow * d.k_d * d.k_h * d.k_w * d.in_c * d.out_c
