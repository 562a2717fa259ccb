"""Host-side mirror of the reference's `models.py` interface, executing on the sm_100a kernels of libqtcnn.

Same class names, constructor signatures, `forward(image_input, numerical_input) -> fp32 logits`, module tree
and `state_dict` keys as the reference, so these classes drop into its training / Grad-CAM scripts:

  QuadtreeCNN                 Quadtree_from scratch/models.py:214-305, resnet/models.py:70-180 (mode=, frozen)
  get_model                   Quadtree_from scratch/models.py:309-325 (and the resnet/ signature via get_model_resnet)

The ResNet-18 container is torchvision's own module tree (what the reference instantiates), so parameter names,
aliasing (`base_cnn.*` / `features_extractor.*` / `global_processor.*`) and hookability of `base_cnn.layer4`
are identical; only the arithmetic is replaced: every BasicBlock, the stem and the fusion head run as
`torch.autograd.Function`s that call the C ABI (bf16 channels-last activations, fp32 parameters/gradients).
There is no cuDNN / ATen fallback for the hot ops: CPU tensors or a missing libqtcnn.so raise.
"""
from __future__ import annotations

import os
import warnings
from typing import Optional

import torch
import torch.nn as nn
import torchvision
from torchvision.models.resnet import BasicBlock

from . import capi, ops
from . import functional as Fn
from .capi import check, ptr, stream
from .ops import BF16, L


def _require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise RuntimeError(f"{what}: the B200 path has no CPU implementation (got a {t.device} tensor); "
                           "move the model and inputs to a CUDA device")


def _zeros_like_param(p):
    return ops.grad_out(p)


# =================================================================================================
# Stem: conv1 7x7/s2 + bn1 + relu + maxpool 3x3/s2 (torchvision resnet.py:197-200)
# =================================================================================================
class _StemFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, conv_w, bn_w, bn_b, bn_mod, training):
        _require_cuda(x, "stem")
        n, c, h, w = x.shape
        cout = conv_w.shape[0]
        if conv_w.shape[1:] != (3, 7, 7) or c != 3:
            raise RuntimeError("stem: expected a 3-channel 7x7 stride-2 convolution (ResNet-18 conv1)")
        dev = x.device
        xp = torch.empty(n, h + 7, w + 8, 4, device=dev, dtype=BF16)
        xf, dtype, scale, shift = ops.stem_source(x)
        with ops.gemm_scope("stem_pack_input", 0.0, float(xf.element_size()) * xf.numel() + 2.0 * xp.numel()):
            check(L().qt_stem_pack_input_ex(ptr(xf), dtype, ptr(scale), ptr(shift), ptr(xp), n, 3, h, w, stream()), "stem_pack_input")
        w8 = ops.packed_stem(conv_w)
        ho, wo = h // 2, w // 2
        y = torch.empty(n, ho, wo, cout, device=dev, dtype=BF16)
        stats = None
        if training:
            stats = torch.empty(L().qt_stem_stat_rows(n, h, w), 2, cout, device=dev, dtype=torch.float32)
        with ops.gemm_scope("stem_fprop", 2.0 * n * ho * wo * 147 * cout):
            check(L().qt_stem_fprop(ptr(xp), ptr(w8), ptr(y), ptr(stats), n, h, w, cout, stream()), "stem_fprop")
        ops._count(2)
        st = ops.bn_finalize(stats, n * ho * wo, bn_mod, cout, dev, training)
        po, qo = capi.out_size(ho, 3, 2, 1), capi.out_size(wo, 3, 2, 1)
        out = torch.empty(n, po, qo, cout, device=dev, dtype=BF16)
        need_bwd = any(ctx.needs_input_grad)  # (grad mode is always off inside Function.forward)
        am = torch.empty(n, po, qo, cout, device=dev, dtype=torch.int8) if need_bwd else None
        yarg = torch.empty_like(out) if need_bwd else None  # conv output at each window's arg-max (backward statistics)
        # bn1 + relu + maxpool in one pass: the 112x112 activated map is never written
        with ops.gemm_scope("stem_tail_fwd", 0.0, 2.0 * y.numel() + (5.0 if need_bwd else 2.0) * out.numel()):  # y in; pooled out (+ yarg, int8 arg-max)
            check(L().qt_bn_relu_maxpool_fwd(ptr(y), ptr(st.scale), ptr(st.shift), ptr(out), ptr(am), ptr(yarg), n, ho, wo, cout, stream()),
                  "bn_relu_maxpool_fwd")
        ops._count()
        if need_bwd:
            ctx.saved = (xp, y, am, yarg, st, conv_w, bn_w, bn_b)
            ctx.dims = (n, h, w, cout, ho, wo)
            ctx.training = training
        return ops.as_nchw_view(out)

    @staticmethod
    def backward(ctx, dout):
        xp, y, am, yarg, st, conv_w, bn_w, bn_b = ctx.saved
        n, h, w, cout, ho, wo = ctx.dims
        dev = y.device
        dout = ops.as_nhwc(dout)
        dgamma = ops.grad_out(bn_w)
        dbeta = ops.grad_out(bn_b)
        dy = torch.empty_like(y)
        if ho % 2 == 0 and wo % 2 == 0:
            # fused: max-pool gather through the arg-max plane + ReLU mask recomputed from y + BatchNorm backward in two
            # passes over 2x2 pixel blocks; neither the activated 112x112 map nor its gradient is ever materialised
            wsb = L().qt_bn_workspace_bytes(cout)
            ws = ops.workspace(wsb, dev)
            # statistics from the pooled gradient + yarg; then one pass over (y, pooled gradient, arg-max plane) + one write of dy
            with ops.gemm_scope("stem_tail_bwd", 0.0, 4.0 * dout.numel() + (2.0 * y.numel() + 3.0 * dout.numel()) + 2.0 * dy.numel()):
                check(L().qt_bn_relu_maxpool_bwd(ptr(dout), ptr(am), ptr(y), ptr(yarg), ptr(st.scale), ptr(st.shift), ptr(st.mean), ptr(st.invstd),
                                                 ptr(bn_w.detach()), n, ho, wo, cout, ptr(dgamma), ptr(dbeta), 0 if ctx.training else 1,
                                                 ptr(dy), ptr(ws), wsb, stream()), "bn_relu_maxpool_bwd")
            ops._count(3)
        else:
            da = torch.empty_like(y)
            check(L().qt_maxpool2d_bwd(ptr(dout), ptr(am), ptr(da), n, ho, wo, cout, 3, 2, 1, stream()), "maxpool_bwd")
            ops._count()
            ops.bn_backward(da, None, y, st, bn_w.detach(), dgamma, dbeta, dy, None, eval_mode=not ctx.training, mask_from_y=True)
        dw = None
        if ctx.needs_input_grad[1]:
            dw = _zeros_like_param(conv_w)
            nbytes = L().qt_stem_wgrad_workspace_bytes(n, h, w, cout)
            ws = ops.workspace(nbytes, dev)
            with ops.gemm_scope("stem_wgrad", 2.0 * n * ho * wo * 147 * cout):
                check(L().qt_stem_wgrad(ptr(xp), ptr(dy), ptr(dw), 0, n, h, w, cout, 3, ptr(ws), ws.numel(), stream()), "stem_wgrad")
            ops._count(3)
        return None, dw, dgamma if ctx.needs_input_grad[2] else None, dbeta if ctx.needs_input_grad[3] else None, None, None


# =================================================================================================
# BasicBlock (torchvision resnet.py:59-105): conv3x3-BN-ReLU-conv3x3-BN (+identity | 1x1 conv-BN) -ReLU
# =================================================================================================
class _BlockFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w1, g1, b1, w2, g2, b2, wd, gd, bd, blk, training):
        """blk: (bn1 module, bn2 module, downsample-bn module | None, stride, downsample kernel, stride, padding). 4-D inputs
        are torchvision's 2-D BasicBlock (resnet.py:59-105), 5-D inputs the video BasicBlock of r3d_18 (video/resnet.py)."""
        _require_cuda(x, "BasicBlock")
        bn1_mod, bn2_mod, bnd_mod, stride, ds_k, ds_s, ds_p = blk
        is3d = x.dim() == 5
        xb = ops.as_channels_last(x)
        n, cin = xb.shape[0], xb.shape[-1]
        sp = tuple(xb.shape[1:-1]) if is3d else (1,) + tuple(xb.shape[1:-1])  # (D, H, W)
        cout = w1.shape[0]
        dev = xb.device
        kd = 3 if is3d else 1
        d1 = ops.conv_nd_desc(n, sp, cin, cout, (kd, 3, 3), (stride if is3d else 1, stride, stride), (1 if is3d else 0, 1, 1))
        so = ops.conv_out_hw(d1)
        d2 = ops.conv_nd_desc(n, so, cout, cout, (kd, 3, 3), (1, 1, 1), (1 if is3d else 0, 1, 1))
        m = n * so[0] * so[1] * so[2]
        oshape = (n,) + (so if is3d else so[1:]) + (cout,)
        y1 = torch.empty(oshape, device=dev, dtype=BF16)
        s1 = ops.conv_fprop(d1, xb, ops.packed_fprop(w1), y1, want_stats=training)
        st1 = ops.bn_finalize(s1, m, bn1_mod, cout, dev, training)
        a1 = torch.empty_like(y1)
        ops.bn_apply(y1, st1, a1, None, True)
        y2 = torch.empty_like(y1)
        s2 = ops.conv_fprop(d2, a1, ops.packed_fprop(w2), y2, want_stats=training)
        st2 = ops.bn_finalize(s2, m, bn2_mod, cout, dev, training)
        dd = yd = std = None
        if wd is not None:
            dd = ops.conv_nd_desc(n, sp, cin, cout, (ds_k if is3d else 1, ds_k, ds_k), (ds_s if is3d else 1, ds_s, ds_s),
                                  (ds_p if is3d else 0, ds_p, ds_p))
            yd = torch.empty_like(y1)
            sd = ops.conv_fprop(dd, xb, ops.packed_fprop(wd), yd, want_stats=training)
            std = ops.bn_finalize(sd, m, bnd_mod, cout, dev, training)
            idn = torch.empty_like(y1)
            ops.bn_apply(yd, std, idn, None, False)
        else:
            idn = xb
        out = torch.empty_like(y1)
        ops.bn_apply(y2, st2, out, idn, True)
        if any(ctx.needs_input_grad):
            ctx.saved = (xb, y1, a1, y2, yd, out, st1, st2, std, w1, w2, wd, g1, g2, gd, b1, b2, bd)
            ctx.descs = (d1, d2, dd)
            ctx.training = training
        return ops.as_channels_first_view(out)

    @staticmethod
    def backward(ctx, dout):
        xb, y1, a1, y2, yd, out, st1, st2, std, w1, w2, wd, g1, g2, gd, b1, b2, bd = ctx.saved
        d1, d2, dd = ctx.descs
        ev = not ctx.training
        dev = xb.device
        cout = y1.shape[-1]
        need = ctx.needs_input_grad
        dout = ops.as_channels_last(dout)

        def vec():
            return torch.empty(cout, device=dev)

        # bn2 + residual ReLU
        dg2, db2 = ops.grad_out(g2), ops.grad_out(b2)
        dy2 = torch.empty_like(y2)
        dz = torch.empty_like(y2)
        ops.bn_backward(dout, out, y2, st2, g2.detach(), dg2, db2, dy2, dz, eval_mode=ev)
        dw2 = None
        if need[4]:
            dw2 = _zeros_like_param(w2)
            ops.conv_wgrad(d2, a1, dy2, dw2)
        da1 = torch.empty_like(a1)
        ops.conv_dgrad(d2, dy2, ops.packed_dgrad(w2), da1)
        # bn1 + ReLU
        dg1, db1 = ops.grad_out(g1), ops.grad_out(b1)
        dy1 = torch.empty_like(y1)
        ops.bn_backward(da1, None, y1, st1, g1.detach(), dg1, db1, dy1, None, eval_mode=ev, mask_from_y=True)
        dw1 = None
        if need[1]:
            dw1 = _zeros_like_param(w1)
            ops.conv_wgrad(d1, xb, dy1, dw1)
        dwd = dgd = dbd = None
        dx = None
        if wd is not None:
            dgd, dbd = ops.grad_out(gd), ops.grad_out(bd)
            dyd = torch.empty_like(yd)
            ops.bn_backward(dz, None, yd, std, gd.detach(), dgd, dbd, dyd, None, eval_mode=ev)
            if need[7]:
                dwd = _zeros_like_param(wd)
                ops.conv_wgrad(dd, xb, dyd, dwd)
            if need[0]:
                # strided 1x1 downsample: its data gradient touches only the even positions, so the main-branch gradient
                # is written first and the downsample gradient accumulated onto it
                dx = torch.empty_like(xb)
                ops.conv_dgrad(d1, dy1, ops.packed_dgrad(w1), dx)
                ops.conv_dgrad(dd, dyd, ops.packed_dgrad(wd), dx, accumulate=True)
        elif need[0]:
            dx = dz  # identity branch gradient; the main-branch dgrad accumulates onto it in place
            ops.conv_dgrad(d1, dy1, ops.packed_dgrad(w1), dx, accumulate=True)
        return (ops.as_channels_first_view(dx) if dx is not None else None, dw1, dg1 if need[2] else None, db1 if need[3] else None,
                dw2, dg2 if need[5] else None, db2 if need[6] else None, dwd, dgd if need[8] else None,
                dbd if need[9] else None, None, None)


class FusedBasicBlock(BasicBlock):
    """torchvision BasicBlock whose forward runs on libqtcnn (same parameters / state_dict / hooks)."""

    def forward(self, x):
        ds = self.downsample
        geom = (self.bn1, self.bn2, ds[1] if ds is not None else None, self.conv1.stride[0],
                ds[0].kernel_size[0] if ds is not None else 1, ds[0].stride[0] if ds is not None else 1,
                ds[0].padding[0] if ds is not None else 0)
        return _BlockFn.apply(x, self.conv1.weight, self.bn1.weight, self.bn1.bias, self.conv2.weight, self.bn2.weight,
                              self.bn2.bias, ds[0].weight if ds is not None else None,
                              ds[1].weight if ds is not None else None, ds[1].bias if ds is not None else None, geom,
                              self.training)


class FusedFeatures(nn.Sequential):
    """`nn.Sequential(conv1, bn1, relu, maxpool, layer1, ...)` of the reference (QS/models.py:222-230) with the
    first four children executed as one fused stem. Children, indices and state_dict keys are unchanged."""

    def forward(self, x):
        mods = list(self.children())
        conv1, bn1 = mods[0], mods[1]
        out = _StemFn.apply(x, conv1.weight, bn1.weight, bn1.bias, bn1, bn1.training)
        for m in mods[4:]:
            out = m(out)
        return out


class _GlobalAvgPoolFn(torch.autograd.Function):
    """AdaptiveAvgPool2d((1,1)) on a channels-last bf16 map (used when global_processor is called directly)."""

    @staticmethod
    def forward(ctx, x):
        xb = ops.as_nhwc(x)
        n, h, w, c = xb.shape
        out = torch.empty(n, c, device=xb.device, dtype=BF16)
        check(L().qt_region_avgpool_fwd(ptr(xb), ptr(out), n, h * w, c, c, stream()), "avgpool")
        ops._count()
        ctx.shape = (n, h, w, c)
        return out.view(n, c, 1, 1)

    @staticmethod
    def backward(ctx, dout):
        n, h, w, c = ctx.shape
        d = dout.reshape(n, c).to(BF16).contiguous()
        dx = torch.empty(n, h, w, c, device=d.device, dtype=BF16)
        check(L().qt_region_avgpool_bwd(ptr(d), ptr(dx), ptr(dx), n, h * w, c, c, 0, stream()), "avgpool_bwd")
        ops._count()
        return ops.as_nchw_view(dx)


class FusedAvgPool(nn.AdaptiveAvgPool2d):
    def forward(self, x):
        return _GlobalAvgPoolFn.apply(x)


_quiet_pretrained = [os.environ.get("QTCNN_QUIET_PRETRAINED", "") == "1"]  # benches / tests use random init on purpose


def make_resnet18() -> nn.Module:
    """torchvision's ResNet-18 module tree (what the reference builds with `models.resnet18(...)`) with fused
    blocks. ImageNet weights are used only if the checkpoint is already in the local torch hub cache — the
    reference's `weights=IMAGENET1K_V1` needs a download that is impossible offline; otherwise torchvision's
    default initialisation (resnet.py:208-213) applies and real weights arrive through `load_state_dict`."""
    net = torchvision.models.resnet18(weights=None)
    ckpt = None
    try:  # only the cache probe is best effort; a checkpoint that exists but does not load is an error
        url = torchvision.models.ResNet18_Weights.IMAGENET1K_V1.url
        ckpt = os.path.join(torch.hub.get_dir(), "checkpoints", os.path.basename(url))
    except Exception:  # pragma: no cover
        ckpt = None
    if ckpt is not None and os.path.exists(ckpt):
        net.load_state_dict(torch.load(ckpt, map_location="cpu"))
    elif not _quiet_pretrained[0]:
        warnings.warn("ResNet-18 IMAGENET1K_V1 weights are not in the torch hub cache (no network): the backbone starts from "
                      "torchvision's random initialisation. The reference always starts from the pretrained checkpoint — load "
                      "real weights with load_state_dict before training, especially for the frozen-backbone variants.",
                      RuntimeWarning, stacklevel=3)
    for m in net.modules():
        if type(m) is BasicBlock:
            m.__class__ = FusedBasicBlock
    net.avgpool.__class__ = FusedAvgPool
    return net


# =================================================================================================
# Fusion head of QuadtreeCNN: quadrant convs + quadtree pooling + numerical MLP + classifier
# (QS/models.py:277-303; `mode` variants resnet/models.py:141-180)
# =================================================================================================
class _QuadHeadFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, base, l4, numerical, qw, qb, m0w, m0b, m3w, m3b, c0w, c0b, c3w, c3b, mode, p_drop, training, labels=None):
        use_img = mode in ("fusion", "image_only")
        use_num = mode in ("fusion", "numerical_only")
        dev = (base if use_img else numerical).device
        _require_cuda(base if use_img else numerical, "QuadtreeCNN head")
        p = float(p_drop) if training else 0.0
        seed1, seed2 = (ops.new_seed(), ops.new_seed()) if p > 0 else (0, 0)
        nimg = 5120 if use_img else 0
        nnum = m3w.shape[0] if use_num else 0
        ldf = nimg + nnum
        if ldf != c0w.shape[1]:
            raise RuntimeError(f"QuadtreeCNN head: classifier expects {c0w.shape[1]} features, built {ldf}")
        q = bb = l4b = h1 = None
        if use_img:
            bb = ops.as_nhwc(base)
            l4b = ops.as_nhwc(l4)
            n, h, w, cin = bb.shape
            cq = qw.shape[0]
            if (h, w) != (14, 14) or l4b.shape[1:] != (7, 7, 512) or cq != 128:
                raise RuntimeError("QuadtreeCNN expects 224x224 inputs (layer3 map 14x14), as the reference's "
                                   "image_feature_dim == 5120 assert does")
            dq = ops.quadrant_desc(n, h, w, cin, cq, 3, 1)
            q = torch.empty(4, n, h // 2, w // 2, cq, device=dev, dtype=BF16)
            ops.conv_fprop(dq, bb, ops.packed_fprop(qw), q, bias=qb.detach(), relu=True)
        else:
            n = numerical.shape[0]
            dq = None
        feat = torch.empty(n, ldf, device=dev, dtype=BF16)
        if use_img:
            # algorithmic bytes: quadrant maps + layer4 map in, 5120 features out (110 KB per image)
            with ops.gemm_scope("quadtree_pool_fwd", 0.0, 2.0 * (q.numel() + l4b.numel() + n * nimg)):
                check(L().qt_quadtree_pool_fwd(ptr(q), ptr(l4b), ptr(feat), n, 7, 7, 128, 49, 512, ldf, stream()), "quadtree_pool_fwd")
            ops._count()
        numf = None
        if use_num:
            numf = numerical.detach().float().contiguous()
            k0 = numf.shape[1]
            nh = m0w.shape[0]
            h1 = torch.empty(n, nh, device=dev)
            check(L().qt_small_linear_fwd(ptr(numf), 0, k0, ptr(m0w.detach()), ptr(m0b.detach()), n, nh, k0, 1, p, seed1, ptr(h1), nh,
                                          None, 0, stream()), "numerical_mlp.0")
            check(L().qt_small_linear_fwd(ptr(h1), 0, nh, ptr(m3w.detach()), ptr(m3b.detach()), n, nnum, nh, 0, 0.0, 0, None, 0,
                                          feat.data_ptr() + 2 * nimg, ldf, stream()), "numerical_mlp.3")
            ops._count(2)
        # classifier.0 on the tensor cores; everything behind it (ReLU + dropout, classifier.3, and — when the caller hands
        # over the labels — the cross-entropy loss) is ONE launch of the fused tail kernel
        nhid = c0w.shape[0]
        hbuf = torch.empty(n, nhid, device=dev)
        ws = ops.workspace(L().qt_linear_workspace_bytes(n, nhid, ldf), dev)
        wf0 = ops.packed_fprop(c0w)
        with ops.gemm_scope("linear_fprop", 2.0 * n * nhid * ldf):
            check(L().qt_linear_fprop(ptr(feat), ldf, ptr(wf0), ptr(c0b.detach()), ptr(hbuf), nhid,
                                      capi.QT_EPI_BIAS | capi.QT_EPI_OUT_F32, n, nhid, ldf, ptr(ws), ws.numel(), stream()), "classifier.0")
        nc = c3w.shape[0]
        logits = torch.empty(n, nc, device=dev)
        lab = lossbuf = None
        if labels is not None:
            lab = labels.detach().to(device=dev, dtype=torch.int64).contiguous()
            lossbuf = torch.empty(n + 1, device=dev)  # [per-row losses | mean]
        from .loss import _counter
        # algorithmic bytes: hidden row read + written back activated (fp32), classifier.3 weights once per CTA from L2
        with ops.gemm_scope("head_tail_fwd", 0.0, 8.0 * hbuf.numel() + 4.0 * c3w.numel()):
            check(L().qt_head_tail_fwd(ptr(hbuf), None, nhid, ptr(c3w.detach()), ptr(c3b.detach()), nc, ptr(lab), n, p, seed2, ptr(logits),
                                       ptr(lossbuf), lossbuf.data_ptr() + 4 * n if lossbuf is not None else None,
                                       ptr(_counter(dev)) if lab is not None else None, stream()), "head_tail_fwd")
        ops._count(3)
        if any(ctx.needs_input_grad):
            ctx.saved = (bb, q, feat, numf, h1, hbuf, logits, lab, qw, m0w, m3w, c0w, c3w, qb, m0b, m3b, c0b, c3b)
            ctx.cfg = (mode, p, seed1, seed2, n, ldf, nimg, nnum, dq)
        if labels is not None:
            ctx.mark_non_differentiable(logits)
            return lossbuf[n], logits
        return logits.view(logits.shape)  # not the saved object itself (no ctx <-> output reference cycle)

    @staticmethod
    def backward(ctx, dout, _dlogits_unused=None):
        bb, q, feat, numf, h1, hbuf, logits, lab, qw, m0w, m3w, c0w, c3w, qb, m0b, m3b, c0b, c3b = ctx.saved
        mode, p, seed1, seed2, n, ldf, nimg, nnum, dq = ctx.cfg
        need = ctx.needs_input_grad
        dev = feat.device
        nc, nhid = c3w.shape
        dh16 = torch.empty(n, nhid, device=dev, dtype=BF16)
        # fused tail backward: (with labels) softmax cross-entropy gradient scaled by the incoming loss gradient on the
        # device, classifier.3 dX through the ReLU / dropout gate -> bf16 operand of the classifier.0 gradient GEMMs
        if lab is not None:
            dl = torch.empty(n, nc, device=dev)
            up = dout.detach().reshape(1).float().contiguous()
            with ops.gemm_scope("head_tail_bwd", 0.0, 4.0 * hbuf.numel() + 2.0 * dh16.numel() + 4.0 * c3w.numel()):
                check(L().qt_head_tail_bwd(ptr(hbuf), nhid, ptr(c3w.detach()), nc, ptr(logits), ptr(lab), 1.0 / n, ptr(up), p, seed2, ptr(dl),
                                           ptr(dh16), n, stream()), "head_tail_bwd")
        else:
            dl = dout.detach().float().contiguous()
            with ops.gemm_scope("head_tail_bwd", 0.0, 4.0 * hbuf.numel() + 2.0 * dh16.numel() + 4.0 * c3w.numel()):
                check(L().qt_head_tail_bwd(ptr(hbuf), nhid, ptr(c3w.detach()), nc, None, None, 1.0, None, p, seed2, ptr(dl), ptr(dh16), n,
                                           stream()), "head_tail_bwd")
        # classifier.3 weight / bias gradient
        dc3w = ops.grad_out(c3w)
        dc3b = ops.grad_out(c3b)
        check(L().qt_small_linear_bwd_dw(ptr(dl), 0, nc, ptr(hbuf), 0, nhid, n, nc, nhid, ptr(dc3w), ptr(dc3b), 0, stream()), "classifier.3 dW")
        ops._count(2)
        # classifier.0
        ws = ops.workspace(L().qt_linear_workspace_bytes(n, nhid, ldf), dev)
        dc0w = ops.grad_out(c0w)
        with ops.gemm_scope("linear_wgrad", 2.0 * n * nhid * ldf):
            check(L().qt_linear_wgrad(ptr(feat), ldf, ptr(dh16), nhid, ptr(dc0w), 0, n, nhid, ldf, ptr(ws), ws.numel(), stream()), "classifier.0 dW")
        dc0b = ops.grad_out(c0b)
        ops.colsum(dh16, dc0b)
        dfeat = torch.empty(n, ldf, device=dev, dtype=BF16)
        wd0 = ops.packed_dgrad(c0w)
        with ops.gemm_scope("linear_dgrad", 2.0 * n * nhid * ldf):
            check(L().qt_linear_dgrad(ptr(dh16), nhid, ptr(wd0), ptr(dfeat), ldf, n, nhid, ldf, ptr(ws), ws.numel(), stream()),
                  "classifier.0 dX")
        ops._count(4)
        dm0w = dm0b = dm3w = dm3b = None
        if nnum:
            nh, k0 = m0w.shape
            dnum = dfeat.data_ptr() + 2 * nimg
            dm3w = ops.grad_out(m3w)
            dm3b = ops.grad_out(m3b)
            check(L().qt_small_linear_bwd_dw(dnum, 1, ldf, ptr(h1), 0, nh, n, nnum, nh, ptr(dm3w), ptr(dm3b), 0, stream()), "numerical_mlp.3 dW")
            dh1 = torch.empty(n, nh, device=dev)
            check(L().qt_small_linear_bwd_dx(dnum, 1, ldf, ptr(m3w.detach()), n, nnum, nh, ptr(h1), nh, p, seed1, ptr(dh1), nh, None, 0,
                                             stream()), "numerical_mlp.3 dX")
            dm0w = ops.grad_out(m0w)
            dm0b = ops.grad_out(m0b)
            check(L().qt_small_linear_bwd_dw(ptr(dh1), 0, nh, ptr(numf), 0, k0, n, nh, k0, ptr(dm0w), ptr(dm0b), 0, stream()), "numerical_mlp.0 dW")
            ops._count(3)
        dbase = dl4 = dqw = dqb = None
        if nimg:
            dqt = torch.empty_like(q)
            dl4b = torch.empty(n, 7, 7, 512, device=dev, dtype=BF16)
            # algorithmic bytes: feature-row gradient + quadrant maps in (arg-max recomputed), both gradients out
            with ops.gemm_scope("quadtree_pool_bwd", 0.0, 2.0 * (n * nimg + q.numel() + dqt.numel() + dl4b.numel())):
                check(L().qt_quadtree_pool_bwd(ptr(dfeat), ptr(q), ptr(dqt), ptr(dl4b), n, 7, 7, 128, 49, 512, ldf, stream()), "quadtree_pool_bwd")
            ops._count()
            dl4 = ops.as_nchw_view(dl4b) if need[1] else None
            if need[3]:
                dqw = _zeros_like_param(qw)
                ops.conv_wgrad(dq, bb, dqt, dqw)
            if need[4]:
                dqb = ops.grad_out(qb)
                ops.colsum(dqt.view(-1, qw.shape[0]), dqb)
            if need[0]:
                dbb = torch.empty_like(bb)
                ops.conv_dgrad(dq, dqt, ops.packed_dgrad(qw), dbb)
                dbase = ops.as_nchw_view(dbb)
        return (dbase, dl4, None, dqw, dqb, dm0w, dm0b, dm3w, dm3b, dc0w, dc0b, dc3w, dc3b, None, None, None, None)


# =================================================================================================
# QuadtreeCNN
# =================================================================================================
class QuadtreeCNN(nn.Module):
    """Drop-in for the reference QuadtreeCNN.

    Constructor / attributes / state_dict keys follow `Quadtree_from scratch/models.py:214-271`; `mode` and
    `freeze_backbone` cover the `resnet/models.py:70-133` variant (frozen backbone, fusion / image_only /
    numerical_only) including its Grad-CAM hook helpers (:131-139)."""

    def __init__(self, num_classes, cnn_feature_dim=512, numerical_feature_dim=47, dropout_rate=0.5, mode="fusion",
                 freeze_backbone=False):
        super().__init__()
        if mode not in ("fusion", "image_only", "numerical_only"):
            raise ValueError(f"Invalid mode: {mode}. Choose from 'fusion', 'image_only', 'numerical_only', "
                             "'standard_resnet_only'.")
        self.mode = mode
        self.base_cnn = make_resnet18()
        if freeze_backbone:
            for param in self.base_cnn.parameters():
                param.requires_grad = False
        b = self.base_cnn
        self.features_extractor = FusedFeatures(b.conv1, b.bn1, b.relu, b.maxpool, b.layer1, b.layer2, b.layer3)
        self.quadrant_processor = nn.Sequential(
            nn.Conv2d(256, cnn_feature_dim // 4, kernel_size=3, padding=1), nn.ReLU(inplace=True),
            nn.MaxPool2d(kernel_size=2, stride=2))
        self.global_processor = nn.Sequential(b.layer4, b.avgpool)
        self.image_feature_dim = 512 + ((cnn_feature_dim // 4) * 3 * 3 * 4)
        assert self.image_feature_dim == 5120, f"Image feature dim mismatch: Expected 5120, got {self.image_feature_dim}"
        self.numerical_mlp = nn.Sequential(
            nn.Linear(numerical_feature_dim, numerical_feature_dim * 2), nn.ReLU(inplace=True), nn.Dropout(dropout_rate),
            nn.Linear(numerical_feature_dim * 2, cnn_feature_dim // 2))
        self.numerical_output_dim = cnn_feature_dim // 2
        self.combined_feature_dim = self.image_feature_dim + self.numerical_output_dim
        if mode == "fusion":
            self.final_classifier_input_dim = self.combined_feature_dim
        elif mode == "image_only":
            self.final_classifier_input_dim = self.image_feature_dim
        else:
            self.final_classifier_input_dim = self.numerical_output_dim
        d = self.final_classifier_input_dim
        self.classifier = nn.Sequential(nn.Linear(d, d // 2), nn.ReLU(inplace=True), nn.Dropout(dropout_rate),
                                        nn.Linear(d // 2, num_classes))
        self.dropout_rate = dropout_rate
        self.gradients = None
        self.activations = None

    # Grad-CAM helpers of the resnet/ variant (resnet/models.py:131-139)
    def save_gradient_hook(self, module, grad_input, grad_output):
        self.gradients = grad_output[0]

    def save_activation_hook(self, module, input, output):
        self.activations = output

    def _run(self, image_input, numerical_input, labels):
        base = l4 = None
        if self.mode in ("fusion", "image_only"):
            # frozen-backbone variant (resnet/models.py:86-88): conv1..layer3 replay from a CUDA graph; trainable backbones,
            # hooked modules and profiling runs take the eager path inside the runner
            runner = self.__dict__.get("_features_graph")
            if runner is None:
                runner = self.__dict__["_features_graph"] = GraphedFrozenForward(self.features_extractor)
            base = runner(image_input)
            l4 = self.base_cnn.layer4(base)  # the module call keeps hooks on base_cnn.layer4 alive
        qp, mlp, cls = self.quadrant_processor[0], self.numerical_mlp, self.classifier
        return _QuadHeadFn.apply(base, l4, numerical_input, qp.weight, qp.bias, mlp[0].weight, mlp[0].bias, mlp[3].weight,
                                 mlp[3].bias, cls[0].weight, cls[0].bias, cls[3].weight, cls[3].bias, self.mode,
                                 self.dropout_rate, self.training, labels)

    def forward(self, image_input, numerical_input):
        """fp32 logits [B, num_classes] with a live autograd graph — the reference's signature; the scripts own the
        criterion. `image_input` may be fp32 / bf16 (already normalised) or uint8 (decoded pixels, normalised on the
        device with ImageNet mean / std like the reference's transform)."""
        return self._run(image_input, numerical_input, None)

    def training_loss(self, image_input, numerical_input, labels):
        """(loss, logits): `nn.CrossEntropyLoss()(self(image_input, numerical_input), labels)` with the loss computed
        inside the fused head-tail kernel (north star: fusion MLP + classifier + cross-entropy as one fused stage).
        `loss` carries the autograd graph, `logits` is detached (accuracy bookkeeping as in Quadtree_train.py:67-69)."""
        return self._run(image_input, numerical_input, labels)


# =================================================================================================
# Level-1 + level-2 models (Quadtree_from scratch/models.py:6-101, 105-210)
# =================================================================================================
class _HierBase(nn.Module):
    """Shared trunk of AttentionHierarchicalCNN / HierarchicalQuadtreeCNN: layer2 map (128x28x28) as the quadtree
    base, layer3->layer4->avgpool as the global branch, 3x3 convs + ReLU + AdaptiveAvgPool on the 4 quadrants
    (14x14) and the 16 sub-quadrants (7x7). As in the reference the ResNet is NOT kept as an attribute, so the
    state_dict has only features_extractor.* / global_processor.* keys for it."""

    def _build_trunk(self, numerical_feature_dim, dropout_rate):
        base_cnn = make_resnet18()
        self.features_extractor = FusedFeatures(base_cnn.conv1, base_cnn.bn1, base_cnn.relu, base_cnn.maxpool,
                                                base_cnn.layer1, base_cnn.layer2)
        self.global_processor = nn.Sequential(base_cnn.layer3, base_cnn.layer4, base_cnn.avgpool)
        self.quadrant_processor = nn.Sequential(nn.Conv2d(128, 128, kernel_size=3, padding=1), nn.ReLU(inplace=True),
                                                nn.AdaptiveAvgPool2d((1, 1)))
        self.sub_quadrant_processor = nn.Sequential(nn.Conv2d(128, 64, kernel_size=3, padding=1), nn.ReLU(inplace=True),
                                                    nn.AdaptiveAvgPool2d((1, 1)))
        self.dropout_rate = dropout_rate

    def _regions(self, image_input):
        base = self.features_extractor(image_input)                     # [B,128,28,28]
        glob = self.global_processor(base).flatten(1).float()           # [B,512]
        qp, sp = self.quadrant_processor[0], self.sub_quadrant_processor[0]
        quad = Fn.RegionConvPool.apply(base, qp.weight, qp.bias, 1)     # [4,B,128]
        sub = Fn.RegionConvPool.apply(base, sp.weight, sp.bias, 2)      # [16,B,64] quadrant-major
        bsz = glob.shape[0]
        quad = quad.permute(1, 0, 2).reshape(bsz, 4 * 128).float()
        sub = sub.permute(1, 0, 2).float()                              # [B,16,64]
        return glob, quad, sub

    def _head(self, image_features, numerical_input):
        mlp, cls = self.numerical_mlp, self.classifier
        num = Fn.SmallLinear.apply(numerical_input, mlp[0].weight, mlp[0].bias, True, self.dropout_rate, self.training)
        comb = torch.cat((image_features, num), dim=1)
        h = Fn.LinearTC.apply(comb, cls[0].weight, cls[0].bias, True, self.dropout_rate, self.training)
        return Fn.SmallLinear.apply(h, cls[3].weight, cls[3].bias, False, 0.0, self.training)


class AttentionHierarchicalCNN(_HierBase):
    """Drop-in for Quadtree_from scratch/models.py:6-101."""

    def __init__(self, num_classes, numerical_feature_dim=47, dropout_rate=0.5):
        super().__init__()
        self._build_trunk(numerical_feature_dim, dropout_rate)
        self.attention_gate = nn.Sequential(nn.Linear(64, 32), nn.ReLU(), nn.Linear(32, 1))
        total_image_feature_dim = 512 + (4 * 128) + 64
        self.numerical_mlp = nn.Sequential(nn.Linear(numerical_feature_dim, 128), nn.ReLU(inplace=True), nn.Dropout(dropout_rate))
        self.classifier = nn.Sequential(nn.Linear(total_image_feature_dim + 128, 1024), nn.ReLU(inplace=True),
                                        nn.Dropout(dropout_rate), nn.Linear(1024, num_classes))

    def forward(self, image_input, numerical_input):
        glob, quad, sub = self._regions(image_input)
        g = self.attention_gate
        hid = Fn.SmallLinear.apply(sub, g[0].weight, g[0].bias, True, 0.0, self.training)         # [B,16,32]
        scores = Fn.SmallLinear.apply(hid, g[2].weight, g[2].bias, False, 0.0, self.training)     # [B,16,1]
        attended = Fn.AttnPool.apply(sub, scores.squeeze(-1))                                      # [B,64]
        return self._head(torch.cat((glob, quad, attended), dim=1), numerical_input)


class HierarchicalQuadtreeCNN(_HierBase):
    """Quadtree_from scratch/models.py:105-210 with the evidently intended bottom-right slices (the reference's
    `w:` / `qw:` slices are empty and its forward raises — SURVEY.md §0.2); constructor and state_dict are
    the reference's."""

    def __init__(self, num_classes, numerical_feature_dim=47, dropout_rate=0.5):
        super().__init__()
        self._build_trunk(numerical_feature_dim, dropout_rate)
        total_image_feature_dim = 512 + (4 * 128) + (16 * 64)
        self.numerical_mlp = nn.Sequential(nn.Linear(numerical_feature_dim, 128), nn.ReLU(inplace=True), nn.Dropout(dropout_rate))
        self.classifier = nn.Sequential(nn.Linear(total_image_feature_dim + 128, 1024), nn.ReLU(inplace=True),
                                        nn.Dropout(dropout_rate), nn.Linear(1024, num_classes))

    def forward(self, image_input, numerical_input):
        glob, quad, sub = self._regions(image_input)
        return self._head(torch.cat((glob, quad, sub.reshape(sub.shape[0], -1)), dim=1), numerical_input)


# =================================================================================================
# StandardResNetCNN (resnet/models.py:7-65)
# =================================================================================================
class StandardResNetCNN(nn.Module):
    def __init__(self, num_classes, dropout_rate=0.5):
        super().__init__()
        self.base_cnn = make_resnet18()
        for param in self.base_cnn.parameters():
            param.requires_grad = False
        b = self.base_cnn
        self.features_extractor = FusedFeatures(b.conv1, b.bn1, b.relu, b.maxpool, b.layer1, b.layer2, b.layer3, b.layer4)
        self.avgpool = b.avgpool
        self.classifier = nn.Sequential(nn.Linear(512, 256), nn.ReLU(inplace=True), nn.Dropout(dropout_rate),
                                        nn.Linear(256, num_classes))
        self.dropout_rate = dropout_rate
        self.gradients = None
        self.activations = None

    def save_gradient_hook(self, module, grad_input, grad_output):
        self.gradients = grad_output[0]

    def save_activation_hook(self, module, input, output):
        self.activations = output

    def forward(self, image_input, numerical_input=None):
        feats = self.features_extractor(image_input)
        f = self.avgpool(feats).flatten(1).float()
        cls = self.classifier
        h = Fn.SmallLinear.apply(f, cls[0].weight, cls[0].bias, True, self.dropout_rate, self.training)
        return Fn.SmallLinear.apply(h, cls[3].weight, cls[3].bias, False, 0.0, self.training)


# =================================================================================================
# Quadtree3DCNN (3dcnn/models.py:96-214): Conv3d stack on the tensor cores, numeric LSTM on the persistent LSTM kernels
# =================================================================================================
class Quadtree3DCNN(nn.Module):
    _POOLS = {"conv3d_block1": (1, 2, 2), "conv3d_block2": (2, 2, 2), "conv3d_block3": (2, 2, 2),
              "conv3d_block4_new": (1, 2, 2), "conv3d_final_features": None}

    def __init__(self, num_classes, sequence_length=8, cnn_3d_feature_dim=1024, numerical_feature_dim=47, dropout_rate=0.6,
                 mode="quadtree_3d_fusion"):
        super().__init__()
        self.mode = mode
        self.sequence_length = sequence_length
        self.cnn_3d_feature_dim = cnn_3d_feature_dim
        self.numerical_feature_dim = numerical_feature_dim

        def block(cin, cout, pool):
            layers = [nn.Conv3d(cin, cout, kernel_size=(3, 3, 3), padding=(1, 1, 1)), nn.BatchNorm3d(cout), nn.ReLU(inplace=True)]
            if pool is not None:
                layers.append(nn.MaxPool3d(kernel_size=pool, stride=pool))
            return nn.Sequential(*layers)

        self.conv3d_block1 = block(3, 32, (1, 2, 2))
        self.conv3d_block2 = block(32, 64, (2, 2, 2))
        self.conv3d_block3 = block(64, 128, (2, 2, 2))
        self.conv3d_block4_new = block(128, 256, (1, 2, 2))
        self.conv3d_final_features = block(256, cnn_3d_feature_dim, None)
        self.global_avg_pool_3d = nn.AdaptiveAvgPool3d((1, 1, 1))
        self.numerical_lstm = nn.LSTM(input_size=numerical_feature_dim, hidden_size=numerical_feature_dim * 4, num_layers=2,
                                      batch_first=True, dropout=dropout_rate)
        self.numerical_lstm_output_dim = numerical_feature_dim * 4
        self.numerical_projection = nn.Sequential(nn.Linear(self.numerical_lstm_output_dim, cnn_3d_feature_dim // 2),
                                                  nn.ReLU(inplace=True), nn.Dropout(dropout_rate))
        self.numerical_final_dim = cnn_3d_feature_dim // 2
        if mode == "quadtree_3d_fusion":
            self.final_classifier_input_dim = cnn_3d_feature_dim + self.numerical_final_dim
        elif mode == "quadtree_3d_image_only":
            self.final_classifier_input_dim = cnn_3d_feature_dim
        else:
            raise ValueError(f"Invalid mode for Quadtree3DCNN: {mode}. Choose from 'quadtree_3d_fusion', 'quadtree_3d_image_only'.")
        d = self.final_classifier_input_dim
        self.classifier = nn.Sequential(nn.Linear(d, d // 2), nn.ReLU(inplace=True), nn.Dropout(dropout_rate),
                                        nn.Linear(d // 2, num_classes))
        self.dropout_rate = dropout_rate
        self.gradients = None
        self.activations = None

    def save_gradient_hook(self, module, grad_input, grad_output):
        self.gradients = grad_output[0]

    def save_activation_hook(self, module, input, output):
        self.activations = output

    def conv_stack(self, image_sequence_input):
        """[B,T,3,H,W] fp32 -> pooled conv features fp32 [B, cnn_3d_feature_dim]."""
        _require_cuda(image_sequence_input, "Quadtree3DCNN")
        x = Fn.PackClip.apply(image_sequence_input)
        for name, pool in self._POOLS.items():
            seq = getattr(self, name)
            conv, bn = seq[0], seq[1]
            x = Fn.Conv3dBnReluPool.apply(x, conv.weight, conv.bias, bn.weight, bn.bias, bn, pool, bn.training)
        return Fn.GlobalAvgPoolND.apply(x)

    def forward(self, image_sequence_input, numerical_sequence_input):
        image_features = self.conv_stack(image_sequence_input)
        if self.mode == "quadtree_3d_fusion":
            lstm_out = Fn.lstm_forward(self.numerical_lstm, numerical_sequence_input.to(image_features.device).float())
            proj = self.numerical_projection[0]
            num = Fn.SmallLinear.apply(lstm_out[:, -1, :], proj.weight, proj.bias, True, self.dropout_rate, self.training)
            combined = torch.cat((image_features, num), dim=1)
        else:
            combined = image_features
        cls = self.classifier
        h = Fn.LinearTC.apply(combined, cls[0].weight, cls[0].bias, True, self.dropout_rate, self.training)
        return Fn.SmallLinear.apply(h, cls[3].weight, cls[3].bias, False, 0.0, self.training)


# =================================================================================================
# r3d_18 backbone models (3dcnn/models.py:220-375): torchvision's VideoResNet module tree (same state_dict keys) with the
# stem and every 3-D BasicBlock executed on libqtcnn. As in the reference the backbone is frozen except layer4.
# =================================================================================================
class _VideoStemFn(torch.autograd.Function):
    """BasicStem of r3d_18: Conv3d(3,64,(3,7,7),s=(1,2,2),p=(1,3,3), no bias) + BatchNorm3d + ReLU. clips: [B,T,3,H,W] in the
    loader's layout (the permute of 3dcnn/models.py:258,345 only renames axes). Forward only — the reference freezes it."""

    @staticmethod
    def forward(ctx, clips, conv_w, bn_mod, training):
        _require_cuda(clips, "r3d_18 stem")
        if conv_w.requires_grad or (bn_mod.weight is not None and bn_mod.weight.requires_grad):
            raise RuntimeError("r3d_18 stem: the reference keeps the stem frozen; training it is not implemented")
        b, t, c, h, w = clips.shape
        cout = conv_w.shape[0]
        if tuple(conv_w.shape[1:]) != (3, 3, 7, 7) or c != 3:
            raise RuntimeError("r3d_18 stem: expected Conv3d(3, C, (3,7,7))")
        dev = clips.device
        xf, dtype, scale, shift = ops.stem_source(clips.reshape(b * t, c, h, w))
        xp = torch.empty(b * t, h + 7, w + 8, 4, device=dev, dtype=BF16)
        check(L().qt_stem_pack_input_ex(ptr(xf), dtype, ptr(scale), ptr(shift), ptr(xp), b * t, 3, h, w, stream()), "stem_pack_input")
        e = ops._entry(conv_w)
        if e.w8 is None:  # [cout][3 depth taps][8 row taps][32]
            w24 = torch.empty(cout, 3, 8, 32, device=dev, dtype=BF16)
            tmp = torch.empty(cout, 8, 32, device=dev, dtype=BF16)
            for kd in range(3):
                wk = conv_w.detach()[:, :, kd].contiguous()
                check(L().qt_wpack_stem(ptr(wk), ptr(tmp), cout, 3, 7, 7, stream()), "wpack_stem")
                w24[:, kd] = tmp
            e.w8 = w24
        ho, wo = h // 2, w // 2
        y = torch.empty(b, t, ho, wo, cout, device=dev, dtype=BF16)
        stats = torch.empty(L().qt_stem3d_stat_rows(b, t, h, w), 2, cout, device=dev) if training else None
        with ops.gemm_scope("stem3d_fprop", 2.0 * b * t * ho * wo * 441 * cout):
            check(L().qt_stem3d_fprop(ptr(xp), ptr(e.w8), ptr(y), ptr(stats), b, t, h, w, cout, stream()), "stem3d_fprop")
        ops._count(2)
        st = ops.bn_finalize(stats, b * t * ho * wo, bn_mod, cout, dev, training)
        out = torch.empty_like(y)
        ops.bn_apply(y, st, out, None, True)
        return ops.as_channels_first_view(out)

    @staticmethod
    def backward(ctx, g):
        return None, None, None, None


class FusedVideoBlock(nn.Module):
    """torchvision.models.video.resnet.BasicBlock (conv1 = Sequential(conv, bn, relu), conv2 = Sequential(conv, bn)) on
    libqtcnn; instances are retargeted in place, so parameters / state_dict keys are torchvision's."""

    def forward(self, x):
        ds = self.downsample
        c1, c2 = self.conv1, self.conv2
        geom = (c1[1], c2[1], ds[1] if ds is not None else None, c1[0].stride[0],
                ds[0].kernel_size[0] if ds is not None else 1, ds[0].stride[0] if ds is not None else 1,
                ds[0].padding[0] if ds is not None else 0)
        return _BlockFn.apply(x, c1[0].weight, c1[1].weight, c1[1].bias, c2[0].weight, c2[1].weight, c2[1].bias,
                              ds[0].weight if ds is not None else None, ds[1].weight if ds is not None else None,
                              ds[1].bias if ds is not None else None, geom, self.training)


class FusedVideoStem(nn.Sequential):
    def forward(self, clips_btchw):
        return _VideoStemFn.apply(clips_btchw, self[0].weight, self[1], self[1].training)


def make_r3d18() -> nn.Module:
    """torchvision's r3d_18 module tree (what the reference builds with `video_models.r3d_18(weights=KINETICS400_V1)`) with
    fused stem / blocks. Kinetics weights only if the checkpoint is already in the torch hub cache (no network here)."""
    from torchvision.models import video as video_models
    from torchvision.models.video.resnet import BasicBlock as VideoBasicBlock
    net = video_models.r3d_18(weights=None)
    ckpt = None
    try:
        url = video_models.R3D_18_Weights.KINETICS400_V1.url
        ckpt = os.path.join(torch.hub.get_dir(), "checkpoints", os.path.basename(url))
    except Exception:  # pragma: no cover
        ckpt = None
    if ckpt is not None and os.path.exists(ckpt):
        net.load_state_dict(torch.load(ckpt, map_location="cpu"))
    elif not _quiet_pretrained[0]:
        warnings.warn("r3d_18 KINETICS400_V1 weights are not in the torch hub cache (no network): the frozen backbone starts from "
                      "random initialisation; load real weights with load_state_dict.", RuntimeWarning, stacklevel=3)
    for mod in net.modules():
        if type(mod) is VideoBasicBlock:
            mod.__class__ = FusedVideoBlock
    net.stem.__class__ = FusedVideoStem
    return net


class _VideoFeatures(nn.Sequential):
    """`nn.Sequential(stem, layer1..layer4)` of HybridQuadtree3DCNN (3dcnn/models.py:276-282): fed with the loader's
    [B,T,3,H,W] clips directly (the stem reads that layout; the reference's permute is a view)."""


class ResNet3DVideo(nn.Module):
    """Drop-in for 3dcnn/models.py:220-262 (r3d_18 fine-tuning: everything frozen except layer4 and the new fc head)."""

    def __init__(self, num_classes, dropout_rate=0.5):
        super().__init__()
        self.r3d_model = make_r3d18()
        for param in self.r3d_model.parameters():
            param.requires_grad = False
        for param in self.r3d_model.layer4.parameters():
            param.requires_grad = True
        num_ftrs = self.r3d_model.fc.in_features
        self.r3d_model.fc = nn.Sequential(nn.Linear(num_ftrs, num_ftrs // 2), nn.ReLU(inplace=True), nn.Dropout(dropout_rate),
                                          nn.Linear(num_ftrs // 2, num_classes))
        self.dropout_rate = dropout_rate

    def forward(self, image_sequence_input, numerical_input=None):
        r = self.r3d_model
        x = r.stem(image_sequence_input)
        for layer in (r.layer1, r.layer2, r.layer3, r.layer4):
            x = layer(x)
        f = Fn.GlobalAvgPoolND.apply(ops.as_channels_last(x))
        fc = r.fc
        h = Fn.SmallLinear.apply(f, fc[0].weight, fc[0].bias, True, self.dropout_rate, self.training)
        return Fn.SmallLinear.apply(h, fc[3].weight, fc[3].bias, False, 0.0, self.training)


class HybridQuadtree3DCNN(nn.Module):
    """Drop-in for 3dcnn/models.py:266-375: frozen r3d_18 extractor (layer4 trainable) + numeric LSTM + fusion classifier."""

    def __init__(self, num_classes, sequence_length=8, numerical_feature_dim=47, dropout_rate=0.6, mode="hybrid_quadtree_3d_fusion"):
        super().__init__()
        self.mode = mode
        self.sequence_length = sequence_length
        self.numerical_feature_dim = numerical_feature_dim
        r3d_base = make_r3d18()
        self.pretrained_image_extractor = _VideoFeatures(r3d_base.stem, r3d_base.layer1, r3d_base.layer2, r3d_base.layer3,
                                                         r3d_base.layer4)
        for param in self.pretrained_image_extractor.parameters():
            param.requires_grad = False
        for param in self.pretrained_image_extractor[4].parameters():
            param.requires_grad = True
        self.cnn_3d_feature_dim = 512
        self.global_avg_pool_3d = nn.AdaptiveAvgPool3d((1, 1, 1))
        self.numerical_lstm = nn.LSTM(input_size=numerical_feature_dim, hidden_size=numerical_feature_dim * 4, num_layers=2,
                                      batch_first=True, dropout=dropout_rate)
        self.numerical_lstm_output_dim = numerical_feature_dim * 4
        self.numerical_projection = nn.Sequential(nn.Linear(self.numerical_lstm_output_dim, self.cnn_3d_feature_dim // 2),
                                                  nn.ReLU(inplace=True), nn.Dropout(dropout_rate))
        self.numerical_final_dim = self.cnn_3d_feature_dim // 2
        if mode == "hybrid_quadtree_3d_fusion":
            self.final_classifier_input_dim = self.cnn_3d_feature_dim + self.numerical_final_dim
        elif mode == "hybrid_quadtree_3d_image_only":
            self.final_classifier_input_dim = self.cnn_3d_feature_dim
        else:
            raise ValueError(f"Invalid mode for HybridQuadtree3DCNN: {mode}. Choose from 'hybrid_quadtree_3d_fusion', "
                             "'hybrid_quadtree_3d_image_only'.")
        d = self.final_classifier_input_dim
        self.classifier = nn.Sequential(nn.Linear(d, d // 2), nn.ReLU(inplace=True), nn.Dropout(dropout_rate), nn.Linear(d // 2, num_classes))
        self.dropout_rate = dropout_rate
        self.gradients = None
        self.activations = None

    def save_gradient_hook(self, module, grad_input, grad_output):
        self.gradients = grad_output[0]

    def save_activation_hook(self, module, input, output):
        self.activations = output

    def forward(self, image_sequence_input, numerical_sequence_input):
        x = self.pretrained_image_extractor(image_sequence_input)
        image_features = Fn.GlobalAvgPoolND.apply(ops.as_channels_last(x))
        if self.mode == "hybrid_quadtree_3d_fusion":
            lstm_out = Fn.lstm_forward(self.numerical_lstm, numerical_sequence_input.to(image_features.device).float())
            proj = self.numerical_projection[0]
            num = Fn.SmallLinear.apply(lstm_out[:, -1, :], proj.weight, proj.bias, True, self.dropout_rate, self.training)
            combined = torch.cat((image_features, num), dim=1)
        else:
            combined = image_features
        cls = self.classifier
        h = Fn.SmallLinear.apply(combined, cls[0].weight, cls[0].bias, True, self.dropout_rate, self.training)
        return Fn.SmallLinear.apply(h, cls[3].weight, cls[3].bias, False, 0.0, self.training)


# =================================================================================================
# CnnLstm (cnn+lstm/models.py:14-89): every frame through the frozen ResNet-18 on the tensor cores; the temporal
# LSTM runs on the persistent LSTM kernels (functional.LSTM)
# =================================================================================================
class GraphedFrozenForward:
    """Forward of a frozen, gradient-free sub-network replayed from a CUDA graph.

    The frozen ResNet-18 of CnnLstm (cnn+lstm/models.py:21-27) is ~80 kernel launches with static shapes, static (frozen)
    weights and no autograd state; issued one by one from Python they take longer on the host than on the GPU once several
    ranks share a box (the 8-GPU line of round 2 ran at 0.67 of linear for that reason). The launches are recorded once
    per (input shape, dtype, train/eval mode, parameter versions) with `torch.cuda.graph` — libqtcnn launches on torch's
    current stream, so they are captured like any other kernel — and replayed with one driver call per step. The input
    is copied into the graph's static buffer; train-mode BatchNorm running statistics are updated by the replayed kernels
    exactly once per step (the warm-up runs needed for capture are rolled back). Falls back to eager execution when
    any parameter of the sub-network is trainable, the input requires a gradient, a module of the sub-network carries
    hooks (Grad-CAM), profiling scopes are active, or `QTCNN_NO_GRAPH=1`."""

    _MAX_GRAPHS = 4  # e.g. train / eval x two batch sizes; each graph owns the activations of one forward

    def __init__(self, module):
        self.module = module
        self.cache = {}  # key -> [graph, static_in, static_out, launches]

    def __deepcopy__(self, memo):  # graphs are neither copyable nor picklable: the copy re-captures on first use
        return None

    def __reduce__(self):
        return (_no_runner, ())

    def usable(self, x) -> bool:
        if os.environ.get("QTCNN_NO_GRAPH", "") == "1" or ops.profiling() or not x.is_cuda:
            return False
        # only fully frozen sub-networks: a trainable weight's bf16 packs are refreshed lazily by eager forwards after an
        # optimizer step, which a replay would bypass (frozen weights change only through `_version`-bumping torch ops or
        # ops.invalidate_packed_weights(), both part of the capture key)
        if any(p.requires_grad for p in self.module.parameters()) or (torch.is_grad_enabled() and x.requires_grad):
            return False
        if torch.cuda.is_current_stream_capturing():
            return False
        for m in self.module.modules():
            if m._forward_hooks or m._forward_pre_hooks or m._backward_hooks or getattr(m, "_backward_pre_hooks", None):
                return False
        return True

    def _make_key(self, x):
        return (tuple(x.shape), x.dtype, x.device.index, self.module.training, ops._hard_epoch,
                tuple((p.data_ptr(), p._version) for p in self.module.parameters()),
                tuple(b.data_ptr() for b in self.module.buffers()))

    def _capture(self, x):
        bufs = [b for b in self.module.buffers()]
        saved = [b.detach().clone() for b in bufs]
        static_in = torch.empty_like(x)
        static_in.copy_(x)
        side = torch.cuda.Stream(device=x.device)
        side.wait_stream(torch.cuda.current_stream(x.device))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(2):  # weight packs, workspaces, kernel attributes: everything lazy happens outside the capture
                self.module(static_in)
        torch.cuda.current_stream(x.device).wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        n0 = ops.launches()
        # thread_local: a DataLoader pin-memory thread may call the CUDA allocator while this thread captures
        with torch.cuda.graph(graph, capture_error_mode="thread_local"), torch.no_grad():
            static_out = self.module(static_in)
        launches = ops.launches() - n0
        ops._count(-launches)  # recorded, not executed
        with torch.no_grad():
            for b, v in zip(bufs, saved):  # roll the warm-up's running-statistics updates back
                b.copy_(v)
        return [graph, static_in, static_out, launches]

    def __call__(self, x):
        if not self.usable(x):
            return self.module(x)
        key = self._make_key(x)
        ent = self.cache.get(key)
        if ent is None:
            while len(self.cache) >= self._MAX_GRAPHS:
                self.cache.pop(next(iter(self.cache)))
            ent = self.cache[key] = self._capture(x)
        graph, static_in, static_out, launches = ent
        static_in.copy_(x)
        graph.replay()
        ops._count(launches)
        return static_out.clone()  # the graph's output buffer is overwritten by the next replay


def _no_runner():
    return None


class CnnLstm(nn.Module):
    def __init__(self, num_classes, sequence_length=4, numerical_feature_dim=47, dropout_rate=0.5, lstm_hidden_size=256):
        super().__init__()
        self.sequence_length = sequence_length
        resnet = make_resnet18()
        self.cnn_backbone = FusedFeatures(*list(resnet.children())[:-1])  # same children / indices / keys as the reference
        for param in self.cnn_backbone.parameters():
            param.requires_grad = False
        self.numerical_mlp = nn.Sequential(nn.Linear(numerical_feature_dim, 128), nn.ReLU(), nn.Linear(128, 128))
        self.lstm = nn.LSTM(input_size=512 + 128, hidden_size=lstm_hidden_size, num_layers=2, batch_first=True,
                            dropout=dropout_rate)
        self.classifier = nn.Sequential(nn.Linear(lstm_hidden_size, 128), nn.ReLU(), nn.Dropout(dropout_rate),
                                        nn.Linear(128, num_classes))
        self.dropout_rate = dropout_rate

    def forward(self, image_sequence, numerical_sequence):
        _require_cuda(image_sequence, "CnnLstm")
        batch_size, seq_len, c, h, w = image_sequence.shape
        c_in = image_sequence.reshape(batch_size * seq_len, c, h, w)
        runner = self.__dict__.get("_backbone_graph")
        if runner is None:
            runner = self.__dict__["_backbone_graph"] = GraphedFrozenForward(self.cnn_backbone)  # not a submodule / state_dict entry
        c_out = runner(c_in).flatten(1).float().view(batch_size, seq_len, -1)   # (B, T, 512)
        mlp = self.numerical_mlp
        num = numerical_sequence.to(c_out.device).float()
        n_out = Fn.SmallLinear.apply(num, mlp[0].weight, mlp[0].bias, True, 0.0, self.training)
        n_out = Fn.SmallLinear.apply(n_out, mlp[2].weight, mlp[2].bias, False, 0.0, self.training)
        lstm_out = Fn.lstm_forward(self.lstm, torch.cat((c_out, n_out), dim=2))
        final_state = lstm_out[:, -1, :]
        cls = self.classifier
        hdn = Fn.SmallLinear.apply(final_state, cls[0].weight, cls[0].bias, True, self.dropout_rate, self.training)
        return Fn.SmallLinear.apply(hdn, cls[3].weight, cls[3].bias, False, 0.0, self.training)


def get_model_seq(model_name, num_classes, device, seq_len=4, num_features=47):
    """`get_model` of cnn+lstm/models.py:147-155 (same argument names). Ji3DCNN is outside SURVEY.md §8."""
    if model_name == "cnn_lstm":
        model = CnnLstm(num_classes, sequence_length=seq_len, numerical_feature_dim=num_features)
    elif model_name == "3d_cnn":
        raise ValueError("get_model: '3d_cnn' (Ji3DCNN) is not on the accelerated path (SURVEY.md §8 scope)")
    else:
        raise ValueError(f"Unknown model name: {model_name}")
    return model.to(device)


def get_model(model_name="quadtree", num_classes=8, device="cuda", print_num_params=True):
    """`get_model` of Quadtree_from scratch/models.py:309-325 (same argument names and printout)."""
    name = model_name.lower()
    if name == "quadtree":
        model = QuadtreeCNN(num_classes=num_classes).to(device)
    elif name == "hierarchical_quadtree":
        model = HierarchicalQuadtreeCNN(num_classes=num_classes).to(device)
    elif name == "attention_hierarchical":
        model = AttentionHierarchicalCNN(num_classes=num_classes).to(device)
    else:
        # the reference's StandardMultimodalCNN is a `pass` stub there and raises TypeError
        raise ValueError(f"get_model: backbone '{model_name}' is outside the B200 hot path (SURVEY.md §8); supported: "
                         "'quadtree', 'hierarchical_quadtree', 'attention_hierarchical'")
    if print_num_params:
        num_params = sum(p.numel() for p in model.parameters() if p.requires_grad)
        print(f"Model: '{name.upper()}' | Trainable Parameters: {num_params / 1e6:.2f} Million")
    return model


def get_model_resnet(num_classes, device, numerical_feature_dim=47, mode="fusion", print_num_params=True):
    """`get_model` of resnet/models.py:183-194 (frozen backbone + mode switch)."""
    if mode == "standard_resnet_only":
        model = StandardResNetCNN(num_classes=num_classes).to(device)
    else:
        model = QuadtreeCNN(num_classes=num_classes, numerical_feature_dim=numerical_feature_dim, mode=mode,
                            freeze_backbone=True).to(device)
    if print_num_params:
        num_params = sum(p.numel() for p in model.parameters() if p.requires_grad)
        print(f"Number of trainable parameters: {num_params / 1e6:.2f} Million (Mode: {mode})")
    return model


def get_model_3d(num_classes, device, numerical_feature_dim=47, mode="fusion", sequence_length=8, print_num_params=True):
    """`get_model` of 3dcnn/models.py:493-522 for the modes on the accelerated path."""
    if mode == "standard_resnet_only":
        model = StandardResNetCNN(num_classes=num_classes).to(device)
    elif mode in ("quadtree_3d_fusion", "quadtree_3d_image_only"):
        model = Quadtree3DCNN(num_classes=num_classes, sequence_length=sequence_length,
                              numerical_feature_dim=numerical_feature_dim, mode=mode, cnn_3d_feature_dim=1024).to(device)
    elif mode == "resnet_3d_video_only":
        model = ResNet3DVideo(num_classes=num_classes).to(device)
    elif mode in ("hybrid_quadtree_3d_fusion", "hybrid_quadtree_3d_image_only"):
        model = HybridQuadtree3DCNN(num_classes=num_classes, sequence_length=sequence_length,
                                    numerical_feature_dim=numerical_feature_dim, mode=mode).to(device)
    else:
        model = QuadtreeCNN(num_classes=num_classes, numerical_feature_dim=numerical_feature_dim, mode=mode,
                            freeze_backbone=True).to(device)
    if print_num_params:
        num_params = sum(p.numel() for p in model.parameters() if p.requires_grad)
        print(f"Number of trainable parameters: {num_params / 1e6:.2f} Million (Mode: {mode})")
    return model
