"""Composable autograd Functions over the C ABI, used by the level-2 (hierarchical / attention), plain-ResNet and
3-D model mirrors. Each Function is one reference operator group with its exact backward; tensors crossing
Function boundaries are either channels-last bf16 maps or small fp32 / bf16 feature matrices."""
from __future__ import annotations

import os

import torch

from . import capi, ops
from .capi import check, ptr, stream
from .ops import BF16, L


# ------------------------------------------------------------------------------------------------------------
# Region convolution + ReLU + AdaptiveAvgPool2d((1,1)) over the quadtree regions of a feature map
# (Quadtree_from scratch/models.py:21-30, 60-79): level 1 = the 4 quadrants, level 2 = the 16 sub-quadrants
# (quadrant-major, each TL,TR,BL,BR). Zero padding is per region, as in the reference (views are convolved
# separately there).
# ------------------------------------------------------------------------------------------------------------
def _region_descs(n, H, W, cin, cout, level):
    """One grouped (4-view) conv descriptor per parent region; outputs are region-major [G][n][qh][qw][cout]."""
    key = ("regions", n, H, W, cin, cout, level)
    cached = ops._desc_cache.get(key)
    if cached is not None:
        return cached
    if H % (2 * level) or W % (2 * level):
        raise RuntimeError("region conv: feature-map size must divide evenly into the quadtree level")
    parents = [(0, 0, H, W)] if level == 1 else [(0, 0, H // 2, W // 2), (0, W // 2, H // 2, W // 2),
                                                 (H // 2, 0, H // 2, W // 2), (H // 2, W // 2, H // 2, W // 2)]
    descs = []
    for ri, (h0, w0, hs, ws) in enumerate(parents):
        qh, qw = hs // 2, ws // 2
        xs = (H * W * cin, 0, W * cin, cin)
        base = (h0 * W + w0) * cin
        xoff = (base, base + qw * cin, base + qh * W * cin, base + qh * W * cin + qw * cin)
        yoff = tuple((ri * 4 + q) * n * qh * qw * cout for q in range(4))
        descs.append(capi.conv_desc(n, (1, qh, qw), cin, cout, (1, 3, 3), (1, 1, 1), (0, 1, 1), x_stride=xs, groups=4,
                                    x_group_off=xoff, y_group_off=yoff))
    out = (descs, 4 * len(parents), parents[0][2] // 2, parents[0][3] // 2)
    ops._desc_cache[key] = out
    return out


class RegionConvPool(torch.autograd.Function):
    """pooled[g, b, :] = mean_{pixels}( relu(conv3x3(region_g(base)) + bias) ) -> bf16 [G, B, cout]."""

    @staticmethod
    def forward(ctx, base, w, b, level):
        xb = ops.as_nhwc(base)
        n, H, W, cin = xb.shape
        cout = w.shape[0]
        descs, G, qh, qw = _region_descs(n, H, W, cin, cout, level)
        y = torch.empty(G, n, qh, qw, cout, device=xb.device, dtype=BF16)
        wf = ops.packed_fprop(w)
        for d in descs:
            ops.conv_fprop(d, xb, wf, y, bias=b.detach(), relu=True)
        pooled = torch.empty(G * n, cout, device=xb.device, dtype=BF16)
        check(L().qt_region_avgpool_fwd(ptr(y), ptr(pooled), G * n, qh * qw, cout, cout, stream()), "region_avgpool_fwd")
        ops._count()
        if any(ctx.needs_input_grad):
            ctx.saved = (xb, y, w, descs)
            ctx.dims = (G, n, qh, qw, cout)
        return pooled.view(G, n, cout)

    @staticmethod
    def backward(ctx, dpooled):
        xb, y, w, descs = ctx.saved
        G, n, qh, qw, cout = ctx.dims
        dp = dpooled.to(BF16).contiguous()
        dy = torch.empty_like(y)
        check(L().qt_region_avgpool_bwd(ptr(dp), ptr(y), ptr(dy), G * n, qh * qw, cout, cout, 1, stream()), "region_avgpool_bwd")
        ops._count()
        dw = db = dbase = None
        if ctx.needs_input_grad[2]:
            db = torch.empty(cout, device=xb.device)
            ops.colsum(dy.view(-1, cout), db)
        if ctx.needs_input_grad[1]:
            dw = ops.grad_out(w)
            for i, d in enumerate(descs):
                ops.conv_wgrad(d, xb, dy, dw, accumulate=i > 0)
        if ctx.needs_input_grad[0]:
            dxb = torch.empty_like(xb)
            wd = ops.packed_dgrad(w)
            for d in descs:
                ops.conv_dgrad(d, dy, wd, dxb)
            dbase = ops.as_nchw_view(dxb)
        return dbase, dw, db, None


# ------------------------------------------------------------------------------------------------------------
# Linear layers
# ------------------------------------------------------------------------------------------------------------
class SmallLinear(torch.autograd.Function):
    """fp32 nn.Linear (+ReLU, +Dropout) for the narrow layers: out = drop(relu(x W^T + b)), fp32 [M, N]."""

    @staticmethod
    def forward(ctx, x, w, b, relu, p_drop, training):
        lead = x.shape[:-1]
        k = x.shape[-1]
        x2 = x.detach().reshape(-1, k)
        is16 = x2.dtype == BF16
        if not is16:
            x2 = x2.float()
        x2 = x2.contiguous()
        m, n = x2.shape[0], w.shape[0]
        p = float(p_drop) if training else 0.0
        seed = ops.new_seed() if p > 0 else 0
        out = torch.empty(m, n, device=x2.device)
        check(L().qt_small_linear_fwd(ptr(x2), 1 if is16 else 0, k, ptr(w.detach()), ptr(b.detach()) if b is not None else None, m, n,
                                      k, 1 if relu else 0, p, seed, ptr(out), n, None, 0, stream()), "small_linear_fwd")
        ops._count()
        if any(ctx.needs_input_grad):
            ctx.saved = (x2, w, out)
            ctx.cfg = (is16, relu, p, seed, lead, k, b is not None)
        return out.view(*lead, n)

    @staticmethod
    def backward(ctx, dout):
        x2, w, out = ctx.saved
        is16, relu, p, seed, lead, k, has_b = ctx.cfg
        m, n = out.shape
        dz = dout.detach().reshape(m, n).float().contiguous()
        if relu or p > 0:
            dz2 = torch.empty_like(dz)
            check(L().qt_relu_dropout_bwd(ptr(dz), ptr(out), ptr(dz2), None, m * n, p, seed, 1 if relu else 0, stream()), "relu_dropout_bwd")
            ops._count()
            dz = dz2
        dw = db = dx = None
        if ctx.needs_input_grad[1] or (has_b and ctx.needs_input_grad[2]):
            dw = ops.grad_out(w)
            db = torch.empty(n, device=dz.device) if has_b else None
            check(L().qt_small_linear_bwd_dw(ptr(dz), 0, n, ptr(x2), 1 if is16 else 0, k, m, n, k, ptr(dw), ptr(db), 0, stream()),
                  "small_linear_bwd_dw")
            ops._count()
        if ctx.needs_input_grad[0]:
            dx = torch.empty(m, k, device=dz.device)
            check(L().qt_small_linear_bwd_dx(ptr(dz), 0, n, ptr(w.detach()), m, n, k, None, 0, 0.0, 0, ptr(dx), k, None, 0, stream()),
                  "small_linear_bwd_dx")
            ops._count()
            dx = dx.view(*lead, k)
        return dx, dw, db, None, None, None


class LinearTC(torch.autograd.Function):
    """nn.Linear (+ReLU, +Dropout) on the tensor cores: x fp32/bf16 [B, K] (K % 8 == 0) -> fp32 [B, N]."""

    @staticmethod
    def forward(ctx, x, w, b, relu, p_drop, training):
        x16 = x.detach().to(BF16).contiguous()
        bsz, k = x16.shape
        n = w.shape[0]
        p = float(p_drop) if training else 0.0
        seed = ops.new_seed() if p > 0 else 0
        out = torch.empty(bsz, n, device=x16.device)
        ws = ops.workspace(L().qt_linear_workspace_bytes(bsz, n, k), x16.device)
        wf = ops.packed_fprop(w)
        with ops.gemm_scope("linear_fprop", 2.0 * bsz * n * k):
            check(L().qt_linear_fprop(ptr(x16), k, ptr(wf), ptr(b.detach()), ptr(out), n, capi.QT_EPI_BIAS | capi.QT_EPI_OUT_F32, bsz, n,
                                      k, ptr(ws), ws.numel(), stream()), "linear_fprop")
        if relu or p > 0:
            check(L().qt_relu_dropout(ptr(out), None, bsz * n, p, seed, 1 if relu else 0, stream()), "relu_dropout")
        ops._count(2)
        if any(ctx.needs_input_grad):
            ctx.saved = (x16, w, out)
            ctx.cfg = (relu, p, seed)
        return out.view(out.shape)  # not the saved object itself (no ctx <-> output reference cycle)

    @staticmethod
    def backward(ctx, dout):
        x16, w, out = ctx.saved
        relu, p, seed = ctx.cfg
        bsz, k = x16.shape
        n = w.shape[0]
        d = dout.detach().float().contiguous()
        dz16 = torch.empty(bsz, n, device=d.device, dtype=BF16)
        check(L().qt_relu_dropout_bwd(ptr(d), ptr(out), None, ptr(dz16), bsz * n, p, seed, 1 if relu else 0, stream()), "relu_dropout_bwd")
        ws = ops.workspace(L().qt_linear_workspace_bytes(bsz, n, k), d.device)
        dw = db = dx = None
        if ctx.needs_input_grad[1]:
            dw = ops.grad_out(w)
            with ops.gemm_scope("linear_wgrad", 2.0 * bsz * n * k):
                check(L().qt_linear_wgrad(ptr(x16), k, ptr(dz16), n, ptr(dw), 0, bsz, n, k, ptr(ws), ws.numel(), stream()), "linear_wgrad")
        if ctx.needs_input_grad[2]:
            db = torch.empty(n, device=d.device)
            ops.colsum(dz16, db)
        if ctx.needs_input_grad[0]:
            dx16 = torch.empty(bsz, k, device=d.device, dtype=BF16)
            wd = ops.packed_dgrad(w)
            with ops.gemm_scope("linear_dgrad", 2.0 * bsz * n * k):
                check(L().qt_linear_dgrad(ptr(dz16), n, ptr(wd), ptr(dx16), k, bsz, n, k, ptr(ws), ws.numel(), stream()), "linear_dgrad")
            dx = dx16.float()
        ops._count(4)
        return dx, dw, db, None, None, None


class LSTM(torch.autograd.Function):
    """Multi-layer `nn.LSTM(batch_first=True)` forward over the whole sequence (zero initial state, inter-layer dropout in
    training mode): x [B, T, I] fp32 -> top layer's output sequence [B, T, H] fp32. `weights` = (weight_ih, weight_hh,
    bias_ih, bias_hh) per layer, torch layout and gate order (3dcnn/models.py:144-158, cnn+lstm/models.py:43-49)."""

    @staticmethod
    def forward(ctx, x, p_drop, training, *weights):
        if not x.is_cuda:
            raise RuntimeError("LSTM: the B200 path has no CPU implementation")
        layers = len(weights) // 4
        xin = x.detach().float().contiguous()
        bsz, T, _ = xin.shape
        dev = xin.device
        p = float(p_drop) if training else 0.0
        saved = []
        cur = xin
        for l in range(layers):
            wih, whh, bih, bhh = weights[4 * l: 4 * l + 4]
            G, I = wih.shape
            H = whh.shape[1]
            whh_t = torch.empty(H, G, device=dev)
            check(L().qt_transpose_f32(ptr(whh.detach().contiguous()), ptr(whh_t), G, H, stream()), "lstm transpose")
            in_p = p if l > 0 else 0.0
            seed = ops.new_seed() if in_p > 0 else 0
            x_used = None
            if in_p > 0:  # nn.LSTM's inter-layer dropout: a dropped-out copy of the previous layer's output (counter-hash mask)
                x_used = cur.clone()
                check(L().qt_relu_dropout(ptr(x_used), None, x_used.numel(), in_p, seed, 0, stream()), "lstm dropout")
            xin_l = x_used if x_used is not None else cur
            # input projection of every time step in one batched product: [B*T, I] x Wih^T + b_ih
            xproj = torch.empty(bsz, T, G, device=dev)
            check(L().qt_small_linear_fwd(ptr(xin_l), 0, I, ptr(wih.detach().contiguous()), ptr(bih.detach()) if bih is not None else None,
                                          bsz * T, G, I, 0, 0.0, 0, ptr(xproj), G, None, 0, stream()), "lstm input projection")
            hseq = torch.empty(bsz, T, H, device=dev)
            hprev = torch.empty_like(hseq)
            cseq = torch.empty_like(hseq)
            gates = torch.empty(bsz, T, G, device=dev)
            check(L().qt_lstm_layer_fwd(ptr(xproj), ptr(whh_t), ptr(bhh.detach()) if bhh is not None else None, bsz, T, H, ptr(hseq),
                                        ptr(hprev), ptr(cseq), ptr(gates), stream()), "lstm_layer_fwd")
            ops._count(4)
            saved.append((x_used if x_used is not None else cur, hprev, cseq, gates, in_p, seed))
            cur = hseq
        if any(ctx.needs_input_grad):
            ctx.saved = saved
            ctx.weights = weights
            ctx.dims = (bsz, T)
        return cur

    @staticmethod
    def backward(ctx, dout):
        weights, saved = ctx.weights, ctx.saved
        bsz, T = ctx.dims
        layers = len(saved)
        d = dout.detach().float().contiguous()
        grads = [None] * len(weights)
        out_p, out_seed = 0.0, 0
        dx = None
        for l in reversed(range(layers)):
            wih, whh, bih, bhh = weights[4 * l: 4 * l + 4]
            x_used, hprev, cseq, gates, in_p, seed = saved[l]
            G, I = wih.shape
            H = whh.shape[1]
            dev = gates.device
            dgates = torch.empty_like(gates)
            check(L().qt_lstm_layer_bwd(ptr(d), out_p, out_seed, ptr(whh.detach().contiguous()), ptr(gates), ptr(cseq), bsz, T, H,
                                        ptr(dgates), stream()), "lstm_layer_bwd")
            rows = bsz * T
            need_ih = ctx.needs_input_grad[3 + 4 * l] or (bih is not None and ctx.needs_input_grad[3 + 4 * l + 2])
            if need_ih:
                dwih = ops.grad_out(wih)
                dbih = ops.grad_out(bih) if bih is not None else None
                check(L().qt_small_linear_bwd_dw(ptr(dgates), 0, G, ptr(x_used), 0, I, rows, G, I, ptr(dwih), ptr(dbih), 0, stream()),
                      "lstm dWih")
                grads[4 * l], grads[4 * l + 2] = dwih, dbih
            if ctx.needs_input_grad[3 + 4 * l + 1] or (bhh is not None and ctx.needs_input_grad[3 + 4 * l + 3]):
                dwhh = ops.grad_out(whh)
                dbhh = ops.grad_out(bhh) if bhh is not None else None
                check(L().qt_small_linear_bwd_dw(ptr(dgates), 0, G, ptr(hprev), 0, H, rows, G, H, ptr(dwhh), ptr(dbhh), 0, stream()),
                      "lstm dWhh")
                grads[4 * l + 1], grads[4 * l + 3] = dwhh, dbhh
            ops._count(3)
            if l > 0 or ctx.needs_input_grad[0]:
                dxl = torch.empty(bsz, T, I, device=dev)
                check(L().qt_small_linear_bwd_dx(ptr(dgates), 0, G, ptr(wih.detach().contiguous()), rows, G, I, None, 0, 0.0, 0, ptr(dxl), I,
                                                 None, 0, stream()), "lstm dX")
                ops._count()
                d, out_p, out_seed = dxl, in_p, seed  # gradient w.r.t. the (dropped-out) input = the layer below's output
                if l == 0:
                    dx = dxl
        return (dx, None, None) + tuple(grads)


def lstm_forward(module, x, training=None):
    """Run an `nn.LSTM` module's parameters through the LSTM Function (same state_dict; batch_first, unidirectional)."""
    if module.bidirectional or not module.batch_first or getattr(module, "proj_size", 0):
        raise RuntimeError("LSTM: only batch_first, unidirectional, projection-free LSTMs (what the reference builds) are implemented")
    ws = []
    for l in range(module.num_layers):
        ws += [getattr(module, f"weight_ih_l{l}"), getattr(module, f"weight_hh_l{l}"),
               getattr(module, f"bias_ih_l{l}") if module.bias else None, getattr(module, f"bias_hh_l{l}") if module.bias else None]
    tr = module.training if training is None else training
    return LSTM.apply(x, float(module.dropout), tr, *ws)


class AttnPool(torch.autograd.Function):
    """softmax(scores) weighted sum of R region vectors (QS/models.py:86-90): x fp32 [B,R,C], scores [B,R]."""

    @staticmethod
    def forward(ctx, x, scores):
        xc = x.detach().float().contiguous()
        sc = scores.detach().float().contiguous()
        bsz, r, c = xc.shape
        wts = torch.empty(bsz, r, device=xc.device)
        out = torch.empty(bsz, c, device=xc.device)
        check(L().qt_attn_pool_fwd(ptr(xc), ptr(sc), ptr(wts), ptr(out), bsz, r, c, stream()), "attn_pool_fwd")
        ops._count()
        ctx.save_for_backward(xc, wts)
        return out

    @staticmethod
    def backward(ctx, dout):
        xc, wts = ctx.saved_tensors
        bsz, r, c = xc.shape
        d = dout.detach().float().contiguous()
        dx = torch.empty_like(xc)
        ds = torch.empty(bsz, r, device=xc.device)
        check(L().qt_attn_pool_bwd(ptr(xc), ptr(wts), ptr(d), ptr(dx), ptr(ds), bsz, r, c, stream()), "attn_pool_bwd")
        ops._count()
        return dx, ds


# ------------------------------------------------------------------------------------------------------------
# Conv3d + BatchNorm3d + ReLU (+ MaxPool3d) block of Quadtree3DCNN (3dcnn/models.py:107-141)
# ------------------------------------------------------------------------------------------------------------
def _conv3d_weights(w, cin_pad, need_dgrad, pair):
    """bf16 GEMM operands of a Conv3d weight [cout, cin, 3, 3, 3], cached per parameter version (ops._entry): the forward
    layout [cout][27][cin_pad] (or the pair layout [cout][2][9][64] of the 32-channel slab path) and, when the input needs a
    gradient, the data-gradient layout [cin_pad][27][cout]. Channel-padded weights (cin = 3 -> 8) are packed from a padded
    fp32 copy; everything else goes through ops.packed_* (and is refreshed by the optimizer kernel)."""
    cout, cin = w.shape[:2]
    if cin == cin_pad:
        wf = ops.packed_pair(w) if pair else ops.packed_fprop(w)
        wd = ops.packed_dgrad(w) if need_dgrad else None
        return wf, wd
    e = ops._entry(w)
    if pair == "c8":  # first layer on the dedicated kernel: resident operand tiles straight from the fp32 parameter
        if getattr(e, "w8", None) is None:
            e.w8 = torch.empty(18, 2, 32, 8, device=w.device, dtype=BF16)
            check(L().qt_wpack_conv3d_c8(ptr(w.detach().contiguous()), ptr(e.w8), cout, cin, stream()), "wpack_conv3d_c8")
            ops._count()
        return e.w8, None
    if getattr(e, "w8", None) is None:  # (the stem slot doubles as the cache of the padded pack)
        wp = torch.zeros(cout, cin_pad, 27, device=w.device, dtype=torch.float32)
        wp[:, :cin] = w.detach().reshape(cout, cin, 27)
        wf = torch.empty(cout, 27, cin_pad, device=w.device, dtype=BF16)
        check(L().qt_wpack_both(ptr(wp), ptr(wf), None, cout, cin_pad, 27, stream()), "wpack_both")
        ops._count()
        e.w8 = wf
    return e.w8, None


class Conv3dBnReluPool(torch.autograd.Function):
    """x: NDHWC bf16 [N,D,H,W,Cin_pad] -> NDHWC bf16 after Conv3d(3x3x3, pad 1, bias) + BN3d(train/eval) + ReLU + pool
    (3dcnn/models.py:107-141). Layers with 32 / 64k input channels run on the persistent slab kernels (conv3x3.cuh /
    wgrad3x3.cuh with three depth taps); the 3-channel first layer is HBM bound and uses the gather kernel."""

    @staticmethod
    def forward(ctx, x, w, b, gamma, beta, bn_mod, pool, training):
        n, D, H, W, cin_pad = x.shape
        cout = w.shape[0]
        dev = x.device
        d = capi.conv_desc(n, (D, H, W), cin_pad, cout, (3, 3, 3), (1, 1, 1), (1, 1, 1))
        need_x_grad = ctx.needs_input_grad[0]
        plan = L().qt_conv_plan(d, 0)
        wf, wd = _conv3d_weights(w, cin_pad, need_x_grad, "c8" if plan == 3 else plan == 2)
        y = torch.empty(n, D, H, W, cout, device=dev, dtype=BF16)
        stats = ops.conv_fprop(d, x, wf, y, bias=b.detach(), relu=False, want_stats=training)
        m = n * D * H * W
        st = ops.bn_finalize(stats, m, bn_mod, cout, dev, training)
        need_grad = any(ctx.needs_input_grad)
        a = am = yarg = None
        fused = pool is not None and _pool_fusable(pool, D, H, W)
        if fused:
            # BatchNorm + ReLU + MaxPool3d in one pass: the full-size activation is never written
            kd, kh, kw = pool
            out = torch.empty(n, D // kd, H // kh, W // kw, cout, device=dev, dtype=BF16)
            if need_grad:
                yarg = torch.empty_like(out)
                am = torch.empty(out.shape, device=dev, dtype=torch.int8)
            check(L().qt_bn_relu_maxpool3d_fwd(ptr(y), ptr(st.scale), ptr(st.shift), ptr(out), ptr(yarg), ptr(am), n, D, H, W, cout,
                                               kd, kh, kw, stream()), "bn_relu_maxpool3d_fwd")
            ops._count()
        else:
            a = torch.empty_like(y)
            ops.bn_apply(y, st, a, None, True)
            if pool is not None:
                kd, kh, kw = pool
                out = torch.empty(n, D // kd, H // kh, W // kw, cout, device=dev, dtype=BF16)
                am = torch.empty(out.shape, device=dev, dtype=torch.int8) if need_grad else None
                check(L().qt_maxpool3d_fwd(ptr(a), ptr(out), ptr(am), n, D, H, W, cout, kd, kh, kw, stream()), "maxpool3d_fwd")
                ops._count()
            else:
                # a distinct tensor object: returning `a` itself would make ctx.saved reference the Function's own output
                # (output -> grad_fn -> ctx -> saved -> output), a cycle only Python's GC can break — the whole step's
                # activations would stay allocated until it runs
                out = a.view(a.shape)
        if need_grad:
            ctx.saved = (x, y, a, am, yarg, st, w, wd, gamma, beta, b, d)
            ctx.cfg = (pool, training, cin_pad, fused)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, y, a, am, yarg, st, w, wd, gamma, beta, b, d = ctx.saved
        pool, training, cin_pad, fused = ctx.cfg
        n, D, H, W, cout = y.shape
        dev = y.device
        dout = dout.to(BF16).contiguous()
        dgamma, dbeta = ops.grad_out(gamma), ops.grad_out(beta)
        db = ops.grad_out(b)
        dy = torch.empty_like(y)
        if fused:
            kd, kh, kw = pool
            ws = ops.workspace(L().qt_bn_workspace_bytes(cout), dev, "bn")
            check(L().qt_bn_relu_maxpool3d_bwd(ptr(dout), ptr(am), ptr(y), ptr(yarg), ptr(st.scale), ptr(st.shift), ptr(st.mean), ptr(st.invstd),
                                               ptr(gamma.detach()), n, D, H, W, cout, kd, kh, kw, ptr(dgamma), ptr(dbeta), ptr(db),
                                               0 if training else 1, ptr(dy), ptr(ws), ws.numel(), stream()),
                  "bn_relu_maxpool3d_bwd")
            ops._count(3)
        else:
            if pool is not None:
                kd, kh, kw = pool
                da = torch.empty_like(y)
                check(L().qt_maxpool3d_bwd(ptr(dout), ptr(am), ptr(da), n, D, H, W, cout, kd, kh, kw, stream()), "maxpool3d_bwd")
                ops._count()
            else:
                da = dout
            ops.bn_backward(da, a, y, st, gamma.detach(), dgamma, dbeta, dy, None, eval_mode=not training)
            ops.colsum(dy.view(-1, cout), db)  # conv bias before train-mode BN: analytically ~0, computed faithfully
        dw = None
        if ctx.needs_input_grad[1]:
            cin = w.shape[1]
            if cin == cin_pad:
                dw = ops.grad_out(w)
                ops.conv_wgrad(d, x, dy, dw)
            else:
                dwp = torch.empty(cout, cin_pad, 27, device=dev)
                ops.conv_wgrad(d, x, dy, dwp)
                dw = dwp[:, :cin].reshape(w.shape).contiguous()
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            ops.conv_dgrad(d, dy, wd, dx)
        return dx, dw, db, dgamma, dbeta, None, None, None


def _pool_fusable(pool, D, H, W):
    """qt_bn_relu_maxpool3d_* cover the pools the reference uses ((1,2,2) and (2,2,2), 3dcnn/models.py:111-135) on even maps."""
    kd, kh, kw = pool
    return os.environ.get("QTCNN_NO_POOL_FUSION", "") != "1" and kd in (1, 2) and kh == 2 and kw == 2 and D % kd == 0 \
        and H % 2 == 0 and W % 2 == 0


class PackClip(torch.autograd.Function):
    """[B,T,3,H,W] fp32 clip -> NDHWC bf16 [B,T,H,W,8] (channel padded); the permute(0,2,1,3,4) of
    3dcnn/models.py:189 is only a change of logical axis order."""

    @staticmethod
    def forward(ctx, clips):
        bsz, t, c, h, w = clips.shape
        xf, dtype, scale, shift = ops.stem_source(clips)
        out = torch.empty(bsz, t, h, w, 8, device=clips.device, dtype=BF16)
        check(L().qt_nchw_to_nhwc_bf16_ex(ptr(xf), dtype, ptr(scale), ptr(shift), ptr(out), bsz * t, c, h * w, 8, stream()), "pack clip")
        ops._count()
        return out

    @staticmethod
    def backward(ctx, g):  # inputs never require gradients in the reference scripts
        return None


class GlobalAvgPoolND(torch.autograd.Function):
    """AdaptiveAvgPool over all spatial positions of a channels-last bf16 tensor [N, ..., C] -> fp32 [N, C]."""

    @staticmethod
    def forward(ctx, x):
        n, c = x.shape[0], x.shape[-1]
        p = x.numel() // (n * c)
        out = torch.empty(n, c, device=x.device, dtype=BF16)
        check(L().qt_region_avgpool_fwd(ptr(x), ptr(out), n, p, c, c, stream()), "avgpool")
        ops._count()
        ctx.shape = tuple(x.shape)
        return out.float()

    @staticmethod
    def backward(ctx, dout):
        shape = ctx.shape
        n, c = shape[0], shape[-1]
        p = 1
        for s in shape[1:-1]:
            p *= s
        d = dout.to(BF16).contiguous()
        dx = torch.empty(shape, device=d.device, dtype=BF16)
        check(L().qt_region_avgpool_bwd(ptr(d), ptr(dx), ptr(dx), n, p, c, c, 0, stream()), "avgpool_bwd")
        ops._count()
        return dx
