"""Per-kernel parity tests (need a B200): every C-ABI entry point is compared with a plain PyTorch fp32
evaluation of the same operator on the same bf16-rounded inputs.

Tolerances: bf16 outputs rel_L2 <= 4e-3 and max_abs <= 2^-7 * max|ref| (one bf16 ulp of the
largest value, plus accumulation-order noise); fp32 outputs (statistics, weight gradients, logits)
rel_L2 <= 1e-4 (fp32 accumulation in a different order than cuDNN/cuBLAS). Index work (pool argmax, quadrant
assignment, feature-row offsets) is bit-exact.
"""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

BF16_REL_L2 = 4e-3
F32_REL_L2 = 1e-4


@pytest.fixture(scope="module")
def C():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    import qtcnn_b200.capi as capi
    capi.lib()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return capi


def rel_l2(a, b):
    a = a.double().flatten()
    b = b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


def report(name, got, ref, tol, bf16_out=False):
    got = got.float()
    ref = ref.float()
    err = rel_l2(got, ref)
    maxabs = float((got - ref).abs().max())
    refmax = float(ref.abs().max())
    msg = f"{name}: rel_l2={err:.3e} max_abs={maxabs:.3e} ref_max={refmax:.3e} shape={tuple(ref.shape)}"
    print(msg)
    if not (err <= tol) or not math.isfinite(err):
        # coarse error map to make a single remote run diagnosable
        d = (got - ref).abs().reshape(ref.shape[0], -1) if ref.dim() > 1 else (got - ref).abs().reshape(1, -1)
        rows = d.max(dim=1).values
        bad_rows = (rows > 10 * tol * max(refmax, 1e-6)).nonzero().flatten()[:32].tolist()
        cols = d.max(dim=0).values
        bad_cols = (cols > 10 * tol * max(refmax, 1e-6)).nonzero().flatten()[:32].tolist()
        pytest.fail(msg + f"\n first bad rows {bad_rows}\n first bad cols {bad_cols}\n got[0,:8]={got.flatten()[:8].tolist()}\n"
                    f" ref[0,:8]={ref.flatten()[:8].tolist()}")
    if bf16_out:
        assert maxabs <= 2.0 ** -7 * refmax + 1e-6, msg


def bf16(t):
    return t.to(torch.bfloat16)


def run(C, rc, what):
    C.check(rc, what)
    torch.cuda.synchronize()
    assert C.lib().qt_take_timeout_flag() == 0, f"{what}: a barrier wait timed out inside the kernel"


def nhwc(t):  # logical NCHW tensor -> dense NHWC bf16 buffer
    return bf16(t).permute(0, 2, 3, 1).contiguous()


def conv_case(C, n, cin, cout, h, w, k, stride, pad, bias=False, relu=False, stats=False, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn(n, cin, h, w, device="cuda", generator=g)
    wt = torch.randn(cout, cin, k, k, device="cuda", generator=g) / math.sqrt(cin * k * k)
    b = torch.randn(cout, device="cuda", generator=g) if bias else None
    xb, wb = bf16(x).float(), bf16(wt).float()
    ref = F.conv2d(xb, wb, b, stride=stride, padding=pad)
    if relu:
        ref = ref.relu()
    ho, wo = ref.shape[2], ref.shape[3]
    x_nhwc = nhwc(x)
    wf = torch.empty(cout, k * k, cin, device="cuda", dtype=torch.bfloat16)
    run(C, C.lib().qt_wpack_fprop(C.ptr(wt.contiguous()), C.ptr(wf), cout, cin, k * k, C.stream()), "wpack_fprop")
    assert torch.equal(wf, bf16(wt).permute(0, 2, 3, 1).reshape(cout, k * k, cin)), "wpack_fprop layout"
    d = C.conv_desc(n, (1, h, w), cin, cout, (1, k, k), (1, stride, stride), (0, pad, pad))
    y = torch.full((n, ho, wo, cout), float("nan"), device="cuda", dtype=torch.bfloat16)
    rows = C.lib().qt_conv_stat_rows(d)
    st = torch.zeros(rows, 2, cout, device="cuda") if stats else None
    flags = (C.QT_EPI_BIAS if bias else 0) | (C.QT_EPI_RELU if relu else 0) | (C.QT_EPI_STATS if stats else 0)
    run(C, C.lib().qt_conv_fprop(d, C.ptr(x_nhwc), C.ptr(wf), C.ptr(y), C.ptr(b), C.ptr(st), flags, None, 0, C.stream()),
        "conv_fprop")
    report(f"conv_fprop n{n} c{cin}->{cout} {h}x{w} k{k} s{stride} p{pad}", y.permute(0, 3, 1, 2), ref, BF16_REL_L2, True)
    if stats:
        yf = y.float().reshape(-1, cout)
        report("  stats sum", st[:, 0].sum(0), yf.sum(0), 1e-4)
        report("  stats sumsq", st[:, 1].sum(0), (yf * yf).sum(0), 1e-4)
    return x, wt, ref


@pytest.mark.parametrize("b,n,k", [(128, 64, 64), (128, 128, 64), (256, 128, 256), (100, 200, 128), (256, 2688, 5376)])
def test_linear_fprop(C, b, n, k):
    g = torch.Generator(device="cuda").manual_seed(1)
    x = bf16(torch.randn(b, k, device="cuda", generator=g))
    w = bf16(torch.randn(n, k, device="cuda", generator=g) / math.sqrt(k))
    bias = torch.randn(n, device="cuda", generator=g)
    ref = (x.float() @ w.float().t() + bias).relu()
    ws_bytes = C.lib().qt_linear_workspace_bytes(b, n, k)
    ws = torch.empty(ws_bytes, device="cuda", dtype=torch.uint8)
    out = torch.full((b, n), float("nan"), device="cuda")
    run(C, C.lib().qt_linear_fprop(C.ptr(x), k, C.ptr(w), C.ptr(bias), C.ptr(out), n,
                                   C.QT_EPI_BIAS | C.QT_EPI_RELU | C.QT_EPI_OUT_F32, b, n, k, C.ptr(ws), ws_bytes,
                                   C.stream()), "linear_fprop")
    report(f"linear_fprop f32 {b}x{n}x{k}", out, ref, F32_REL_L2)
    out16 = torch.full((b, n), float("nan"), device="cuda", dtype=torch.bfloat16)
    run(C, C.lib().qt_linear_fprop(C.ptr(x), k, C.ptr(w), C.ptr(bias), C.ptr(out16), n, C.QT_EPI_BIAS | C.QT_EPI_RELU,
                                   b, n, k, C.ptr(ws), ws_bytes, C.stream()), "linear_fprop bf16")
    report(f"linear_fprop bf16 {b}x{n}x{k}", out16, ref, BF16_REL_L2, True)


@pytest.mark.parametrize("b,n,k", [(128, 64, 128), (256, 2688, 5376), (96, 200, 256)])
def test_linear_dgrad_wgrad(C, b, n, k):
    g = torch.Generator(device="cuda").manual_seed(2)
    x = bf16(torch.randn(b, k, device="cuda", generator=g))
    w = torch.randn(n, k, device="cuda", generator=g) / math.sqrt(k)
    dy = bf16(torch.randn(b, n, device="cuda", generator=g))
    wt = torch.empty(k, n, device="cuda", dtype=torch.bfloat16)
    run(C, C.lib().qt_wpack_dgrad(C.ptr(w), C.ptr(wt), n, k, 1, C.stream()), "wpack_dgrad")
    assert torch.equal(wt, bf16(w).t().contiguous()), "wpack_dgrad layout"
    ws_bytes = C.lib().qt_linear_workspace_bytes(b, n, k)
    ws = torch.empty(ws_bytes, device="cuda", dtype=torch.uint8)
    dx = torch.full((b, k), float("nan"), device="cuda", dtype=torch.bfloat16)
    run(C, C.lib().qt_linear_dgrad(C.ptr(dy), n, C.ptr(wt), C.ptr(dx), k, b, n, k, C.ptr(ws), ws_bytes, C.stream()),
        "linear_dgrad")
    report(f"linear_dgrad {b}x{n}x{k}", dx, dy.float() @ bf16(w).float(), BF16_REL_L2, True)
    dw = torch.full((n, k), float("nan"), device="cuda")
    run(C, C.lib().qt_linear_wgrad(C.ptr(x), k, C.ptr(dy), n, C.ptr(dw), 0, b, n, k, C.ptr(ws), ws_bytes, C.stream()),
        "linear_wgrad")
    report(f"linear_wgrad {b}x{n}x{k}", dw, dy.float().t() @ x.float(), F32_REL_L2)


@pytest.mark.parametrize("n,cin,cout,h,w,k,stride,pad", [
    (2, 64, 64, 8, 8, 3, 1, 1),
    (3, 64, 64, 56, 56, 3, 1, 1),
    (4, 64, 128, 28, 28, 3, 2, 1),
    (4, 64, 128, 28, 28, 1, 2, 0),
    (5, 128, 128, 14, 14, 3, 1, 1),
    (8, 256, 512, 14, 14, 3, 2, 1),
    (6, 512, 512, 7, 7, 3, 1, 1),
    (2, 32, 64, 9, 11, 3, 1, 1),
    (7, 512, 512, 2, 2, 3, 1, 1),      # tiny / odd maps (what 64x64 inputs leave at layer4)
    (5, 256, 256, 4, 4, 3, 1, 1),
    (5, 128, 128, 3, 5, 3, 1, 1),
    (3, 64, 64, 5, 3, 3, 1, 1),
])
def test_conv_fprop(C, n, cin, cout, h, w, k, stride, pad):
    conv_case(C, n, cin, cout, h, w, k, stride, pad, stats=True)


def test_conv_fprop_bias_relu(C):
    conv_case(C, 4, 256, 128, 7, 7, 3, 1, 1, bias=True, relu=True)


@pytest.mark.parametrize("n,cin,cout,h,w,k,stride,pad", [
    (2, 64, 64, 8, 8, 3, 1, 1),
    (3, 64, 64, 30, 30, 3, 1, 1),
    (7, 512, 512, 2, 2, 3, 1, 1),      # tiny maps: the slab kernels' single-carry advance does not apply -> generic kernels
    (5, 256, 256, 4, 4, 3, 1, 1),
    (5, 128, 128, 3, 5, 3, 1, 1),
    (3, 128, 256, 28, 28, 3, 1, 1),
    (4, 64, 128, 28, 28, 3, 2, 1),
    (4, 64, 128, 28, 28, 1, 2, 0),
    (6, 512, 512, 7, 7, 3, 1, 1),
    (5, 256, 128, 7, 7, 3, 1, 1),
])
def test_conv_dgrad_wgrad(C, n, cin, cout, h, w, k, stride, pad):
    g = torch.Generator(device="cuda").manual_seed(3)
    x = bf16(torch.randn(n, cin, h, w, device="cuda", generator=g)).float().requires_grad_(True)
    wt = torch.randn(cout, cin, k, k, device="cuda", generator=g) / math.sqrt(cin * k * k)
    wq = bf16(wt).float().requires_grad_(True)
    y = F.conv2d(x, wq, None, stride=stride, padding=pad)
    dy = bf16(torch.randn_like(y))
    y.backward(dy.float())
    ho, wo = y.shape[2], y.shape[3]
    d = C.conv_desc(n, (1, h, w), cin, cout, (1, k, k), (1, stride, stride), (0, pad, pad))
    dy_nhwc = dy.permute(0, 2, 3, 1).contiguous()
    x_nhwc = nhwc(x.detach())
    wd = torch.empty(cin, k * k, cout, device="cuda", dtype=torch.bfloat16)
    run(C, C.lib().qt_wpack_dgrad(C.ptr(wt.contiguous()), C.ptr(wd), cout, cin, k * k, C.stream()), "wpack_dgrad")
    assert torch.equal(wd, bf16(wt).permute(1, 2, 3, 0).reshape(cin, k * k, cout)), "wpack_dgrad layout"
    # dgrad (fresh) and dgrad accumulated onto an existing gradient
    dx = torch.zeros(n, h, w, cin, device="cuda", dtype=torch.bfloat16)
    acc = 1 if (stride > 1 and k == 1) else 0
    run(C, C.lib().qt_conv_dgrad(d, C.ptr(dy_nhwc), C.ptr(wd), C.ptr(dx), acc, C.stream()), "conv_dgrad")
    report(f"conv_dgrad c{cin}->{cout} {h}x{w} k{k} s{stride}", dx.permute(0, 3, 1, 2), x.grad, BF16_REL_L2, True)
    base = bf16(torch.randn(n, h, w, cin, device="cuda", generator=g))
    dx2 = base.clone()
    run(C, C.lib().qt_conv_dgrad(d, C.ptr(dy_nhwc), C.ptr(wd), C.ptr(dx2), 1, C.stream()), "conv_dgrad accumulate")
    report("  conv_dgrad accumulate", dx2.permute(0, 3, 1, 2), x.grad + base.float().permute(0, 3, 1, 2), 6e-3)
    # wgrad
    ws_bytes = C.lib().qt_conv_wgrad_workspace_bytes(d)
    ws = torch.empty(max(ws_bytes, 16), device="cuda", dtype=torch.uint8)
    dw = torch.full((cout, cin, k, k), float("nan"), device="cuda")
    run(C, C.lib().qt_conv_wgrad(d, C.ptr(x_nhwc), C.ptr(dy_nhwc), C.ptr(dw), 0, C.ptr(ws), ws_bytes, C.stream()),
        "conv_wgrad")
    report(f"conv_wgrad c{cin}->{cout} {h}x{w} k{k} s{stride}", dw, wq.grad, F32_REL_L2)


def test_quadrant_conv_groups(C):
    """The four quadrant views of a 14x14 map as one grouped launch (zero halo at the internal seams)."""
    n, cin, cout = 3, 256, 128
    g = torch.Generator(device="cuda").manual_seed(4)
    base = bf16(torch.randn(n, cin, 14, 14, device="cuda", generator=g))
    wt = torch.randn(cout, cin, 3, 3, device="cuda", generator=g) / math.sqrt(cin * 9)
    b = torch.randn(cout, device="cuda", generator=g)
    quads = [base[:, :, :7, :7], base[:, :, :7, 7:], base[:, :, 7:, :7], base[:, :, 7:, 7:]]
    ref = torch.stack([F.conv2d(q.float(), bf16(wt).float(), b, padding=1).relu() for q in quads])  # [4,n,cout,7,7]
    x_nhwc = base.permute(0, 2, 3, 1).contiguous()
    wf = torch.empty(cout, 9, cin, device="cuda", dtype=torch.bfloat16)
    run(C, C.lib().qt_wpack_fprop(C.ptr(wt), C.ptr(wf), cout, cin, 9, C.stream()), "wpack")
    xs = (14 * 14 * cin, 0, 14 * cin, cin)
    xoff = (0, 7 * cin, 7 * 14 * cin, 7 * 14 * cin + 7 * cin)
    yoff = tuple(q * n * 49 * cout for q in range(4))
    d = C.conv_desc(n, (1, 7, 7), cin, cout, (1, 3, 3), (1, 1, 1), (0, 1, 1), x_stride=xs, groups=4, x_group_off=xoff,
                    y_group_off=yoff)
    y = torch.full((4, n, 7, 7, cout), float("nan"), device="cuda", dtype=torch.bfloat16)
    run(C, C.lib().qt_conv_fprop(d, C.ptr(x_nhwc), C.ptr(wf), C.ptr(y), C.ptr(b), None,
                                 C.QT_EPI_BIAS | C.QT_EPI_RELU, None, 0, C.stream()), "quadrant conv")
    report("quadrant conv fprop", y.permute(0, 1, 4, 2, 3), ref, BF16_REL_L2, True)


@pytest.mark.parametrize("n,h,w", [(3, 64, 96), (2, 224, 224), (1, 38, 50)])
def test_stem(C, n, h, w):
    """(64x96, 224x224): dedicated overlapping-row kernels; 38x50 (Wo % 16 != 0): generic weight-gradient path."""
    cout = 64
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn(n, 3, h, w, device="cuda", generator=g)
    wt = torch.randn(cout, 3, 7, 7, device="cuda", generator=g) / math.sqrt(147)
    xq = bf16(x).float().requires_grad_(True)
    wq = bf16(wt).float().requires_grad_(True)
    ref = F.conv2d(xq, wq, None, stride=2, padding=3)
    xp = torch.empty(n, h + 7, w + 8, 4, device="cuda", dtype=torch.bfloat16)
    run(C, C.lib().qt_stem_pack_input(C.ptr(x), C.ptr(xp), n, 3, h, w, C.stream()), "stem_pack_input")
    expect = torch.zeros(n, h + 7, w + 8, 4, device="cuda", dtype=torch.bfloat16)   # 3 rows / cols of zeros in front
    expect[:, 3:3 + h, 3:3 + w, :3] = bf16(x).permute(0, 2, 3, 1)
    assert torch.equal(xp, expect), "stem_pack_input layout"
    w8 = torch.empty(cout, 8, 32, device="cuda", dtype=torch.bfloat16)
    run(C, C.lib().qt_wpack_stem(C.ptr(wt), C.ptr(w8), cout, 3, 7, 7, C.stream()), "wpack_stem")
    y = torch.full((n, h // 2, w // 2, cout), float("nan"), device="cuda", dtype=torch.bfloat16)
    rows = C.lib().qt_stem_stat_rows(n, h, w)
    st = torch.zeros(rows, 2, cout, device="cuda")
    run(C, C.lib().qt_stem_fprop(C.ptr(xp), C.ptr(w8), C.ptr(y), C.ptr(st), n, h, w, cout, C.stream()), "stem_fprop")
    report("stem fprop", y.permute(0, 3, 1, 2), ref, BF16_REL_L2, True)
    report("  stem stats", st[:, 0].sum(0), y.float().reshape(-1, cout).sum(0), 1e-4)
    dy = bf16(torch.randn_like(ref))
    ref.backward(dy.float())
    ws_bytes = C.lib().qt_stem_wgrad_workspace_bytes(n, h, w, cout)
    ws = torch.empty(ws_bytes, device="cuda", dtype=torch.uint8)
    dw = torch.full((cout, 3, 7, 7), float("nan"), device="cuda")
    run(C, C.lib().qt_stem_wgrad(C.ptr(xp), C.ptr(dy.permute(0, 2, 3, 1).contiguous()), C.ptr(dw), 0, n, h, w, cout, 3,
                                 C.ptr(ws), ws_bytes, C.stream()), "stem_wgrad")
    report("stem wgrad", dw, wq.grad, F32_REL_L2)


@pytest.mark.parametrize("m,c,residual,relu", [(1000, 64, False, True), (4096, 256, True, True), (300, 512, True, False)])
def test_batchnorm_train(C, m, c, residual, relu):
    g = torch.Generator(device="cuda").manual_seed(6)
    y = bf16(torch.randn(m, c, device="cuda", generator=g) * 2 + 0.5)
    res = bf16(torch.randn(m, c, device="cuda", generator=g)) if residual else None
    gamma = torch.rand(c, device="cuda", generator=g) + 0.5
    beta = torch.randn(c, device="cuda", generator=g)
    rm, rv = torch.zeros(c, device="cuda"), torch.ones(c, device="cuda")
    # reference
    yf = y.float().requires_grad_(True)
    rm_ref, rv_ref = rm.clone(), rv.clone()
    z = F.batch_norm(yf, rm_ref, rv_ref, gamma.clone().requires_grad_(True), beta, True, 0.1, 1e-5)
    gref = gamma.clone().requires_grad_(True)
    bref = beta.clone().requires_grad_(True)
    z = F.batch_norm(yf, None, None, gref, bref, True, 0.1, 1e-5)
    o = z + res.float() if residual else z
    o = o.relu() if relu else o
    dout = bf16(torch.randn(m, c, device="cuda", generator=g))
    o.backward(dout.float())
    # ours
    rows = 37
    partial = torch.zeros(rows, 2, c, device="cuda")
    run(C, C.lib().qt_bn_stats(C.ptr(y), m, c, C.ptr(partial), rows, C.stream()), "bn_stats")
    ws_bytes = C.lib().qt_bn_workspace_bytes(c)
    ws = torch.empty(ws_bytes, device="cuda", dtype=torch.uint8)
    mean, invstd, scale, shift = (torch.empty(c, device="cuda") for _ in range(4))
    run(C, C.lib().qt_bn_finalize(C.ptr(partial), rows, c, float(m), C.ptr(gamma), C.ptr(beta), 1e-5, 0.1, C.ptr(rm),
                                  C.ptr(rv), C.ptr(mean), C.ptr(invstd), C.ptr(scale), C.ptr(shift), C.ptr(ws), ws_bytes,
                                  C.stream()), "bn_finalize")
    report("bn mean", mean, y.float().mean(0), 1e-5)
    report("bn invstd", invstd, 1.0 / torch.sqrt(y.float().var(0, unbiased=False) + 1e-5), 1e-5)
    report("bn running_mean", rm, rm_ref, 1e-5)
    report("bn running_var", rv, rv_ref, 1e-5)
    out = torch.empty_like(y)
    run(C, C.lib().qt_bn_apply(C.ptr(y), C.ptr(scale), C.ptr(shift), C.ptr(res), C.ptr(out), m, c, int(relu), C.stream()),
        "bn_apply")
    report("bn apply", out, o.detach(), BF16_REL_L2, True)
    dgamma, dbeta = torch.empty(c, device="cuda"), torch.empty(c, device="cuda")
    dy = torch.empty_like(y)
    dz = torch.empty_like(y)
    act = out if relu else None
    run(C, C.lib().qt_bn_backward(C.ptr(dout), C.ptr(act), C.ptr(y), C.ptr(mean), C.ptr(invstd), C.ptr(gamma), None, None, m, c,
                                  C.ptr(dgamma), C.ptr(dbeta), 0, 0, C.ptr(dy), C.ptr(dz), C.ptr(ws), ws_bytes, C.stream()),
        "bn_backward")
    # the reference mask comes from fp32 activations; ours from the bf16-rounded ones -> compare loosely on dy
    report("bn dgamma", dgamma, gref.grad, 2e-2)
    report("bn dbeta", dbeta, bref.grad, 2e-2)
    report("bn dy", dy, yf.grad, 2e-2)


def test_maxpool_3x3s2(C):
    n, h, w, c = 3, 20, 24, 64
    g = torch.Generator(device="cuda").manual_seed(7)
    x = bf16(torch.randn(n, c, h, w, device="cuda", generator=g)).relu()  # many exact ties at 0, like after ReLU
    xf = x.float().requires_grad_(True)
    ref, idx = F.max_pool2d(xf, 3, 2, 1, return_indices=True)
    ho, wo = ref.shape[2], ref.shape[3]
    x_nhwc = x.permute(0, 2, 3, 1).contiguous()
    out = torch.empty(n, ho, wo, c, device="cuda", dtype=torch.bfloat16)
    am = torch.empty(n, ho, wo, c, device="cuda", dtype=torch.int8)
    run(C, C.lib().qt_maxpool2d_fwd(C.ptr(x_nhwc), C.ptr(out), C.ptr(am), n, h, w, c, 3, 2, 1, C.stream()), "maxpool fwd")
    assert torch.equal(out.permute(0, 3, 1, 2).float(), ref.detach()), "maxpool values must be bit-exact"
    # argmax: decode window code -> flat input index, must equal ATen's indices
    amn = am.permute(0, 3, 1, 2).long()
    hh = torch.arange(ho, device="cuda").view(1, 1, -1, 1) * 2 - 1 + amn // 3
    ww = torch.arange(wo, device="cuda").view(1, 1, 1, -1) * 2 - 1 + amn % 3
    assert torch.equal(hh * w + ww, idx), "maxpool argmax must match ATen's first-maximum rule"
    dout = bf16(torch.randn_like(ref))
    ref.backward(dout.float())
    dx = torch.empty(n, h, w, c, device="cuda", dtype=torch.bfloat16)
    run(C, C.lib().qt_maxpool2d_bwd(C.ptr(dout.permute(0, 2, 3, 1).contiguous()), C.ptr(am), C.ptr(dx), n, h, w, c, 3, 2,
                                    1, C.stream()), "maxpool bwd")
    report("maxpool bwd", dx.permute(0, 3, 1, 2), xf.grad, BF16_REL_L2)


def test_quadtree_pool(C):
    b, cq, cg = 5, 128, 512
    g = torch.Generator(device="cuda").manual_seed(8)
    q = bf16(torch.randn(4, b, cq, 7, 7, device="cuda", generator=g)).relu()
    l4 = bf16(torch.randn(b, cg, 7, 7, device="cuda", generator=g)).relu()
    qf = q.float().requires_grad_(True)
    lf = l4.float().requires_grad_(True)
    parts = [F.adaptive_avg_pool2d(lf, 1).flatten(1)] + [F.max_pool2d(qf[i], 2, 2).flatten(1) for i in range(4)]
    ref = torch.cat(parts, dim=1)  # [b, 5120] exactly as models.py:291-294
    ldf = 5376
    feat = torch.zeros(b, ldf, device="cuda", dtype=torch.bfloat16)
    q_nhwc = q.permute(0, 1, 3, 4, 2).contiguous()
    l_nhwc = l4.permute(0, 2, 3, 1).contiguous()
    run(C, C.lib().qt_quadtree_pool_fwd(C.ptr(q_nhwc), C.ptr(l_nhwc), C.ptr(feat), b, 7, 7, cq, 49, cg, ldf, C.stream()),
        "quadtree_pool_fwd")
    assert torch.equal(feat[:, 512:5120].float(), ref[:, 512:].detach()), "quadrant max-pool / flatten order must be bit-exact"
    report("quadtree global avg", feat[:, :512], ref[:, :512], BF16_REL_L2, True)
    assert float(feat[:, 5120:].abs().max()) == 0.0
    dfeat = bf16(torch.randn(b, ldf, device="cuda", generator=g))
    ref.backward(dfeat[:, :5120].float())
    dq = torch.empty_like(q_nhwc)
    dl = torch.empty_like(l_nhwc)
    run(C, C.lib().qt_quadtree_pool_bwd(C.ptr(dfeat), C.ptr(q_nhwc), C.ptr(dq), C.ptr(dl), b, 7, 7, cq, 49, cg, ldf,
                                        C.stream()), "quadtree_pool_bwd")
    # reference gradient w.r.t. the post-ReLU map, then the ReLU mask (q > 0) our kernel fuses
    ref_dq = qf.grad * (q.float() > 0)
    assert torch.equal(dq.permute(0, 1, 4, 2, 3).float(), ref_dq), "quadrant gradient routing must be bit-exact"
    report("quadtree dl4", dl.permute(0, 3, 1, 2), lf.grad, BF16_REL_L2, True)


@pytest.mark.parametrize("b,k,n", [
    (33, 47, 94),       # head MLP: warp-per-output kernels
    (256, 640, 1024),   # CnnLstm input projection over all time steps (cnn+lstm/models.py:43-49): tiled kernel (sgemm.cuh)
    (512, 47, 752),     # Quadtree3DCNN LSTM layer-1 projection at B=32, T=16 (3dcnn/models.py:144-150), ragged K
    (500, 188, 750),    # ragged in every dimension
])
def test_small_linears(C, b, k, n):
    g = torch.Generator(device="cuda").manual_seed(9)
    x = torch.rand(b, k, device="cuda", generator=g) * (180 if k == 47 else 1)
    w = (torch.randn(n, k, device="cuda", generator=g) / math.sqrt(k)).requires_grad_(True)
    bias = torch.randn(n, device="cuda", generator=g).requires_grad_(True)
    xr = x.clone().requires_grad_(True)
    ref = F.linear(xr, w, bias).relu()
    guards = []

    def guarded(*shape, dtype=torch.float32):
        """Output tensor inside a sentinel-filled buffer: an out-of-bounds store of a ragged tile shows up in the guard bands
        (compute-sanitizer is not available on the GPU pool)."""
        numel, pad = math.prod(shape), 4096
        buf = torch.full((numel + 2 * pad,), -7.0, device="cuda", dtype=dtype)
        guards.append((buf, pad, numel))
        return buf[pad:pad + numel].view(*shape)

    out = guarded(b, n)
    out16 = guarded(b, n, dtype=torch.bfloat16)
    run(C, C.lib().qt_small_linear_fwd(C.ptr(x), 0, k, C.ptr(w), C.ptr(bias), b, n, k, 1, 0.0, 0, C.ptr(out), n,
                                       C.ptr(out16), n, C.stream()), "small_linear_fwd")
    report("small_linear fwd", out, ref, 1e-5)
    dy = torch.randn(b, n, device="cuda", generator=g)
    ref.backward(dy)
    # dz through the ReLU using the stored output, then dW/db/dx
    dz = guarded(b, n)
    eye = torch.eye(n, device="cuda")
    run(C, C.lib().qt_small_linear_bwd_dx(C.ptr(dy), 0, n, C.ptr(eye), b, n, n, C.ptr(out), n, 0.0, 0, C.ptr(dz), n, None,
                                          0, C.stream()), "relu mask via bwd_dx")
    dw = guarded(n, k)
    db = guarded(n)
    run(C, C.lib().qt_small_linear_bwd_dw(C.ptr(dz), 0, n, C.ptr(x), 0, k, b, n, k, C.ptr(dw), C.ptr(db), 0, C.stream()),
        "small_linear_bwd_dw")
    report("small_linear dw", dw, w.grad, 1e-5)
    report("small_linear db", db, bias.grad, 1e-5)
    dx = guarded(b, k)
    run(C, C.lib().qt_small_linear_bwd_dx(C.ptr(dz), 0, n, C.ptr(w), b, n, k, None, 0, 0.0, 0, C.ptr(dx), k, None, 0,
                                          C.stream()), "small_linear_bwd_dx")
    report("small_linear dx", dx, xr.grad, 1e-5)
    for buf, pad, numel in guards:
        assert bool((buf[:pad] == -7.0).all()) and bool((buf[pad + numel:] == -7.0).all()), "store outside the output tensor"


def test_tiled_linear_dropout_and_bf16_operands(C):
    """Tiled path (sgemm.cuh): counter-hash dropout is keyed on the element index exactly like the warp-per-output kernels (the
    backward gate reads the stored output), bf16-stored operands, accumulate into dw / db."""
    b, k, n = 128, 256, 192
    g = torch.Generator(device="cuda").manual_seed(3)
    x16 = torch.randn(b, k, device="cuda", generator=g).to(torch.bfloat16)
    w = torch.randn(n, k, device="cuda", generator=g) / math.sqrt(k)
    out = torch.empty(b, n, device="cuda")
    out16 = torch.empty(b, n, device="cuda", dtype=torch.bfloat16)
    lib = C.lib()
    run(C, lib.qt_small_linear_fwd(C.ptr(x16), 1, k, C.ptr(w), None, b, n, k, 1, 0.5, 77, C.ptr(out), n, C.ptr(out16), n, C.stream()),
        "fwd")
    zp = x16.float() @ w.t()
    z = zp.relu()
    clear = zp.abs() > 1e-4   # away from the ReLU edge, where summation order could flip the sign
    kept = out > 0
    assert 0.2 < float(kept.float().mean()) < 0.3   # half by ReLU, half of those by dropout
    report("tiled fwd kept*2", out[kept & clear], 2 * z[kept & clear], 1e-5)
    assert torch.equal(out16, out.to(torch.bfloat16))
    # same mask as the element-wise kernel keyed on the same index
    h = z.clone()
    run(C, lib.qt_relu_dropout(C.ptr(h), None, b * n, 0.5, 77, 1, C.stream()), "relu_dropout")
    assert torch.equal((h > 0)[clear], kept[clear])
    # gate of the backward = stored output; eye weight isolates the gate
    dy16 = torch.randn(b, n, device="cuda", generator=g).to(torch.bfloat16)
    eye = torch.eye(n, device="cuda")
    dz = torch.empty(b, n, device="cuda")
    run(C, lib.qt_small_linear_bwd_dx(C.ptr(dy16), 1, n, C.ptr(eye), b, n, n, C.ptr(out), n, 0.5, 77, C.ptr(dz), n, None, 0, C.stream()),
        "gate")
    assert torch.equal(dz, torch.where(kept, 2 * dy16.float(), torch.zeros_like(dz)))
    dw = torch.ones(n, k, device="cuda")
    db = torch.ones(n, device="cuda")
    run(C, lib.qt_small_linear_bwd_dw(C.ptr(dy16), 1, n, C.ptr(x16), 1, k, b, n, k, C.ptr(dw), C.ptr(db), 1, C.stream()), "dw acc")
    report("tiled dw (accumulate)", dw, 1 + dy16.float().t() @ x16.float(), 1e-5)
    report("tiled db (accumulate)", db, 1 + dy16.float().sum(0), 1e-5)


def test_dropout_statistics(C):
    n = 1 << 20
    h = torch.ones(n, device="cuda")
    h16 = torch.empty(n, device="cuda", dtype=torch.bfloat16)
    run(C, C.lib().qt_relu_dropout(C.ptr(h), C.ptr(h16), n, 0.5, 1234, 1, C.stream()), "relu_dropout")
    keep = float((h > 0).float().mean())
    assert abs(keep - 0.5) < 5e-3, keep
    assert set(h.unique().tolist()) == {0.0, 2.0}
    h2 = torch.ones(n, device="cuda")
    run(C, C.lib().qt_relu_dropout(C.ptr(h2), None, n, 0.5, 1234, 1, C.stream()), "relu_dropout again")
    assert torch.equal(h, h2), "same seed -> same mask (backward regenerates it)"


def test_region_avgpool_and_misc(C):
    r, p, c = 40, 49, 64
    g = torch.Generator(device="cuda").manual_seed(10)
    x = bf16(torch.randn(r, p, c, device="cuda", generator=g)).relu()
    out = torch.empty(r, c, device="cuda", dtype=torch.bfloat16)
    run(C, C.lib().qt_region_avgpool_fwd(C.ptr(x), C.ptr(out), r, p, c, c, C.stream()), "region_avgpool_fwd")
    report("region avgpool", out, x.float().mean(1), BF16_REL_L2, True)
    dout = bf16(torch.randn(r, c, device="cuda", generator=g))
    dx = torch.empty_like(x)
    run(C, C.lib().qt_region_avgpool_bwd(C.ptr(dout), C.ptr(x), C.ptr(dx), r, p, c, c, 1, C.stream()), "region_avgpool_bwd")
    report("region avgpool bwd", dx, (dout.float() / p).unsqueeze(1) * (x.float() > 0), BF16_REL_L2, True)
    # colsum / add / relu_backward
    ws_bytes = C.lib().qt_bn_workspace_bytes(c)
    ws = torch.empty(ws_bytes, device="cuda", dtype=torch.uint8)
    cs = torch.empty(c, device="cuda")
    x2 = x.reshape(-1, c)
    run(C, C.lib().qt_colsum(C.ptr(x2), x2.shape[0], c, C.ptr(cs), 0, C.ptr(ws), ws_bytes, C.stream()), "colsum")
    report("colsum", cs, x2.float().sum(0), 1e-5)
    a, b2 = bf16(torch.randn(4096, device="cuda", generator=g)), bf16(torch.randn(4096, device="cuda", generator=g))
    o = torch.empty_like(a)
    run(C, C.lib().qt_add_bf16(C.ptr(a), C.ptr(b2), C.ptr(o), 4096, C.stream()), "add")
    assert torch.equal(o, bf16(a.float() + b2.float()))
    run(C, C.lib().qt_relu_backward(C.ptr(a), C.ptr(b2), C.ptr(o), 4096, C.stream()), "relu_backward")
    assert torch.equal(o, torch.where(b2.float() > 0, a, torch.zeros_like(a)))


def test_colsum_wide(C):
    m, c = 300, 2688
    g = torch.Generator(device="cuda").manual_seed(12)
    x = bf16(torch.randn(m, c, device="cuda", generator=g))
    ws_bytes = C.lib().qt_bn_workspace_bytes(c)
    ws = torch.empty(ws_bytes, device="cuda", dtype=torch.uint8)
    out = torch.empty(c, device="cuda")
    run(C, C.lib().qt_colsum(C.ptr(x), m, c, C.ptr(out), 0, C.ptr(ws), ws_bytes, C.stream()), "colsum wide")
    report("colsum 2688", out, x.float().sum(0), 1e-5)


@pytest.mark.parametrize("cout,cin,taps", [(64, 64, 9), (128, 64, 1), (2688, 5376, 1), (96, 40, 27)])
def test_wpack_both(C, cout, cin, taps):
    g = torch.Generator(device="cuda").manual_seed(13)
    w = torch.randn(cout, cin, taps, device="cuda", generator=g)
    wf = torch.empty(cout, taps, cin, device="cuda", dtype=torch.bfloat16)
    wd = torch.empty(cin, taps, cout, device="cuda", dtype=torch.bfloat16)
    run(C, C.lib().qt_wpack_both(C.ptr(w), C.ptr(wf), C.ptr(wd), cout, cin, taps, C.stream()), "wpack_both")
    assert torch.equal(wf, bf16(w).permute(0, 2, 1).contiguous())
    assert torch.equal(wd, bf16(w).permute(1, 2, 0).contiguous())


@pytest.mark.parametrize("with_yarg", [True, False])
def test_fused_stem_tail(C, with_yarg):
    """bn1 + relu + maxpool(3,2,1) fused forward / backward vs torch autograd on the same bf16 conv output; with `yarg` the
    backward statistics come from the pooled-size tensors (the path the models use), without it from a pass over y."""
    n, h, w, c = 3, 22, 26, 64
    g = torch.Generator(device="cuda").manual_seed(21)
    y = bf16(torch.randn(n, c, h, w, device="cuda", generator=g) * 1.5 + 0.3)
    gamma = torch.rand(c, device="cuda", generator=g) + 0.5
    beta = 0.3 * torch.randn(c, device="cuda", generator=g)
    yf = y.float().requires_grad_(True)
    gref, bref = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    z = F.batch_norm(yf, None, None, gref, bref, True, 0.1, 1e-5)
    act = z.relu()
    act = act + (bf16(act).float() - act).detach()  # the product stores bf16 activations: ties break on the rounded values
    pooled = F.max_pool2d(act, 3, 2, 1)
    dpool = bf16(torch.randn_like(pooled))
    pooled.backward(dpool.float())
    # our statistics
    y_nhwc = y.permute(0, 2, 3, 1).contiguous()
    m = n * h * w
    rows = 29
    partial = torch.zeros(rows, 2, c, device="cuda")
    run(C, C.lib().qt_bn_stats(C.ptr(y_nhwc), m, c, C.ptr(partial), rows, C.stream()), "bn_stats")
    ws_bytes = C.lib().qt_bn_workspace_bytes(c)
    ws = torch.empty(ws_bytes, device="cuda", dtype=torch.uint8)
    mean, invstd, scale, shift = (torch.empty(c, device="cuda") for _ in range(4))
    run(C, C.lib().qt_bn_finalize(C.ptr(partial), rows, c, float(m), C.ptr(gamma), C.ptr(beta), 1e-5, 0.1, None, None, C.ptr(mean),
                                  C.ptr(invstd), C.ptr(scale), C.ptr(shift), C.ptr(ws), ws_bytes, C.stream()), "bn_finalize")
    ho, wo = pooled.shape[2], pooled.shape[3]
    out = torch.empty(n, ho, wo, c, device="cuda", dtype=torch.bfloat16)
    am = torch.empty(n, ho, wo, c, device="cuda", dtype=torch.int8)
    yarg = torch.empty_like(out) if with_yarg else None
    run(C, C.lib().qt_bn_relu_maxpool_fwd(C.ptr(y_nhwc), C.ptr(scale), C.ptr(shift), C.ptr(out), C.ptr(am), C.ptr(yarg), n, h, w, c,
                                          C.stream()), "bn_relu_maxpool_fwd")
    report("fused stem tail fwd", out.permute(0, 3, 1, 2), pooled.detach(), BF16_REL_L2, True)
    dy = torch.empty_like(y_nhwc)
    dgamma, dbeta = torch.empty(c, device="cuda"), torch.empty(c, device="cuda")
    run(C, C.lib().qt_bn_relu_maxpool_bwd(C.ptr(dpool.permute(0, 2, 3, 1).contiguous()), C.ptr(am), C.ptr(y_nhwc), C.ptr(yarg), C.ptr(scale),
                                          C.ptr(shift), C.ptr(mean), C.ptr(invstd), C.ptr(gamma), n, h, w, c, C.ptr(dgamma),
                                          C.ptr(dbeta), 0, C.ptr(dy), C.ptr(ws), ws_bytes, C.stream()), "bn_relu_maxpool_bwd")
    report("fused stem tail dgamma", dgamma, gref.grad, 2e-2)
    report("fused stem tail dbeta", dbeta, bref.grad, 2e-2)
    report("fused stem tail dy", dy.permute(0, 3, 1, 2), yf.grad, 2e-2)


@pytest.mark.parametrize("n,cin,cout,h,w", [(3, 64, 64, 56, 56), (3, 128, 128, 28, 28), (5, 256, 256, 14, 14), (9, 512, 512, 7, 7),
                                            (7, 512, 512, 2, 2), (2, 64, 64, 21, 37)])
def test_conv3x3_writes_stay_in_bounds(C, n, cin, cout, h, w):
    """Guard bands around every output of the slab / generic conv kernels (fprop with statistics, data gradient with and
    without accumulation, weight gradient): nothing outside the tensor may change (compute-sanitizer is not available on
    the pool, so out-of-range stores are caught this way)."""
    g = torch.Generator(device="cuda").manual_seed(9)
    guard = 4096

    def banded(shape, dtype, fill):
        numel = math.prod(shape)
        buf = torch.full((numel + 2 * guard,), fill, device="cuda", dtype=dtype)
        return buf, buf[guard:guard + numel].view(*shape)

    def bands_intact(buf, fill):
        return bool((buf[:guard] == fill).all()) and bool((buf[-guard:] == fill).all())

    x = bf16(torch.randn(n, h, w, cin, device="cuda", generator=g))
    dy = bf16(torch.randn(n, h, w, cout, device="cuda", generator=g))
    wt = torch.randn(cout, cin, 3, 3, device="cuda", generator=g) / math.sqrt(9 * cin)
    wf = torch.empty(cout, 9, cin, device="cuda", dtype=torch.bfloat16)
    wd = torch.empty(cin, 9, cout, device="cuda", dtype=torch.bfloat16)
    run(C, C.lib().qt_wpack_both(C.ptr(wt), C.ptr(wf), C.ptr(wd), cout, cin, 9, C.stream()), "wpack_both")
    d = C.conv_desc(n, (1, h, w), cin, cout, (1, 3, 3), (1, 1, 1), (0, 1, 1))
    ybuf, y = banded((n, h, w, cout), torch.bfloat16, 7.0)
    rows = C.lib().qt_conv_stat_rows(d)
    sbuf, st = banded((rows, 2, cout), torch.float32, 7.0)
    run(C, C.lib().qt_conv_fprop(d, C.ptr(x), C.ptr(wf), C.ptr(y), None, C.ptr(st), C.QT_EPI_STATS, None, 0, C.stream()), "fprop")
    assert bands_intact(ybuf, 7.0) and bands_intact(sbuf, 7.0)
    assert not bool((y == 7.0).all()) and bool(torch.isfinite(y.float()).all())
    dxbuf, dx = banded((n, h, w, cin), torch.bfloat16, 7.0)
    for acc in (0, 1):
        run(C, C.lib().qt_conv_dgrad(d, C.ptr(dy), C.ptr(wd), C.ptr(dx), acc, C.stream()), "dgrad")
        assert bands_intact(dxbuf, 7.0)
    ws_bytes = C.lib().qt_conv_wgrad_workspace_bytes(d)
    wsbuf, ws = banded((max(ws_bytes, 16),), torch.uint8, 7)
    dwbuf, dw = banded((cout, cin, 3, 3), torch.float32, 7.0)
    run(C, C.lib().qt_conv_wgrad(d, C.ptr(x), C.ptr(dy), C.ptr(dw), 0, C.ptr(ws), ws_bytes, C.stream()), "wgrad")
    assert bands_intact(dwbuf, 7.0) and bands_intact(wsbuf, 7)
    ref = torch.nn.grad.conv2d_weight(x.float().permute(0, 3, 1, 2), (cout, cin, 3, 3), dy.float().permute(0, 3, 1, 2), padding=1)
    report(f"wgrad (banded) c{cin}->{cout} {h}x{w}", dw, ref, F32_REL_L2)
    assert C.lib().qt_take_timeout_flag() == 0


def test_wpack_multi_matches_single_packs(C):
    """One qt_wpack_multi launch over a mixed table (3x3, 1x1 / linear, 3x3x3, ragged channel counts, one item without the
    data-gradient layout) produces exactly the bf16 layouts of the per-weight reference transforms."""
    import ctypes
    g = torch.Generator(device="cuda").manual_seed(31)
    shapes = [(64, 64, 9), (128, 64, 1), (72, 40, 9), (256, 128, 9), (200, 136, 1), (32, 8, 27), (2688, 544, 1)]
    items = (C.WpackItem * len(shapes))()
    keep, first, max_taps = [], 0, 1
    for i, (cout, cin, taps) in enumerate(shapes):
        w = torch.randn(cout, cin, taps, device="cuda", generator=g)
        wf = torch.zeros(cout, taps, cin, device="cuda", dtype=torch.bfloat16)
        wd = torch.zeros(cin, taps, cout, device="cuda", dtype=torch.bfloat16) if i != 4 else None
        keep.append((w, wf, wd))
        it = items[i]
        it.w, it.wf, it.wd = w.data_ptr(), wf.data_ptr(), (wd.data_ptr() if wd is not None else None)
        it.cout, it.cin, it.taps = cout, cin, taps
        nb = C.lib().qt_wpack_item_plan(ctypes.byref(it))
        assert nb > 0
        it.first_block = first
        first += nb
        max_taps = max(max_taps, taps)
    table = torch.frombuffer(bytearray(bytes(items)), dtype=torch.uint8).cuda()
    run(C, C.lib().qt_wpack_multi(C.ptr(table), len(shapes), first, max_taps, C.stream()), "wpack_multi")
    for w, wf, wd in keep:
        assert torch.equal(wf, bf16(w).permute(0, 2, 1).contiguous())
        if wd is not None:
            assert torch.equal(wd, bf16(w).permute(1, 2, 0).contiguous())
    bad = C.WpackItem()
    bad.cout, bad.cin, bad.taps = 8, 8, 99
    assert C.lib().qt_wpack_item_plan(ctypes.byref(bad)) < 0
