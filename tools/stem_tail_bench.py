"""Times the fused vs unfused stem tail (bn1+relu+maxpool) at the real shape (B200 only)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qtcnn_b200.capi as C
lib = C.lib(); st = C.stream()
n, h, w, c = 256, 112, 112, 64
y = torch.randn(n, h, w, c, device="cuda").to(torch.bfloat16)
a = torch.empty_like(y); out = torch.empty(n, 56, 56, c, device="cuda", dtype=torch.bfloat16)
am = torch.empty(n, 56, 56, c, device="cuda", dtype=torch.int8)
yarg = torch.empty_like(out)
scale = torch.rand(c, device="cuda") + 0.5; shift = torch.randn(c, device="cuda") * 0.1
mean = torch.zeros(c, device="cuda"); invstd = torch.ones(c, device="cuda"); gamma = torch.ones(c, device="cuda")
dpool = torch.randn(n, 56, 56, c, device="cuda").to(torch.bfloat16)
da = torch.empty_like(y); dy = torch.empty_like(y)
dg = torch.empty(c, device="cuda"); db = torch.empty(c, device="cuda")
wsb = lib.qt_bn_workspace_bytes(c); ws = torch.empty(wsb, device="cuda", dtype=torch.uint8)
def t(fn, it=10):
    for _ in range(2): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / it * 1e3
def unf_fwd():
    C.check(lib.qt_bn_apply(C.ptr(y), C.ptr(scale), C.ptr(shift), None, C.ptr(a), n*h*w, c, 1, st))
    C.check(lib.qt_maxpool2d_fwd(C.ptr(a), C.ptr(out), C.ptr(am), n, h, w, c, 3, 2, 1, st))
def fus_fwd():
    C.check(lib.qt_bn_relu_maxpool_fwd(C.ptr(y), C.ptr(scale), C.ptr(shift), C.ptr(out), C.ptr(am), C.ptr(yarg), n, h, w, c, st))
def unf_bwd():
    C.check(lib.qt_maxpool2d_bwd(C.ptr(dpool), C.ptr(am), C.ptr(da), n, h, w, c, 3, 2, 1, st))
    C.check(lib.qt_bn_backward(C.ptr(da), C.ptr(a), C.ptr(y), C.ptr(mean), C.ptr(invstd), C.ptr(gamma), None, None, n*h*w, c, C.ptr(dg), C.ptr(db), 0, 0, C.ptr(dy), None, C.ptr(ws), wsb, st))
def fus_bwd():
    C.check(lib.qt_bn_relu_maxpool_bwd(C.ptr(dpool), C.ptr(am), C.ptr(y), C.ptr(yarg), C.ptr(scale), C.ptr(shift), C.ptr(mean), C.ptr(invstd), C.ptr(gamma), n, h, w, c, C.ptr(dg), C.ptr(db), 0, C.ptr(dy), C.ptr(ws), wsb, st))
print(f"fwd unfused {t(unf_fwd):.0f} us  fused {t(fus_fwd):.0f} us | bwd unfused {t(unf_bwd):.0f} us  fused {t(fus_bwd):.0f} us")
