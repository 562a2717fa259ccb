import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qtcnn_b200.capi as C
lib = C.lib()
n, D, H, W = 32, 16, 112, 112
d = C.conv_desc(n, (D, H, W), 8, 32, (3, 3, 3), (1, 1, 1), (1, 1, 1))
x = torch.randn(n, D, H, W, 8, device="cuda").to(torch.bfloat16)
dy = torch.randn(n, D, H, W, 32, device="cuda").to(torch.bfloat16)
dw = torch.empty(32, 8, 27, device="cuda")
wsb = lib.qt_conv_wgrad_workspace_bytes(d)
ws = torch.empty(wsb, device="cuda", dtype=torch.uint8)
for mode in (0, 1, 2):
    lib.qt_set_tuning(9, mode)
    for _ in range(3):
        C.check(lib.qt_conv_wgrad(d, C.ptr(x), C.ptr(dy), C.ptr(dw), 0, C.ptr(ws), wsb, C.stream()))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        C.check(lib.qt_conv_wgrad(d, C.ptr(x), C.ptr(dy), C.ptr(dw), 0, C.ptr(ws), wsb, C.stream()))
    e1.record(); torch.cuda.synchronize()
    print("debug mode", mode, "us per launch", e0.elapsed_time(e1) * 100)
