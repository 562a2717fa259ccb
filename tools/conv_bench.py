"""Micro-benchmark of the conv GEMM kernels on the ResNet-18 / QuadtreeCNN layer shapes (B200 only).
python tools/conv_bench.py [--batch 256] [--which fprop,dgrad,wgrad] [--v1]  -> TFLOP/s per layer and pass."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qtcnn_b200.capi as C  # noqa: E402

LAYERS = [  # name, cin, cout, h, w, k, stride, pad
    ("layer1 3x3", 64, 64, 56, 56, 3, 1, 1),
    ("layer2.0.c1 s2", 64, 128, 56, 56, 3, 2, 1),
    ("layer2 3x3", 128, 128, 28, 28, 3, 1, 1),
    ("layer2 ds 1x1", 64, 128, 56, 56, 1, 2, 0),
    ("layer3.0.c1 s2", 128, 256, 28, 28, 3, 2, 1),
    ("layer3 3x3", 256, 256, 14, 14, 3, 1, 1),
    ("layer4.0.c1 s2", 256, 512, 14, 14, 3, 2, 1),
    ("layer4 3x3", 512, 512, 7, 7, 3, 1, 1),
]


ITERS = 20


def timeit(fn, iters=None):
    iters = iters or ITERS
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--which", default="fprop,dgrad,wgrad")
    ap.add_argument("--v1", action="store_true", help="disable the persistent 3x3 kernel")
    ap.add_argument("--layers", default="", help="substring filter on layer names")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--tune", default="", help="key=value,... passed to qt_set_tuning")
    a = ap.parse_args()
    global ITERS
    ITERS = a.iters
    lib = C.lib()
    lib.qt_set_conv3x3_enabled(0 if a.v1 else 1)
    for kv in filter(None, a.tune.split(",")):
        k, v = kv.split("=")
        lib.qt_set_tuning(int(k), int(v))
    n = a.batch
    st = C.stream()
    for name, cin, cout, h, w, k, s, p in LAYERS:
        if a.layers and a.layers not in name:
            continue
        d = C.conv_desc(n, (1, h, w), cin, cout, (1, k, k), (1, s, s), (0, p, p))
        ho, wo = C.out_size(h, k, s, p), C.out_size(w, k, s, p)
        x = torch.randn(n, h, w, cin, device="cuda").to(torch.bfloat16)
        y = torch.empty(n, ho, wo, cout, device="cuda", dtype=torch.bfloat16)
        dy = torch.randn(n, ho, wo, cout, device="cuda").to(torch.bfloat16)
        dx = torch.zeros_like(x)
        wf = (torch.randn(cout, k * k, cin, device="cuda") * 0.05).to(torch.bfloat16)
        wd = (torch.randn(cin, k * k, cout, device="cuda") * 0.05).to(torch.bfloat16)
        dw = torch.empty(cout, cin, k, k, device="cuda")
        rows = lib.qt_conv_stat_rows(d)
        stats = torch.empty(rows, 2, cout, device="cuda")
        wsb = lib.qt_conv_wgrad_workspace_bytes(d)
        ws = torch.empty(max(wsb, 16), device="cuda", dtype=torch.uint8)
        flops = 2.0 * n * ho * wo * k * k * cin * cout
        out = [f"{name:16s}"]
        if "fprop" in a.which:
            ms = timeit(lambda: C.check(lib.qt_conv_fprop(d, C.ptr(x), C.ptr(wf), C.ptr(y), None, C.ptr(stats), C.QT_EPI_STATS, None, 0, st)))
            out.append(f"fprop {ms*1e3:7.1f} us {flops/ms/1e9:7.1f} TF")
        if "dgrad" in a.which:
            acc = 1 if (s > 1 and k == 1) else 0
            ms = timeit(lambda: C.check(lib.qt_conv_dgrad(d, C.ptr(dy), C.ptr(wd), C.ptr(dx), acc, st)))
            out.append(f"dgrad {ms*1e3:7.1f} us {flops/ms/1e9:7.1f} TF")
        if "wgrad" in a.which:
            ms = timeit(lambda: C.check(lib.qt_conv_wgrad(d, C.ptr(x), C.ptr(dy), C.ptr(dw), 0, C.ptr(ws), wsb, st)))
            out.append(f"wgrad {ms*1e3:7.1f} us {flops/ms/1e9:7.1f} TF")
        print(" | ".join(out), flush=True)
    assert lib.qt_take_timeout_flag() == 0, "barrier timeout inside a kernel"


if __name__ == "__main__":
    main()
