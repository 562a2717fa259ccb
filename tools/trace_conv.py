"""Where do the roles of the persistent conv3x3 kernel wait? Needs the developer build of the library:
    nvcc ... -DQT_TRACE -o build/libqtcnn_trace.so csrc/qtcnn.cu ; QTCNN_LIB=build/libqtcnn_trace.so python tools/trace_conv.py
Prints, per layer and pass, the mean over CTAs of the cycles each role spent in its barrier waits and in total."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qtcnn_b200.capi as C  # noqa: E402

LAYERS = [("layer1", 64, 64, 56), ("layer2", 128, 128, 28), ("layer3", 256, 256, 14), ("layer4", 512, 512, 7)]
NAMES = ["mma:acc_empty", "mma:a_full", "mma:b_full", "mma:total", "epi:acc_full", "epi:total", "prod:a_empty", "prod:total"]


def main():
    lib = C.lib()
    rd = lib.qt_debug_read_trace
    rd.restype = ctypes.c_int
    rd.argtypes = [ctypes.c_void_p, ctypes.c_int]
    n = 256
    st = C.stream()
    buf = (ctypes.c_longlong * (148 * 16))()
    for name, cin, cout, h in LAYERS:
        d = C.conv_desc(n, (1, h, h), cin, cout, (1, 3, 3), (1, 1, 1), (0, 1, 1))
        x = torch.randn(n, h, h, cin, device="cuda").to(torch.bfloat16)
        y = torch.empty(n, h, h, cout, device="cuda", dtype=torch.bfloat16)
        dx = torch.zeros_like(x)
        wf = (torch.randn(cout, 9, cin, device="cuda") * 0.05).to(torch.bfloat16)
        stats = torch.empty(lib.qt_conv_stat_rows(d), 2, cout, device="cuda")
        for which in ("fprop", "dgrad"):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for it in range(13):
                if it == 3:
                    e0.record()
                if which == "fprop":
                    C.check(lib.qt_conv_fprop(d, C.ptr(x), C.ptr(wf), C.ptr(y), None, C.ptr(stats), C.QT_EPI_STATS, None, 0, st))
                else:
                    C.check(lib.qt_conv_dgrad(d, C.ptr(y), C.ptr(wf), C.ptr(dx), 1, st))
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            rd(buf, 148 * 16)
            vals = torch.tensor(list(buf), dtype=torch.float64).view(148, 16)
            mean = vals.mean(0)
            print(f"{name} {which}: " + "  ".join(f"{k}={mean[i]/1e3:.1f}k" for i, k in enumerate(NAMES)), flush=True)
            t0 = vals[:, 8].min()
            print(f"    wall (us, relative to the first CTA entry): entry max {float(vals[:, 8].max() - t0)/1e3:.1f}  prologue done mean {float(vals[:, 9].mean() - t0)/1e3:.1f} "
                  f"max {float(vals[:, 9].max() - t0)/1e3:.1f}  mma done mean {float(vals[:, 10].mean() - t0)/1e3:.1f} max {float(vals[:, 10].max() - t0)/1e3:.1f}  "
                  f"epilogue done max {float(vals[:, 11].max() - t0)/1e3:.1f}  exit max {float(vals[:, 12].max() - t0)/1e3:.1f}; event time/launch {ms*1e3:.1f} us", flush=True)


if __name__ == "__main__":
    main()
