"""Stand-alone timing of the quadtree-stage kernels at the headline shape (B = 256): 24 rotating buffer sets (> L2 in total),
CUDA events around the whole batch of launches -> microseconds per launch and achieved algorithmic GB/s."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qtcnn_b200.capi as C  # noqa: E402


def main():
    lib = C.lib()
    b, ldf, sets = 256, 5376, 24
    bf = torch.bfloat16
    q = [torch.randn(4, b, 7, 7, 128, device="cuda").to(bf).relu() for _ in range(sets)]
    l4 = [torch.randn(b, 7, 7, 512, device="cuda").to(bf).relu() for _ in range(sets)]
    feat = [torch.zeros(b, ldf, device="cuda", dtype=bf) for _ in range(sets)]
    dfeat = [torch.randn(b, ldf, device="cuda").to(bf) for _ in range(sets)]
    dq = [torch.empty_like(q[0]) for _ in range(sets)]
    dl = [torch.empty_like(l4[0]) for _ in range(sets)]
    st = C.stream()

    def fwd(i):
        C.check(lib.qt_quadtree_pool_fwd(C.ptr(q[i]), C.ptr(l4[i]), C.ptr(feat[i]), b, 7, 7, 128, 49, 512, ldf, st))

    def bwd(i):
        C.check(lib.qt_quadtree_pool_bwd(C.ptr(dfeat[i]), C.ptr(q[i]), C.ptr(dq[i]), C.ptr(dl[i]), b, 7, 7, 128, 49, 512, ldf, st))

    for name, fn, nbytes in (("quadtree_pool_fwd", fwd, 2.0 * (q[0].numel() + l4[0].numel() + b * 5120)),
                             ("quadtree_pool_bwd", bwd, 2.0 * (b * 5120 + q[0].numel() + dq[0].numel() + dl[0].numel()))):
        for i in range(sets):
            fn(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record()
        for _ in range(reps):
            for i in range(sets):
                fn(i)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (reps * sets)
        print(f"{name}: {us:.2f} us per launch, {nbytes/1e6:.1f} MB algorithmic -> {nbytes/us/1e3:.0f} GB/s", flush=True)


if __name__ == "__main__":
    main()
