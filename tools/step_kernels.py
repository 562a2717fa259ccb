"""Per-kernel time table of one training / inference step of a bench workload (torch.profiler / CUPTI; development aid —
numbers under a profiler are never bench values).  python tools/step_kernels.py --workload quadtree3d_train [--batch 32]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("QTCNN_QUIET_PRETRAINED", "1")
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="quadtree_train")
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--top", type=int, default=40)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    w = bench.make_workload(a.workload, a.batch or None, dev)
    batch = tuple(t.to(dev) for t in w["host_fp32"])

    def step():
        return w["step"](*batch)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(a.steps):
            step()
        torch.cuda.synchronize()
    rows = []
    for e in prof.key_averages():
        t = getattr(e, "device_time_total", None)
        if t is None:
            t = e.cuda_time_total
        if t > 0:
            rows.append((t / a.steps, e.count / a.steps, e.key))
    rows.sort(reverse=True)
    total = sum(r[0] for r in rows)
    print(f"{a.workload}: {total/1e3:.3f} ms of kernels per step, {sum(r[1] for r in rows):.0f} launches per step")
    for t, n, k in rows[:a.top]:
        print(f"{t:10.1f} us {100*t/total:5.1f}%  x{n:6.1f}  {k[:110]}")


if __name__ == "__main__":
    main()
