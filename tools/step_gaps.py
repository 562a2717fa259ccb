"""Idle gaps between consecutive kernels of one bench workload step (torch.profiler / CUPTI timestamps; development aid).
python tools/step_gaps.py [--workload quadtree_train]"""
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
    ap.add_argument("--steps", type=int, default=4)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    w = bench.make_workload(a.workload, None, dev)
    batch = tuple(t.to(dev) for t in w["host_fp32"])
    for _ in range(5):
        w["step"](*batch)
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(a.steps):
            w["step"](*batch)
        torch.cuda.synchronize()
    ev = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start),
                key=lambda e: e.time_range.start)
    gaps = []
    busy = 0.0
    for p, n in zip(ev, ev[1:]):
        busy += p.time_range.end - p.time_range.start
        gaps.append((n.time_range.start - p.time_range.end, p.name[:60], n.name[:60]))
    span = ev[-1].time_range.end - ev[0].time_range.start
    tot = sum(g[0] for g in gaps if g[0] > 0)
    print(f"{len(ev)} kernels over {span/1e3:.3f} ms ({span/1e3/a.steps:.3f} per step); idle {tot/1e3/a.steps:.3f} ms per step")
    hist = {}
    for g, _, _ in gaps:
        b = "<1" if g < 1 else "1-2" if g < 2 else "2-4" if g < 4 else "4-10" if g < 10 else "10-50" if g < 50 else ">50"
        hist[b] = hist.get(b, [0, 0.0])
        hist[b][0] += 1
        hist[b][1] += max(g, 0)
    for k in ("<1", "1-2", "2-4", "4-10", "10-50", ">50"):
        if k in hist:
            print(f"  gaps {k:6s} us: {hist[k][0] / a.steps:7.1f} per step, {hist[k][1] / a.steps:8.1f} us per step")
    for g, p, n in sorted(gaps, reverse=True)[:16]:
        print(f"  {g:8.1f} us after {p}  ->  {n}")


if __name__ == "__main__":
    main()
