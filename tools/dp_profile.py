"""Rank-0 view of one data-parallel step (torchrun): host issue time per step, device time per step, per-kernel table.
Development aid for profiles/r02_scaling.md."""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    w = bench.make_workload(os.environ.get("WORKLOAD", "quadtree_train"), None, dev, world, rank)
    batch = tuple(t.to(dev) for t in w["host_fp32"])
    step = lambda: w["step"](*batch)  # noqa: E731
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    n = 10
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        step()
    host = (time.perf_counter() - t0) / n * 1e3
    e1.record()
    torch.cuda.synchronize()
    devms = e0.elapsed_time(e1) / n
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            step()
        torch.cuda.synchronize()
    if rank == 0:
        rows = []
        for e in prof.key_averages():
            t = getattr(e, "device_time_total", None)
            if t is None:
                t = e.cuda_time_total
            if t > 0:
                rows.append((t / 3, e.count / 3, e.key))
        rows.sort(reverse=True)
        total = sum(r[0] for r in rows)
        print(f"world {world}: host issue {host:.2f} ms/step, device {devms:.2f} ms/step, kernel sum {total/1e3:.2f} ms/step, "
              f"{sum(r[1] for r in rows):.0f} launches")
        for t, c, k in rows[:int(os.environ.get("TOP", "12"))]:
            print(f"{t:10.1f} us x{c:5.1f}  {k[:100]}")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
