import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
dev = torch.device("cuda", 0)
w = bench.make_workload(os.environ.get("WORKLOAD", "quadtree3d_train"), None, dev)
batch = tuple(t.to(dev) for t in w["host_fp32"])
def stats():
    s = torch.cuda.memory_stats()
    return s["num_device_alloc"], s["num_device_free"], s["reserved_bytes.all.current"] >> 20, s["allocated_bytes.all.peak"] >> 20
for i in range(14):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    w["step"](*batch)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"step {i}: host {(t1-t0)*1e3:.2f} ms, total {(t2-t0)*1e3:.2f} ms, (cudaMalloc, cudaFree, reserved MB, peak MB) = {stats()}", flush=True)
