"""How long does NCCL take for the gradient exchange of one QuadtreeCNN step when nothing else runs? (torchrun, one rank per GPU)
Buckets as parallel.DataParallelGrads builds them: 58 / 24 / 21 MB fp32, SUM and AVG, plus one flat 104 MB call and bf16."""
import os

import torch
import torch.distributed as dist


def main():
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sizes = [14_450_000, 6_100_000, 5_400_000]
    bufs = [torch.randn(n, device="cuda") for n in sizes]
    flat = torch.randn(sum(sizes), device="cuda")
    half = flat.to(torch.bfloat16)

    def timed(fn, iters=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    def buckets(op):
        works = [dist.all_reduce(b, op=op, async_op=True) for b in bufs]
        for w in works:
            w.wait()

    res = {
        "3 buckets SUM": timed(lambda: buckets(dist.ReduceOp.SUM)),
        "3 buckets AVG": timed(lambda: buckets(dist.ReduceOp.AVG)),
        "flat 104 MB SUM": timed(lambda: dist.all_reduce(flat, op=dist.ReduceOp.SUM)),
        "flat 52 MB bf16 SUM": timed(lambda: dist.all_reduce(half, op=dist.ReduceOp.SUM)),
    }
    if rank == 0:
        n = dist.get_world_size()
        for k, ms in res.items():
            nbytes = half.numel() * 2 if "bf16" in k else flat.numel() * 4
            print(f"N={n} {k:22s} {ms:7.3f} ms  algbw {nbytes/ms/1e6:7.1f} GB/s  busbw {nbytes/ms/1e6*2*(n-1)/n:7.1f} GB/s", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
