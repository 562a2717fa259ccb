"""World-size-2 gloo test (CPU) of the data-parallel gradient bucketing used for N > 1 GPUs:
bucket plan learned from the first backward, gradients produced inside flat bucket views, async all-reduce
per bucket, unused parameters ignored, result == average of the per-rank gradients."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from qtcnn_b200 import ops, parallel
        torch.manual_seed(100 + rank)  # different initial weights: broadcast must equalise them
        model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))
        unused = torch.nn.Linear(4, 4)  # like base_cnn.fc: registered, never used
        model.add_module("unused", unused)
        dp = parallel.DataParallelGrads(model, bucket_bytes=64)  # tiny buckets -> several of them
        w0 = model[0].weight.detach().clone()
        results = []
        for step in range(3):
            g = torch.Generator().manual_seed(10 * step + rank)
            x = torch.randn(4, 6, generator=g)
            for p in model.parameters():
                p.grad = None
            used = [p for n, p in model.named_parameters() if not n.startswith("unused")]
            # local gradients without touching .grad (autograd.grad does not fire the accumulate hooks)
            local = [g.clone() for g in torch.autograd.grad(model[2](model[1](model[0](x))).square().mean(), used)]
            out = model[2](model[1](model[0](x)))
            out.square().mean().backward()
            dp.finish()
            synced = [p.grad.detach().clone() for p in model.parameters() if p.grad is not None]
            # reference: explicit all-reduce of the local gradients
            ok = True
            for l, s in zip(local, synced):
                ref = l.clone()
                dist.all_reduce(ref)
                ref /= world
                ok = ok and torch.allclose(ref, s, atol=1e-6)
            in_views = all(p.grad.data_ptr() == dp.bucket_of[p].views[p].data_ptr() for p in dp.bucket_of) if step > 0 else True
            results.append((ok, in_views, len(dp.buckets), dp.collectives_last_step))
        view_ptr = dp.bucket_of[model[0].weight].views[model[0].weight].data_ptr()
        # (advisor finding, round 1) the view is handed out only while p.grad is None; with a live gradient a fresh tensor
        # is returned so that autograd's accumulation adds into the view instead of a kernel overwriting it
        live_is_fresh = ops.grad_out(model[0].weight).data_ptr() != view_ptr
        model[0].weight.grad = None
        none_is_view = ops.grad_out(model[0].weight).data_ptr() == view_ptr
        # zero_grad(set_to_none=False): gradients stay allocated (the views) and must not double
        x = torch.randn(4, 6, generator=torch.Generator().manual_seed(77 + rank))
        used = [p for n, p in model.named_parameters() if not n.startswith("unused")]
        for p in model.parameters():
            p.grad = None
        model[2](model[1](model[0](x))).square().mean().backward()
        dp.finish()
        first = [p.grad.detach().clone() for p in used]
        for p in used:
            p.grad.zero_()
        model[2](model[1](model[0](x))).square().mean().backward()
        dp.finish()
        keep_ok = all(torch.allclose(a, p.grad, atol=1e-7) for a, p in zip(first, used))
        # micro-batch accumulation: no_sync() for all but the last backward
        for p in used:
            p.grad.zero_()
        with dp.no_sync():
            (0.5 * model[2](model[1](model[0](x))).square().mean()).backward()
        (0.5 * model[2](model[1](model[0](x))).square().mean()).backward()
        dp.finish()
        accum_ok = all(torch.allclose(a, p.grad, atol=1e-6) for a, p in zip(first, used))
        # a second synchronising backward before finish() must raise instead of silently corrupting the buckets
        for p in used:
            p.grad.zero_()
        model[2](model[1](model[0](x))).square().mean().backward()
        try:
            model[2](model[1](model[0](x))).square().mean().backward()
            raised = False
        except RuntimeError:
            raised = True
        for b in dp.buckets:  # drain the collectives the first backward issued
            if b.work is not None:
                b.work.wait()
        q.put((rank, results, w0.tolist(), unused.weight.grad is None, live_is_fresh and none_is_view, keep_ok, accum_ok, raised))
    finally:
        dist.destroy_process_group()


def test_bucketed_allreduce_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    outs.sort()
    assert outs[0][2] == outs[1][2], "parameters must be broadcast from rank 0"
    for rank, results, _, unused_none, view_ok, keep_ok, accum_ok, raised in outs:
        assert unused_none and view_ok
        assert keep_ok, "zero_grad(set_to_none=False) changed the synchronised gradient"
        assert accum_ok, "no_sync() accumulation differs from the single-backward gradient"
        assert raised, "a second synchronising backward before finish() must raise"
        for ok, in_views, nb, ncoll in results:
            assert ok and in_views
        assert results[-1][2] >= 2 and results[-1][3] == results[-1][2]  # several buckets, one collective each
