"""Data-parallel gradient synchronisation on real devices, two ranks. With >= 2 GPUs: one rank per GPU over NCCL.
On a one-GPU box (the driver's test box) both ranks share cuda:0 and talk through gloo on CUDA tensors — NCCL refuses
two ranks on one device, the bucket / hook / view logic under test is the same.

Checked: parameters are broadcast from rank 0 (packed bf16 weight copies made before wrapping are dropped), after every
synchronised step both ranks hold bit-identical gradients equal to the mean of the per-rank local gradients, gradients
live inside the flat bucket views from step 2 on, `zero_grad(set_to_none=False)` and `no_sync()` micro-batch
accumulation give the same result as the plain path."""
import copy
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q, ndev):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["QTCNN_QUIET_PRETRAINED"] = "1"
    dev = torch.device("cuda", rank % ndev)
    torch.cuda.set_device(dev)
    if ndev >= world:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import torch.nn.functional as F
        from oracle import quadtree_oracle as O
        from oracle.loading import load_oracle_params
        from qtcnn_b200 import models as M
        from qtcnn_b200 import parallel

        def local_grads(m, batch):
            images, numerical, labels = batch
            for p in m.parameters():
                p.grad = None
            F.cross_entropy(m(images.to(dev), numerical.to(dev)), labels.to(dev)).backward()
            return {n: p.grad.detach().clone() for n, p in m.named_parameters() if p.grad is not None}

        model = M.QuadtreeCNN(num_classes=8, dropout_rate=0.0)
        load_oracle_params(model, O.make_params("quadtree", 8, seed=rank))  # different weights: broadcast must fix it
        model = model.to(dev).train()
        with torch.no_grad():  # a forward BEFORE wrapping packs bf16 weight copies from the rank's own (soon stale) weights
            model(*[t.to(dev) for t in O.synthetic_batch(2, 7)[:2]])
        dp = parallel.DataParallelGrads(model)
        shadow = copy.deepcopy(model)  # un-wrapped twin (same broadcast weights) for the local gradients
        out = []
        for step in range(4):
            batch = O.synthetic_batch(4, 100 + 10 * step + rank)
            want = local_grads(shadow, batch)
            for g in want.values():
                dist.all_reduce(g)
                g.div_(world)
            images, numerical, labels = batch
            if step == 2:
                model.zero_grad(set_to_none=False)  # gradients stay allocated (the bucket views): must not double
            else:
                for p in model.parameters():
                    p.grad = None
            if step == 3:
                # two micro-batches: the first accumulates without communication, the second triggers the all-reduce
                half = [t[:2] for t in batch], [t[2:] for t in batch]
                with dp.no_sync():
                    (0.5 * F.cross_entropy(model(half[0][0].to(dev), half[0][1].to(dev)), half[0][2].to(dev))).backward()
                (0.5 * F.cross_entropy(model(half[1][0].to(dev), half[1][1].to(dev)), half[1][2].to(dev))).backward()
                dp.finish()
                got = {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}
                # BatchNorm statistics differ between a batch of 4 and two batches of 2, so compare with the same
                # accumulation done locally
                for p in shadow.parameters():
                    p.grad = None
                for hb in half:
                    (0.5 * F.cross_entropy(shadow(hb[0].to(dev), hb[1].to(dev)), hb[2].to(dev))).backward()
                want = {n: p.grad.detach().clone() for n, p in shadow.named_parameters() if p.grad is not None}
                for g in want.values():
                    dist.all_reduce(g)
                    g.div_(world)
            else:
                F.cross_entropy(model(images.to(dev), numerical.to(dev)), labels.to(dev)).backward()
                dp.finish()
                got = {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}
            same_keys = got.keys() == want.keys()
            worst = max(float((got[n] - want[n]).abs().max() / (want[n].abs().max() + 1e-30)) for n in want)
            in_views = all(p.grad.data_ptr() == dp.bucket_of[p].views[p].data_ptr() for p in dp.bucket_of) if step > 0 else True
            digest = float(torch.cat([g.flatten() for g in got.values()]).double().sum())
            out.append((same_keys, worst, in_views, len(dp.buckets), digest))
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


def test_two_rank_gradient_sync():
    ndev = torch.cuda.device_count()
    if ndev < 1:
        pytest.skip("needs a GPU")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, ndev)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=900) for _ in range(2))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for (rank, out) in res:
        for step, (same_keys, worst, in_views, nb, digest) in enumerate(out):
            assert same_keys and in_views, (rank, step)
            # step 3 sums two micro-batch gradients in a different order than the twin (fresh tensor + in-place add)
            assert worst <= (1e-5 if step == 3 else 0.0), (rank, step, worst)
        assert out[-1][3] >= 3
    assert [o[4] for o in res[0][1]] == [o[4] for o in res[1][1]], "ranks must hold bit-identical gradients"
