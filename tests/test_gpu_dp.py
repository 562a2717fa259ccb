"""Two-GPU check of the data-parallel path (skipped with fewer than 2 devices): after one synchronised step
both ranks hold identical gradients equal to the average of the per-rank gradients, produced inside the flat
bucket views."""
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


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import torch.nn.functional as F
        from oracle import quadtree_oracle as O
        from qtcnn_b200 import models as M
        from qtcnn_b200 import parallel
        dev = torch.device("cuda", rank)
        model = M.QuadtreeCNN(num_classes=8, dropout_rate=0.0)
        M.load_oracle_params(model, O.make_params("quadtree", 8, seed=rank))  # different weights: broadcast must fix it
        model = model.to(dev).train()
        dp = parallel.DataParallelGrads(model)
        out = []
        for step in range(3):
            images, numerical, labels = O.synthetic_batch(4, 100 + 10 * step + rank)
            for p in model.parameters():
                p.grad = None
            loss = F.cross_entropy(model(images.to(dev), numerical.to(dev)), labels.to(dev))
            loss.backward()
            dp.finish()
            g = torch.cat([p.grad.flatten() for p in model.parameters() if p.grad is not None])
            gsum = g.clone()
            dist.all_reduce(gsum)
            same = bool(torch.allclose(gsum / world, g, rtol=0, atol=0))  # already averaged => identical on all ranks
            in_views = all(p.grad.data_ptr() == dp.bucket_of[p].views[p].data_ptr() for p in dp.bucket_of) if step > 0 else True
            out.append((same, in_views, len(dp.buckets), float(g.norm())))
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


def test_two_gpu_gradient_sync():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=600) for _ in range(2))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for (rank, out) in res:
        for same, in_views, nb, norm in out:
            assert same and in_views and norm > 0
        assert out[-1][2] >= 3
    assert [o[3] for o in res[0][1]] == [o[3] for o in res[1][1]], "ranks must hold identical gradients"
