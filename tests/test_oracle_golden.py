"""Pins oracle/quadtree_oracle.py to the reference: replays the fixtures that tests/golden/make_golden.py
produced by running the unmodified reference modules (CPU, fp32) on seeded parameters and inputs.

Tolerance: fp32 CPU vs fp32 CPU of the same operators in a different composition — logits/loss 1e-4 relative
(+1e-5 abs), gradient norms 2e-3 relative, sampled gradient elements 5e-3 relative + 3 % of the tensor rms
(B=2 train-mode BatchNorm amplifies fp32 accumulation-order noise in the early layers) (MKL-DNN conv algorithms differ between nn.Module and
functional call paths only by accumulation order).
"""
import glob
import json
import os

import pytest
import torch

from oracle import quadtree_oracle as O

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.json")))


def close(a, b, rel, abs_=1e-6):
    return abs(a - b) <= abs_ + rel * max(abs(a), abs(b))


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-5] for p in GOLD])
def test_oracle_matches_reference_fixture(path):
    torch.set_num_threads(max(1, min(8, os.cpu_count() or 1)))
    with open(path) as f:
        gold = json.load(f)
    case = gold["case"]
    kind, mode = case["kind"], case.get("mode", "fusion")
    training = case.get("training", True)
    p = O.make_params(kind, 8, seed=case["param_seed"], mode=mode)
    if kind in ("quadtree3d", "cnn_lstm", "resnet3d_video", "hybrid3d"):
        images, numerical, labels = O.synthetic_batch(case["batch"], case["seed"], seq_len=case["seq_len"],
                                                      clip_size=case["clip"])
    else:
        images, numerical, labels = O.synthetic_batch(case["batch"], case["seed"])
    kw = {"mode": mode} if kind in ("quadtree", "quadtree3d", "hybrid3d") else {}
    ref_logits = torch.tensor(gold["logits"], dtype=torch.float64)
    if not training:
        with torch.no_grad():
            logits = O.FORWARDS[kind](p, images, numerical, training=False, **kw)
        assert torch.allclose(logits.double(), ref_logits, rtol=1e-4, atol=1e-5)
        return
    logits, loss, grads, nb = O.loss_and_grads(kind, p, (images, numerical), labels, training=True, **kw)
    assert torch.allclose(logits.double(), ref_logits, rtol=1e-4, atol=1e-5), (logits, ref_logits)
    assert close(float(loss), gold["loss"], 1e-4)
    assert gold["grads"], "fixture without gradients"
    for name, d in gold["grads"].items():
        assert name in grads, f"oracle produced no gradient for {name}"
        g = grads[name].double().flatten()
        assert close(float(g.norm()), d["norm"], 2e-3), (name, float(g.norm()), d["norm"])
        for i, v in zip(d["idx"], d["val"]):
            assert close(float(g[i]), v, 5e-3, 1e-6 + 3e-2 * d["norm"] / max(1.0, g.numel() ** 0.5)), (name, i, float(g[i]), v)
    # parameters the reference never reaches (base_cnn.fc) must not get a gradient in the oracle either
    if kind == "quadtree" and "file" not in case:
        assert "base_cnn.fc.weight" not in grads
    for name, d in gold["buffers"].items():
        b = nb.get(name)
        if b is None:  # buffer of a module the forward never calls (e.g. unused layers)
            b = p[name]
        assert close(float(b.double().norm()), d["norm"], 1e-4), name


def test_quadrant_order_and_odd_split():
    """Region assignment is integer work: TL, TR, BL, BR, second half takes the extra row/col."""
    x = torch.arange(2 * 1 * 5 * 7, dtype=torch.float32).reshape(2, 1, 5, 7)
    q = O.quadrants(x)
    assert [tuple(t.shape[2:]) for t in q] == [(2, 3), (2, 4), (3, 3), (3, 4)]
    assert torch.equal(q[0], x[:, :, :2, :3]) and torch.equal(q[3], x[:, :, 2:, 3:])


def test_feature_row_layout():
    """[global512 | TL | TR | BL | BR] with c*9+h*3+w inside a quadrant (QS/models.py:291-294)."""
    p = O.make_params("quadtree", 8, seed=0)
    images, numerical, _ = O.synthetic_batch(1, 5)
    taps = {}
    with torch.no_grad():
        O.quadtree_forward(p, images, numerical, training=False, taps=taps)
    f = taps["image_features"]
    assert f.shape == (1, 5120)
    base = taps["base_features"]
    assert base.shape == (1, 256, 14, 14)
    q_tr = base[:, :, :7, 7:]
    y = torch.nn.functional.max_pool2d(torch.relu(torch.nn.functional.conv2d(
        q_tr, p["quadrant_processor.0.weight"], p["quadrant_processor.0.bias"], padding=1)), 2, 2)
    c, h, w = 17, 2, 1
    assert torch.allclose(f[0, 512 + 1152 + c * 9 + h * 3 + w], y[0, c, h, w])
    assert torch.allclose(f[0, :512], taps["layer4"].mean(dim=(2, 3))[0], atol=1e-6)
