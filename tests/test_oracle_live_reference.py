"""Extra pin of the oracle against the LIVE reference modules, on seeds / batch sizes the committed fixtures do not
cover. Runs only where /root/reference exists (the build container); skipped everywhere else — the committed fixtures
under tests/golden/ are what travels."""
import importlib.util
import os

import pytest
import torch

from oracle import quadtree_oracle as O

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")


@pytest.fixture(scope="module")
def gold():
    path = os.path.join(os.path.dirname(__file__), "golden", "make_golden.py")
    spec = importlib.util.spec_from_file_location("make_golden", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


CASES = [
    {"name": "live_quadtree_b3", "kind": "quadtree", "batch": 3, "seed": 2024, "param_seed": 11},
    {"name": "live_attention_b3_eval", "kind": "attention_hierarchical", "batch": 3, "seed": 2025, "param_seed": 12, "training": False},
    {"name": "live_cnn_lstm_b2", "kind": "cnn_lstm", "batch": 2, "seed": 2026, "param_seed": 13, "seq_len": 2, "clip": 96},
]


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_oracle_matches_live_reference(gold, case):
    torch.set_num_threads(max(1, min(8, os.cpu_count() or 1)))
    ref = gold.run_case(case)
    kind = case["kind"]
    p = O.make_params(kind, 8, seed=case["param_seed"])
    if kind == "cnn_lstm":
        images, numerical, labels = O.synthetic_batch(case["batch"], case["seed"], seq_len=case["seq_len"], clip_size=case["clip"])
    else:
        images, numerical, labels = O.synthetic_batch(case["batch"], case["seed"])
    ref_logits = torch.tensor(ref["logits"], dtype=torch.float64)
    if not case.get("training", True):
        with torch.no_grad():
            logits = O.FORWARDS[kind](p, images, numerical, training=False)
        assert torch.allclose(logits.double(), ref_logits, rtol=1e-4, atol=1e-5)
        return
    logits, loss, grads, _ = O.loss_and_grads(kind, p, (images, numerical), labels, training=True)
    assert torch.allclose(logits.double(), ref_logits, rtol=1e-4, atol=1e-5)
    assert abs(float(loss) - ref["loss"]) <= 1e-4 * max(1.0, abs(ref["loss"]))
    for name, d in ref["grads"].items():
        g = grads[name].double().flatten()
        assert abs(float(g.norm()) - d["norm"]) <= 2e-3 * max(d["norm"], 1e-12) + 1e-9, name
