"""BASELINE.json's full configuration (QuadtreeCNN, 224x224, batch 256) through size-independent properties and one
full-size comparison with the oracle (evaluated in fp32 on the GPU; its pin to the reference is the CPU golden test):

* a train step at batch 256 matches the oracle within the stated tolerance (logits <= 3e-2 * max|logit|, loss <= 2e-3,
  gradient cosine >= torch-bf16-autocast cosine - 0.02 and rel_L2 <= 1.25 x autocast's on every parameter (tests/parity.py), BatchNorm running statistics <= 2e-2 relative);
* the step is bitwise reproducible (fixed-order reductions everywhere: two runs give identical logits and gradients);
* eval-mode forward is equivariant under a permutation of the batch, bit for bit (no cross-sample coupling, no
  position-dependent accumulation order).
"""
import pytest
import torch
import torch.nn.functional as F

import parity
from oracle.loading import load_oracle_params

pytestmark = pytest.mark.gpu
B = 256


@pytest.fixture(scope="module")
def setup():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    from oracle import quadtree_oracle as O
    from qtcnn_b200 import models as M
    p = O.make_params("quadtree", 8, seed=0)
    images, numerical, labels = O.synthetic_batch(B, 1234)
    return O, M, p, images.cuda(), numerical.cuda(), labels.cuda()


def make_model(M, p, train):
    model = M.QuadtreeCNN(num_classes=8, dropout_rate=0.0)
    load_oracle_params(model, p)
    return model.cuda().train(train)


def cos(a, b):
    return float(F.cosine_similarity(a.double().flatten(), b.double().flatten(), dim=0))


def train_step(model, images, numerical, labels):
    model.zero_grad(set_to_none=True)
    logits = model(images, numerical)
    loss = F.cross_entropy(logits, labels)
    loss.backward()
    torch.cuda.synchronize()
    grads = {n: q.grad.detach().clone() for n, q in model.named_parameters() if q.grad is not None}
    return logits.detach().clone(), loss.detach().clone(), grads


def test_full_size_train_step_vs_oracle(setup):
    O, M, p, images, numerical, labels = setup
    pg = {k: v.cuda() for k, v in p.items()}
    ref_logits, ref_loss, ref_g, ref_nb = O.loss_and_grads("quadtree", pg, (images, numerical), labels, training=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):  # what the reference's own framework does in bf16
        _, _, ac_g, _ = O.loss_and_grads("quadtree", pg, (images, numerical), labels, training=True)
    model = make_model(M, p, True)
    logits, loss, grads = train_step(model, images, numerical, labels)
    parity.assert_logits_loss(logits, loss, ref_logits, ref_loss)
    seen, report = 0, []
    for name, g in grads.items():
        if name.startswith(("features_extractor.", "global_processor.")) or name not in ref_g:
            continue
        seen += parity.assert_grad(name, g, ref_g[name], ac_g[name], report)
    parity.print_worst(report)
    assert seen >= 60
    sd = model.state_dict()
    for name, ref in ref_nb.items():
        if name.endswith("running_mean") or name.endswith("running_var"):
            err = float((sd[name].double() - ref.double()).norm() / (ref.double().norm() + 1e-12))
            assert err <= 2e-2, (name, err)
    # every gradient the oracle produces exists here and vice versa (base_cnn.fc never receives one)
    mine = {n for n in grads if not n.startswith(("features_extractor.", "global_processor."))}
    assert mine == set(ref_g), (sorted(mine ^ set(ref_g))[:6])


def test_full_size_step_is_bitwise_reproducible(setup):
    O, M, p, images, numerical, labels = setup
    a = train_step(make_model(M, p, True), images, numerical, labels)
    b = train_step(make_model(M, p, True), images, numerical, labels)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    assert a[2].keys() == b[2].keys()
    for name in a[2]:
        assert torch.equal(a[2][name], b[2][name]), name


def test_full_size_eval_is_permutation_equivariant(setup):
    O, M, p, images, numerical, labels = setup
    model = make_model(M, p, False)
    g = torch.Generator(device="cuda").manual_seed(3)
    perm = torch.randperm(B, device="cuda", generator=g)
    with torch.no_grad():
        base = model(images, numerical)
        shuffled = model(images[perm].contiguous(), numerical[perm].contiguous())
    assert torch.equal(shuffled, base[perm])
    assert bool(torch.isfinite(base).all())


def test_level12_inference_at_config_batch():
    """BASELINE.json configs[1]: AttentionHierarchicalCNN (level-1 2x2 + level-2 4x4 quadtree) inference, 224x224, batch 256,
    eval mode with non-trivial running statistics, against the fp32 oracle evaluated on the GPU; plus batch-permutation
    equivariance (region assignment and pooling never couple samples)."""
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from oracle import quadtree_oracle as O
    from qtcnn_b200 import models as M
    p = O.make_params("attention_hierarchical", 8, seed=1)
    g = torch.Generator().manual_seed(7)
    for k in list(p):
        if k.endswith("running_mean"):
            p[k] = 0.1 * torch.randn(p[k].shape, generator=g)
        if k.endswith("running_var"):
            p[k] = 0.5 + torch.rand(p[k].shape, generator=g)
    images, numerical, _ = O.synthetic_batch(B, 99)
    images, numerical = images.cuda(), numerical.cuda()
    model = M.AttentionHierarchicalCNN(num_classes=8)
    load_oracle_params(model, p)
    model = model.cuda().eval()
    with torch.no_grad():
        out = model(images, numerical)
        ref = O.attention_hier_forward({k: v.cuda() for k, v in p.items()}, images, numerical, training=False)
        perm = torch.randperm(B, device="cuda", generator=torch.Generator(device="cuda").manual_seed(3))
        shuffled = model(images[perm].contiguous(), numerical[perm].contiguous())
    assert out.shape == (B, 8) and out.dtype == torch.float32
    err = float((out - ref).abs().max())
    print(f"level-1+2 inference B={B}: logits max abs err {err:.3e} / max |logit| {float(ref.abs().max()):.3f}")
    assert err <= parity.LOGIT_TOL * float(ref.abs().max())
    assert torch.equal(shuffled, out[perm])
