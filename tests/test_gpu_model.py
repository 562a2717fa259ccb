"""End-to-end parity of the drop-in QuadtreeCNN (CUDA path through the C ABI) against the oracle.

The oracle (fp32, functional) is evaluated on the GPU in fp32 for speed; its pin to the reference is the CPU
golden test. Because the product computes in bf16, gradients are judged relative to what PyTorch's own bf16
autocast does on the same oracle graph — the contract lives in tests/parity.py:
logits max_abs <= 3e-2 * max|logit|, loss |d| <= 2e-3, per-parameter gradient cosine >= autocast cosine - 0.02,
rel_L2 <= 1.25 x autocast rel_L2.
"""
import pytest
import torch
import torch.nn.functional as F

import parity
from oracle.loading import load_oracle_params

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    from oracle import quadtree_oracle as O
    from qtcnn_b200 import models as M
    return O, M


def cos(a, b):
    return float(F.cosine_similarity(a.double().flatten(), b.double().flatten(), dim=0))


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def oracle_on_gpu(O, kind, p, inputs, labels, autocast=False, **kw):
    dev = torch.device("cuda")
    pg = {k: v.to(dev) for k, v in p.items()}
    ins = tuple(t.to(dev) for t in inputs)
    if autocast:
        with torch.autocast("cuda", dtype=torch.bfloat16):
            return O.loss_and_grads(kind, pg, ins, labels.to(dev), training=True, **kw)
    return O.loss_and_grads(kind, pg, ins, labels.to(dev), training=True, **kw)


def test_quadtree_train_step_matches_oracle(env):
    O, M = env
    B = 16
    p = O.make_params("quadtree", 8, seed=0)
    images, numerical, labels = O.synthetic_batch(B, 1234)
    ref_logits, ref_loss, ref_g, ref_nb = oracle_on_gpu(O, "quadtree", p, (images, numerical), labels)
    ac_logits, ac_loss, ac_g, _ = oracle_on_gpu(O, "quadtree", p, (images, numerical), labels, autocast=True)

    model = M.QuadtreeCNN(num_classes=8, dropout_rate=0.0)
    load_oracle_params(model, p)
    model = model.cuda().train()
    logits = model(images.cuda(), numerical.cuda())
    assert logits.dtype == torch.float32 and logits.shape == (B, 8)
    loss = F.cross_entropy(logits, labels.cuda())
    loss.backward()
    torch.cuda.synchronize()

    lmax = float(ref_logits.abs().max())
    lerr = float((logits.detach() - ref_logits).abs().max())
    aerr = float((ac_logits.float() - ref_logits).abs().max())
    print(f"logits max_abs err ours {lerr:.3e} autocast {aerr:.3e} (max |logit| {lmax:.3f}); loss ours {float(loss):.5f} "
          f"autocast {float(ac_loss):.5f} fp32 {float(ref_loss):.5f}")
    parity.assert_logits_loss(logits, loss, ref_logits, ref_loss)

    named = dict(model.named_parameters())
    worst = []
    for name, g_ref in ref_g.items():
        prm = named[name]
        assert prm.grad is not None, f"no gradient for {name}"
        c_o, c_a = cos(prm.grad, g_ref), cos(ac_g[name], g_ref)
        r_o, r_a = rel(prm.grad, g_ref), rel(ac_g[name], g_ref)
        worst.append((c_o - c_a, name, c_o, c_a, r_o, r_a))
    worst.sort()
    for dlt, name, c_o, c_a, r_o, r_a in worst[:12]:
        print(f"  {name:40s} cos ours {c_o:.4f} autocast {c_a:.4f} | rel_l2 ours {r_o:.3f} autocast {r_a:.3f}")
    for name, g_ref in ref_g.items():
        parity.assert_grad(name, named[name].grad, g_ref, ac_g[name])
    # base_cnn.fc is registered but never used by the reference forward: no gradient
    assert named["base_cnn.fc.weight"].grad is None
    # BN running statistics were updated like nn.BatchNorm2d would
    sd = model.state_dict()
    for k in ("base_cnn.bn1.running_mean", "base_cnn.layer3.1.bn2.running_var", "base_cnn.layer4.0.downsample.1.running_mean"):
        assert rel(sd[k], ref_nb[k]) < 2e-2, (k, rel(sd[k], ref_nb[k]))
    assert int(sd["base_cnn.bn1.num_batches_tracked"]) == 1


def test_eval_forward_and_state_dict_roundtrip(env):
    O, M = env
    p = O.make_params("quadtree", 8, seed=3)
    # non-trivial running statistics
    g = torch.Generator().manual_seed(7)
    for k in list(p):
        if k.endswith("running_mean"):
            p[k] = 0.1 * torch.randn(p[k].shape, generator=g)
        if k.endswith("running_var"):
            p[k] = 0.5 + torch.rand(p[k].shape, generator=g)
    images, numerical, labels = O.synthetic_batch(4, 99)
    with torch.no_grad():
        ref = O.quadtree_forward({k: v.cuda() for k, v in p.items()}, images.cuda(), numerical.cuda(), training=False)
    model = M.QuadtreeCNN(num_classes=8)
    load_oracle_params(model, p)
    model = model.cuda().eval()
    with torch.no_grad():
        out = model(images.cuda(), numerical.cuda())
    err = float((out - ref).abs().max() / ref.abs().max())
    print("eval logits rel max err", err)
    assert err < 3e-2
    # state_dict keys (252 incl. aliases) and a load into a second instance give identical outputs
    sd = model.state_dict()
    assert len(sd) == 252 and "features_extractor.6.1.bn2.weight" in sd and "global_processor.0.0.conv1.weight" in sd
    m2 = M.QuadtreeCNN(num_classes=8).cuda().eval()
    m2.load_state_dict(sd)
    with torch.no_grad():
        out2 = m2(images.cuda(), numerical.cuda())
    assert torch.equal(out, out2)


@pytest.mark.parametrize("mode", ["fusion", "image_only", "numerical_only"])
def test_frozen_mode_variants(env, mode):
    """resnet/models.py QuadtreeCNN: frozen backbone (BN still in train mode), mode switch."""
    O, M = env
    p = O.make_params("quadtree", 8, seed=5, mode=mode)
    images, numerical, labels = O.synthetic_batch(8, 321)
    ref_logits, ref_loss, ref_g, _ = oracle_on_gpu(O, "quadtree", p, (images, numerical), labels, mode=mode)
    _, _, ac_g, _ = oracle_on_gpu(O, "quadtree", p, (images, numerical), labels, autocast=True, mode=mode)
    model = M.get_model_resnet(8, "cuda", mode=mode, print_num_params=False)
    model.dropout_rate = 0.0
    load_oracle_params(model, p)
    model.train()
    logits = model(images.cuda(), numerical.cuda())
    loss = F.cross_entropy(logits, labels.cuda())
    loss.backward()
    parity.assert_logits_loss(logits, loss, ref_logits, ref_loss)
    named = dict(model.named_parameters())
    assert named["base_cnn.layer1.0.conv1.weight"].grad is None  # frozen
    checked, report = 0, []
    for name, g_ref in ref_g.items():  # every trainable parameter, autocast-relative bounds (tests/parity.py)
        if name.startswith("base_cnn."):
            continue
        assert named[name].grad is not None, name
        checked += parity.assert_grad(name, named[name].grad, g_ref, ac_g[name], report)
    parity.print_worst(report)
    assert checked >= (4 if mode == "numerical_only" else 6)


def test_gradcam_hooks_on_layer4(env):
    """resnet/grad_cam_analysis.py:251-286: forward + full-backward hooks on base_cnn.layer4, eval mode,
    backward from a one-hot gradient on the logits."""
    O, M = env
    p = O.make_params("quadtree", 8, seed=2)
    images, numerical, _ = O.synthetic_batch(1, 17)
    model = M.QuadtreeCNN(num_classes=8, freeze_backbone=False)
    load_oracle_params(model, p)
    model = model.cuda().eval()
    target = model.base_cnn.layer4
    h1 = target.register_forward_hook(model.save_activation_hook)
    h2 = target.register_full_backward_hook(model.save_gradient_hook)
    x = images.cuda().requires_grad_(True)
    out = model(x, numerical.cuda())
    one_hot = torch.zeros_like(out)
    one_hot[0, int(out.argmax())] = 1
    out.backward(gradient=one_hot, retain_graph=True)
    h1.remove(); h2.remove()
    assert model.activations is not None and tuple(model.activations.shape) == (1, 512, 7, 7)
    assert model.gradients is not None and tuple(model.gradients.shape) == (1, 512, 7, 7)
    # reference value of d logit / d layer4 from the oracle
    pg = {k: v.cuda() for k, v in p.items()}
    taps = {}
    logits = O.quadtree_forward(pg, images.cuda().requires_grad_(True), numerical.cuda(), training=False, taps=taps)
    gref = torch.autograd.grad(logits[0, int(out.argmax())], taps["layer4"])[0] if taps["layer4"].requires_grad else None
    if gref is not None:
        assert cos(model.gradients.float(), gref) > 0.98
    cam = F.relu((model.gradients.float().mean(dim=(2, 3), keepdim=True) * model.activations.float()).sum(1))
    assert torch.isfinite(cam).all()


def test_dropout_training_runs_and_is_seeded(env):
    O, M = env
    p = O.make_params("quadtree", 8, seed=0)
    images, numerical, labels = O.synthetic_batch(4, 5)
    model = M.QuadtreeCNN(num_classes=8, dropout_rate=0.5)
    load_oracle_params(model, p)
    model = model.cuda().train()
    torch.manual_seed(11)
    a = model(images.cuda(), numerical.cuda()).detach()
    torch.manual_seed(11)
    b = model(images.cuda(), numerical.cuda()).detach()
    c = model(images.cuda(), numerical.cuda()).detach()
    assert torch.equal(a, b) and not torch.equal(a, c)
    loss = F.cross_entropy(model(images.cuda(), numerical.cuda()), labels.cuda())
    loss.backward()
    assert all(torch.isfinite(q.grad).all() for q in model.parameters() if q.grad is not None)


def test_weights_refresh_after_fused_adam_step(env):
    """Fused Adam does not bump Tensor._version; the packed bf16 weight copies must still follow the update."""
    O, M = env
    p = O.make_params("quadtree", 8, seed=0)
    images, numerical, labels = O.synthetic_batch(4, 77)
    model = M.QuadtreeCNN(num_classes=8, dropout_rate=0.0)
    load_oracle_params(model, p)
    model = model.cuda().train()
    opt = torch.optim.Adam([q for q in model.parameters() if q.requires_grad], lr=1e-2, fused=True)
    x, nf, y = images.cuda(), numerical.cuda(), labels.cuda()
    before = model(x, nf).detach().clone()
    F.cross_entropy(model(x, nf), y).backward()
    opt.step()
    after = model(x, nf).detach()
    assert float((after - before).abs().max()) > 1e-3, "forward ignored the optimizer update (stale packed weights)"
    # and it matches the oracle evaluated with the updated fp32 parameters
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    ref = O.quadtree_forward(sd, x, nf, training=True)
    assert float((after - ref).abs().max()) <= 4e-2 * float(ref.abs().max())
