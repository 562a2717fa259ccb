"""Parity of the remaining drop-in classes against the oracle (evaluated in fp32 on the GPU):
AttentionHierarchicalCNN / HierarchicalQuadtreeCNN (level-1 + level-2 quadtree), StandardResNetCNN (frozen) and
Quadtree3DCNN (Conv3d stack, both modes). Tolerances: the SURVEY §8(d) contract in tests/parity.py (autocast-relative)."""
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


def to_cuda(p):
    return {k: v.cuda() for k, v in p.items()}


def oracle_names(model, name):
    """product parameter name -> oracle name for classes that keep the ResNet only via features_extractor/global_processor"""
    back = {"features_extractor.0": "conv1", "features_extractor.1": "bn1", "features_extractor.4": "layer1",
            "features_extractor.5": "layer2", "global_processor.0": "layer3", "global_processor.1": "layer4"}
    if hasattr(model, "base_cnn"):
        return name
    for a, b in back.items():
        if name.startswith(a + "."):
            return "base_cnn." + b + name[len(a):]
    return name


def check_train_step(O, kind, model, p, inputs, labels, **kw):
    (ref_logits, ref_loss, ref_g, _), (_, _, ac_g, _) = parity.oracle_fp32_and_autocast(O, kind, p, inputs, labels, **kw)
    logits = model(*(t.cuda() for t in inputs))
    loss = F.cross_entropy(logits, labels.cuda())
    loss.backward()
    print(f"{kind}: loss {float(loss.detach()):.5f} vs {float(ref_loss):.5f}")
    parity.assert_logits_loss(logits, loss.detach(), ref_logits, ref_loss)
    checked, report = 0, []
    seen = set()
    for name, prm in model.named_parameters():
        if id(prm) in seen or not prm.requires_grad:
            continue
        seen.add(id(prm))
        on = oracle_names(model, name)
        if on not in ref_g:
            assert prm.grad is None, f"{name} received a gradient the oracle does not produce"
            continue
        assert prm.grad is not None, f"no gradient for {name}"
        checked += parity.assert_grad(name, prm.grad, ref_g[on], ac_g[on], report)
    parity.print_worst(report)
    assert checked > 5
    return logits


def test_attention_hierarchical_train_and_eval(env):
    O, M = env
    p = O.make_params("attention_hierarchical", 8, seed=1)
    images, numerical, labels = O.synthetic_batch(8, 99)
    model = M.AttentionHierarchicalCNN(num_classes=8, dropout_rate=0.0)
    load_oracle_params(model, p)
    model = model.cuda().train()
    assert len(model.state_dict()) == len([k for k in model.state_dict()])
    check_train_step(O, "attention_hierarchical", model, p, (images, numerical), labels)
    model.eval()
    sd = {oracle_names(model, k): v for k, v in model.state_dict().items()}
    with torch.no_grad():
        out = model(images.cuda(), numerical.cuda())
        ref = O.attention_hier_forward(sd, images.cuda(), numerical.cuda(), training=False)
    assert float((out - ref).abs().max()) <= parity.LOGIT_TOL * float(ref.abs().max())


def test_hierarchical_quadtree_train(env):
    O, M = env
    p = O.make_params("hierarchical_quadtree", 8, seed=4)
    images, numerical, labels = O.synthetic_batch(8, 5)
    model = M.get_model("hierarchical_quadtree", 8, "cuda", print_num_params=False)
    model.dropout_rate = 0.0
    load_oracle_params(model, p)
    model.train()
    assert sum(q.numel() for q in model.parameters()) == 13_641_480  # SURVEY §0.2 [probe]
    check_train_step(O, "hierarchical_quadtree", model, p, (images, numerical), labels)


def test_standard_resnet_frozen(env):
    O, M = env
    p = O.make_params("standard_resnet", 8, seed=2)
    images, numerical, labels = O.synthetic_batch(8, 55)
    model = M.get_model_resnet(8, "cuda", mode="standard_resnet_only", print_num_params=False)
    model.dropout_rate = 0.0
    load_oracle_params(model, p)
    model.train()
    (ref_logits, ref_loss, ref_g, _), (_, _, ac_g, _) = parity.oracle_fp32_and_autocast(O, "standard_resnet", p, (images, numerical), labels)
    logits = model(images.cuda(), None)
    loss = F.cross_entropy(logits, labels.cuda())
    loss.backward()
    parity.assert_logits_loss(logits, loss.detach(), ref_logits, ref_loss)
    named = dict(model.named_parameters())
    assert named["base_cnn.layer4.1.conv2.weight"].grad is None
    report = []
    for n in ("classifier.0.weight", "classifier.3.weight", "classifier.0.bias", "classifier.3.bias"):
        parity.assert_grad(n, named[n].grad, ref_g[n], ac_g[n], report)
    parity.print_worst(report)


@pytest.mark.parametrize("mode", ["quadtree_3d_fusion", "quadtree_3d_image_only"])
def test_quadtree3d(env, mode):
    O, M = env
    p = O.make_params("quadtree3d", 8, seed=6, mode=mode)
    clips, numerical, labels = O.synthetic_batch(4, 11, seq_len=4, clip_size=32)
    model = M.Quadtree3DCNN(num_classes=8, sequence_length=4, dropout_rate=0.0, mode=mode)
    load_oracle_params(model, p)
    model = model.cuda().train()
    assert sum(q.numel() for q in M.Quadtree3DCNN(8).parameters()) == 9_992_024  # SURVEY §8a14 [probe] (fusion)
    check_train_step(O, "quadtree3d", model, p, (clips, numerical), labels, mode=mode)
    # the conv stack alone (north-star: Conv3d forward/backward of the 3-D model)
    with torch.no_grad():
        model.eval()
        feats = model.conv_stack(clips.cuda())
        sd = model.state_dict()
        ref = O.quadtree3d_conv_stack({k: v for k, v in sd.items()}, clips.cuda(), training=False)
    assert float((feats - ref).abs().max()) <= parity.LOGIT_TOL * float(ref.abs().max()) + 1e-3


def test_cnn_lstm(env):
    """CnnLstm (cnn+lstm/models.py:14-89): frames through the frozen tensor-core ResNet-18 (BatchNorm in
    train mode over B*T frames), per-step MLP, torch LSTM, classifier — logits and every trainable gradient vs oracle."""
    O, M = env
    p = O.make_params("cnn_lstm", 8, seed=8)
    frames, numerical, labels = O.synthetic_batch(3, 21, seq_len=4, clip_size=64)
    model = M.get_model_seq("cnn_lstm", 8, "cuda", seq_len=4)
    model.dropout_rate = 0.0
    model.lstm.dropout = 0.0
    load_oracle_params(model, p)
    model.train()
    assert sum(q.numel() for q in model.parameters()) == 12_678_984
    assert sum(q.numel() for q in model.parameters() if q.requires_grad) == 1_502_472
    check_train_step(O, "cnn_lstm", model, p, (frames, numerical), labels)
    assert model.cnn_backbone[7][1].conv2.weight.grad is None  # frozen backbone
    with pytest.raises(ValueError):
        M.get_model_seq("3d_cnn", 8, "cuda")


def test_maxpool3d_exact(env):
    import qtcnn_b200.capi as C
    n, d, h, w, c = 2, 4, 6, 10, 16
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(n, c, d, h, w, device="cuda", generator=g).to(torch.bfloat16).relu()
    xf = x.float().requires_grad_(True)
    ref = F.max_pool3d(xf, (2, 2, 2), (2, 2, 2))
    xb = x.permute(0, 2, 3, 4, 1).contiguous()
    out = torch.empty(n, d // 2, h // 2, w // 2, c, device="cuda", dtype=torch.bfloat16)
    am = torch.empty(out.shape, device="cuda", dtype=torch.int8)
    C.check(C.lib().qt_maxpool3d_fwd(C.ptr(xb), C.ptr(out), C.ptr(am), n, d, h, w, c, 2, 2, 2, C.stream()))
    assert torch.equal(out.permute(0, 4, 1, 2, 3).float(), ref.detach())
    dout = torch.randn_like(ref).to(torch.bfloat16)
    ref.backward(dout.float())
    dx = torch.empty_like(xb)
    C.check(C.lib().qt_maxpool3d_bwd(C.ptr(dout.permute(0, 2, 3, 4, 1).contiguous()), C.ptr(am), C.ptr(dx), n, d, h, w, c, 2, 2, 2,
                                     C.stream()))
    assert torch.equal(dx.permute(0, 4, 1, 2, 3).float(), xf.grad)


@pytest.mark.parametrize("kind", ["resnet3d_video", "hybrid3d"])
def test_r3d18_models(env, kind):
    """ResNet3DVideo / HybridQuadtree3DCNN (3dcnn/models.py:220-375; torchvision r3d_18, frozen except layer4): the stem's
    3x7x7 strided Conv3d, the 3-D BasicBlocks (3x3x3 stride 1 on the slab kernels, stride 2 and 1x1x1 downsample on the gather
    kernels), BatchNorm3d in train mode on the frozen layers, logits and every trainable gradient against the oracle."""
    O, M = env
    mode = "hybrid_quadtree_3d_fusion"
    p = O.make_params(kind, 8, seed=9, mode=mode)
    # B = 6: layer4 works on 1 x 4 x 4 maps here, so its train-mode BatchNorm sees only 16 B values per channel; at B = 2 a
    # single ReLU flip from bf16 rounding moves those statistics by more than the autocast-relative 1.25x bound allows
    clips, numerical, labels = O.synthetic_batch(6, 31, seq_len=8, clip_size=64)
    if kind == "resnet3d_video":
        model = M.get_model_3d(8, "cuda", mode="resnet_3d_video_only", print_num_params=False)
        kw = {}
    else:
        model = M.get_model_3d(8, "cuda", mode=mode, sequence_length=8, print_num_params=False)
        model.numerical_lstm.dropout = 0.0
        kw = {"mode": mode}
    model.dropout_rate = 0.0
    load_oracle_params(model, p)
    model.train()
    sd = model.state_dict()
    assert all(k in sd for k in p), [k for k in p if k not in sd][:4]  # torchvision's r3d_18 key names
    trainable = [n for n, q in model.named_parameters() if q.requires_grad]
    assert any("layer4" in n or ".4." in n for n in trainable) and not any("layer3" in n or ".3.0.conv" in n for n in trainable)
    check_train_step(O, kind, model, p, (clips, numerical), labels, **kw)
