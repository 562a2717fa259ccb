"""CPU oracle for the QuadtreeCNN hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain-PyTorch fp32 *functional* restatement of the reference models' arithmetic (no nn.Module from the
reference, no CUDA, no kernels from this repo). Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it, and only as the checker / baseline.

Pinned against the reference: `tests/golden/make_golden.py` (run in the build container, where
/root/reference is mounted) loads the same seeded parameters into the UNMODIFIED reference modules and stores
their logits / losses / gradient digests under `tests/golden/*.json`; `tests/test_oracle_golden.py` replays
them through this file. The reference ships no tests or golden vectors of its own, so these
reference-generated fixtures are the pin.

Reference files restated here (paths relative to the reference root; "QS" = "Quadtree_from scratch"):
  QS/models.py:214-305   QuadtreeCNN            -> quadtree_forward
  QS/models.py:6-101     AttentionHierarchicalCNN -> attention_hier_forward
  QS/models.py:105-210   HierarchicalQuadtreeCNN (intended slicing) -> hier_forward
  resnet/models.py:7-180 StandardResNetCNN / QuadtreeCNN(mode=...) -> standard_resnet_forward / quadtree_forward(mode=)
  3dcnn/models.py:96-214 Quadtree3DCNN          -> quadtree3d_forward
  3dcnn/models.py:220-375 ResNet3DVideo / HybridQuadtree3DCNN (torchvision video r3d_18) -> resnet3d_video_forward / hybrid3d_forward
  torchvision/models/resnet.py:59-105,266-284 (BasicBlock, ResNet._forward_impl; torchvision 0.26.0)
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]
BN_EPS = 1e-5
BN_MOMENTUM = 0.1


# ------------------------------------------------------------------------------------------------
# Seeded parameter construction (independent of the reference so fixtures can be regenerated anywhere).
# Names are torchvision's ResNet-18 names under "base_cnn." plus the reference heads' own names.
# ------------------------------------------------------------------------------------------------
def _conv_w(g, cout, cin, *k):
    fan_out = cout * math.prod(k)
    return torch.randn(cout, cin, *k, generator=g) * math.sqrt(2.0 / fan_out)


def _linear(g, p, name, nout, nin):
    bound = 1.0 / math.sqrt(nin)
    p[name + ".weight"] = (torch.rand(nout, nin, generator=g) * 2 - 1) * bound
    p[name + ".bias"] = (torch.rand(nout, generator=g) * 2 - 1) * bound


def _bn(g, p, name, c, ndim_tag=""):
    p[name + ".weight"] = 0.5 + torch.rand(c, generator=g)
    p[name + ".bias"] = 0.2 * torch.randn(c, generator=g)
    p[name + ".running_mean"] = torch.zeros(c)
    p[name + ".running_var"] = torch.ones(c)
    p[name + ".num_batches_tracked"] = torch.zeros((), dtype=torch.long)


def resnet18_params(g, prefix="base_cnn.") -> Params:
    p: Params = {}
    p[prefix + "conv1.weight"] = _conv_w(g, 64, 3, 7, 7)
    _bn(g, p, prefix + "bn1", 64)
    cin = 64
    for li, cout in enumerate([64, 128, 256, 512], start=1):
        for bi in range(2):
            stride = 2 if (li > 1 and bi == 0) else 1
            b = f"{prefix}layer{li}.{bi}."
            p[b + "conv1.weight"] = _conv_w(g, cout, cin, 3, 3)
            _bn(g, p, b + "bn1", cout)
            p[b + "conv2.weight"] = _conv_w(g, cout, cout, 3, 3)
            _bn(g, p, b + "bn2", cout)
            if stride != 1 or cin != cout:
                p[b + "downsample.0.weight"] = _conv_w(g, cout, cin, 1, 1)
                _bn(g, p, b + "downsample.1", cout)
            cin = cout
    _linear(g, p, prefix + "fc", 1000, 512)
    return p


def make_params(kind: str, num_classes: int = 8, seed: int = 0, numerical_feature_dim: int = 47,
                mode: str = "fusion", cnn_3d_feature_dim: int = 1024) -> Params:
    g = torch.Generator().manual_seed(seed)
    if kind == "quadtree":  # QS/models.py:216-271 (mode variants resnet/models.py:115-126)
        p = resnet18_params(g)
        p["quadrant_processor.0.weight"] = _conv_w(g, 128, 256, 3, 3)
        p["quadrant_processor.0.bias"] = 0.1 * torch.randn(128, generator=g)
        _linear(g, p, "numerical_mlp.0", numerical_feature_dim * 2, numerical_feature_dim)
        _linear(g, p, "numerical_mlp.3", 256, numerical_feature_dim * 2)
        din = {"fusion": 5376, "image_only": 5120, "numerical_only": 256}[mode]
        _linear(g, p, "classifier.0", din // 2, din)
        _linear(g, p, "classifier.3", num_classes, din // 2)
        return p
    if kind in ("attention_hierarchical", "hierarchical_quadtree"):  # QS/models.py:7-54, 107-165
        p = resnet18_params(g)
        p["quadrant_processor.0.weight"] = _conv_w(g, 128, 128, 3, 3)
        p["quadrant_processor.0.bias"] = 0.1 * torch.randn(128, generator=g)
        p["sub_quadrant_processor.0.weight"] = _conv_w(g, 64, 128, 3, 3)
        p["sub_quadrant_processor.0.bias"] = 0.1 * torch.randn(64, generator=g)
        if kind == "attention_hierarchical":
            _linear(g, p, "attention_gate.0", 32, 64)
            _linear(g, p, "attention_gate.2", 1, 32)
            img = 512 + 4 * 128 + 64
        else:
            img = 512 + 4 * 128 + 16 * 64
        _linear(g, p, "numerical_mlp.0", 128, numerical_feature_dim)
        _linear(g, p, "classifier.0", 1024, img + 128)
        _linear(g, p, "classifier.3", num_classes, 1024)
        return p
    if kind == "standard_resnet":  # resnet/models.py:7-54
        p = resnet18_params(g)
        _linear(g, p, "classifier.0", 256, 512)
        _linear(g, p, "classifier.3", num_classes, 256)
        return p
    if kind == "quadtree3d":  # 3dcnn/models.py:96-181
        p = {}
        chans = [(3, 32), (32, 64), (64, 128), (128, 256), (256, cnn_3d_feature_dim)]
        names = ["conv3d_block1", "conv3d_block2", "conv3d_block3", "conv3d_block4_new", "conv3d_final_features"]
        for (ci, co), nm in zip(chans, names):
            p[nm + ".0.weight"] = _conv_w(g, co, ci, 3, 3, 3)
            p[nm + ".0.bias"] = 0.1 * torch.randn(co, generator=g)
            _bn(g, p, nm + ".1", co)
        hid = numerical_feature_dim * 4
        for layer, nin in ((0, numerical_feature_dim), (1, hid)):
            bound = 1.0 / math.sqrt(hid)
            p[f"numerical_lstm.weight_ih_l{layer}"] = (torch.rand(4 * hid, nin, generator=g) * 2 - 1) * bound
            p[f"numerical_lstm.weight_hh_l{layer}"] = (torch.rand(4 * hid, hid, generator=g) * 2 - 1) * bound
            p[f"numerical_lstm.bias_ih_l{layer}"] = (torch.rand(4 * hid, generator=g) * 2 - 1) * bound
            p[f"numerical_lstm.bias_hh_l{layer}"] = (torch.rand(4 * hid, generator=g) * 2 - 1) * bound
        _linear(g, p, "numerical_projection.0", cnn_3d_feature_dim // 2, hid)
        din = cnn_3d_feature_dim + cnn_3d_feature_dim // 2 if mode == "quadtree_3d_fusion" else cnn_3d_feature_dim
        _linear(g, p, "classifier.0", din // 2, din)
        _linear(g, p, "classifier.3", num_classes, din // 2)
        return p
    if kind in ("resnet3d_video", "hybrid3d"):  # 3dcnn/models.py:220-262, 266-340 (torchvision r3d_18 names)
        pre = "r3d_model." if kind == "resnet3d_video" else "pretrained_image_extractor."
        p = {}
        stem = pre + ("stem." if kind == "resnet3d_video" else "0.")
        p[stem + "0.weight"] = _conv_w(g, 64, 3, 3, 7, 7)
        _bn(g, p, stem + "1", 64)
        cin = 64
        for li, cout in enumerate([64, 128, 256, 512], start=1):
            for bi in range(2):
                stride = 2 if (li > 1 and bi == 0) else 1
                b = f"{pre}layer{li}.{bi}." if kind == "resnet3d_video" else f"{pre}{li}.{bi}."
                p[b + "conv1.0.weight"] = _conv_w(g, cout, cin, 3, 3, 3)
                _bn(g, p, b + "conv1.1", cout)
                p[b + "conv2.0.weight"] = _conv_w(g, cout, cout, 3, 3, 3)
                _bn(g, p, b + "conv2.1", cout)
                if stride != 1 or cin != cout:
                    p[b + "downsample.0.weight"] = _conv_w(g, cout, cin, 1, 1, 1)
                    _bn(g, p, b + "downsample.1", cout)
                cin = cout
        if kind == "resnet3d_video":
            _linear(g, p, pre + "fc.0", 256, 512)
            _linear(g, p, pre + "fc.3", num_classes, 256)
            return p
        hid = numerical_feature_dim * 4
        for layer, nin in ((0, numerical_feature_dim), (1, hid)):
            bound = 1.0 / math.sqrt(hid)
            p[f"numerical_lstm.weight_ih_l{layer}"] = (torch.rand(4 * hid, nin, generator=g) * 2 - 1) * bound
            p[f"numerical_lstm.weight_hh_l{layer}"] = (torch.rand(4 * hid, hid, generator=g) * 2 - 1) * bound
            p[f"numerical_lstm.bias_ih_l{layer}"] = (torch.rand(4 * hid, generator=g) * 2 - 1) * bound
            p[f"numerical_lstm.bias_hh_l{layer}"] = (torch.rand(4 * hid, generator=g) * 2 - 1) * bound
        _linear(g, p, "numerical_projection.0", 256, hid)
        din = 768 if mode == "hybrid_quadtree_3d_fusion" else 512
        _linear(g, p, "classifier.0", din // 2, din)
        _linear(g, p, "classifier.3", num_classes, din // 2)
        return p
    if kind == "cnn_lstm":  # cnn+lstm/models.py:14-57 (ResNet-18 children[:-1] as cnn_backbone.{0,1,4,5,6,7})
        r = resnet18_params(g, prefix="")
        p = {}
        for k, v in r.items():
            head, rest = k.split(".", 1)
            if head in _SEQ_INDEX:
                p[f"cnn_backbone.{_SEQ_INDEX[head]}.{rest}"] = v
        _linear(g, p, "numerical_mlp.0", 128, numerical_feature_dim)
        _linear(g, p, "numerical_mlp.2", 128, 128)
        hid, nin0 = 256, 512 + 128
        for layer, nin in ((0, nin0), (1, hid)):
            bound = 1.0 / math.sqrt(hid)
            p[f"lstm.weight_ih_l{layer}"] = (torch.rand(4 * hid, nin, generator=g) * 2 - 1) * bound
            p[f"lstm.weight_hh_l{layer}"] = (torch.rand(4 * hid, hid, generator=g) * 2 - 1) * bound
            p[f"lstm.bias_ih_l{layer}"] = (torch.rand(4 * hid, generator=g) * 2 - 1) * bound
            p[f"lstm.bias_hh_l{layer}"] = (torch.rand(4 * hid, generator=g) * 2 - 1) * bound
        _linear(g, p, "classifier.0", 128, hid)
        _linear(g, p, "classifier.3", num_classes, 128)
        return p
    raise ValueError(kind)


# position of the ResNet-18 children inside `nn.Sequential(*list(resnet.children())[:-1])` (cnn+lstm/models.py:23)
_SEQ_INDEX = {"conv1": 0, "bn1": 1, "layer1": 4, "layer2": 5, "layer3": 6, "layer4": 7}


def synthetic_batch(batch: int, seed: int = 1234, image_size: int = 224, num_classes: int = 8, seq_len: int = 0,
                    clip_size: int = 112):
    """SURVEY.md §8(d) synthetic inputs: N(0,1) images, pose vector with the real feature ranges
    (img process/1_prepare_still_image_dataset.py:101-113), labels in [0, num_classes)."""
    g = torch.Generator().manual_seed(seed)

    def pose(*lead):
        u = torch.rand(*lead, 47, generator=g)
        scale = torch.cat([torch.ones(33), torch.full((10,), 180.0), torch.full((3,), 4.0), torch.full((1,), 5.0)])
        return u * scale

    if seq_len:
        images = torch.randn(batch, seq_len, 3, clip_size, clip_size, generator=g)
        numerical = pose(batch, seq_len)
    else:
        images = torch.randn(batch, 3, image_size, image_size, generator=g)
        numerical = pose(batch)
    labels = torch.randint(0, num_classes, (batch,), generator=g)
    return images, numerical, labels


# ------------------------------------------------------------------------------------------------
# Building blocks
# ------------------------------------------------------------------------------------------------
def _bn_apply(p: Params, name: str, x: torch.Tensor, training: bool, new_buffers: Optional[dict]):
    """nn.BatchNorm{2,3}d: batch statistics + running-stat update in training, running stats in eval."""
    w, b = p[name + ".weight"], p[name + ".bias"]
    rm, rv = p[name + ".running_mean"], p[name + ".running_var"]
    dims = [0] + list(range(2, x.dim()))
    shape = [1, -1] + [1] * (x.dim() - 2)
    if training:
        mean = x.mean(dim=dims)
        var = x.var(dim=dims, unbiased=False)
        if new_buffers is not None:
            n = x.numel() / x.shape[1]
            with torch.no_grad():
                new_buffers[name + ".running_mean"] = (1 - BN_MOMENTUM) * rm + BN_MOMENTUM * mean.detach()
                new_buffers[name + ".running_var"] = (1 - BN_MOMENTUM) * rv + BN_MOMENTUM * var.detach() * n / max(n - 1, 1)
                new_buffers[name + ".num_batches_tracked"] = p[name + ".num_batches_tracked"] + 1
    else:
        mean, var = rm, rv
    return (x - mean.view(shape)) / torch.sqrt(var.view(shape) + BN_EPS) * w.view(shape) + b.view(shape)


def _basic_block(p, prefix, x, stride, training, nb):
    """torchvision BasicBlock.forward (resnet.py:89-105)."""
    out = F.conv2d(x, p[prefix + "conv1.weight"], None, stride=stride, padding=1)
    out = F.relu(_bn_apply(p, prefix + "bn1", out, training, nb))
    out = F.conv2d(out, p[prefix + "conv2.weight"], None, stride=1, padding=1)
    out = _bn_apply(p, prefix + "bn2", out, training, nb)
    if prefix + "downsample.0.weight" in p:
        idt = F.conv2d(x, p[prefix + "downsample.0.weight"], None, stride=stride)
        idt = _bn_apply(p, prefix + "downsample.1", idt, training, nb)
    else:
        idt = x
    return F.relu(out + idt)


def _layer(p, prefix, li, x, training, nb):
    x = _basic_block(p, f"{prefix}layer{li}.0.", x, 1 if li == 1 else 2, training, nb)
    return _basic_block(p, f"{prefix}layer{li}.1.", x, 1, training, nb)


def _stem(p, prefix, x, training, nb):
    x = F.conv2d(x, p[prefix + "conv1.weight"], None, stride=2, padding=3)
    x = F.relu(_bn_apply(p, prefix + "bn1", x, training, nb))
    return F.max_pool2d(x, 3, 2, 1)


def quadrants(x):
    """Level-1 split, order TL, TR, BL, BR; odd sizes give the second half the extra row/col
    (QS/models.py:277-282)."""
    h, w = x.shape[2], x.shape[3]
    return [x[:, :, : h // 2, : w // 2], x[:, :, : h // 2, w // 2:], x[:, :, h // 2:, : w // 2], x[:, :, h // 2:, w // 2:]]


def _dropout(x, p, training, rate):
    # parity runs use rate 0 (or eval); a non-zero rate here uses torch's generator, not the kernels' hash.
    return F.dropout(x, rate, training) if rate > 0 else x


# ------------------------------------------------------------------------------------------------
# Models
# ------------------------------------------------------------------------------------------------
def quadtree_forward(p: Params, image, numerical, training=True, mode="fusion", dropout_rate=0.0,
                     new_buffers: Optional[dict] = None, taps: Optional[dict] = None):
    """QuadtreeCNN.forward — QS/models.py:273-305; `mode` as in resnet/models.py:141-180."""
    pre = "base_cnn."
    image_features = numerical_features = None
    if mode in ("fusion", "image_only"):
        x = _stem(p, pre, image, training, new_buffers)
        for li in (1, 2, 3):
            x = _layer(p, pre, li, x, training, new_buffers)
        base = x  # [B,256,14,14]
        qf = []
        for q in quadrants(base):
            y = F.relu(F.conv2d(q, p["quadrant_processor.0.weight"], p["quadrant_processor.0.bias"], padding=1))
            qf.append(F.max_pool2d(y, 2, 2).flatten(1))
        l4 = _layer(p, pre, 4, base, training, new_buffers)
        glob = F.adaptive_avg_pool2d(l4, 1).flatten(1)
        image_features = torch.cat([glob] + qf, dim=1)
        if taps is not None:
            taps["base_features"] = base
            taps["layer4"] = l4
            taps["image_features"] = image_features
    if mode in ("fusion", "numerical_only"):
        h = F.relu(F.linear(numerical, p["numerical_mlp.0.weight"], p["numerical_mlp.0.bias"]))
        h = _dropout(h, p, training, dropout_rate)
        numerical_features = F.linear(h, p["numerical_mlp.3.weight"], p["numerical_mlp.3.bias"])
    if mode == "fusion":
        comb = torch.cat((image_features, numerical_features), dim=1)
    elif mode == "image_only":
        comb = image_features
    else:
        comb = numerical_features
    h = F.relu(F.linear(comb, p["classifier.0.weight"], p["classifier.0.bias"]))
    h = _dropout(h, p, training, dropout_rate)
    return F.linear(h, p["classifier.3.weight"], p["classifier.3.bias"])


def _level12(p, base):
    """Level-1 quadrant vectors (4 x 128) and level-2 sub-quadrant vectors (16 x 64), quadrant-major order
    (QS/models.py:60-79)."""
    quads = quadrants(base)
    qv = [F.adaptive_avg_pool2d(F.relu(F.conv2d(q, p["quadrant_processor.0.weight"], p["quadrant_processor.0.bias"],
                                                padding=1)), 1).flatten(1) for q in quads]
    sv = []
    for q in quads:
        for sq in quadrants(q):
            sv.append(F.adaptive_avg_pool2d(F.relu(F.conv2d(sq, p["sub_quadrant_processor.0.weight"],
                                                            p["sub_quadrant_processor.0.bias"], padding=1)), 1).flatten(1))
    return qv, sv


def _hier_trunk(p, image, training, nb):
    pre = "base_cnn."
    x = _stem(p, pre, image, training, nb)
    for li in (1, 2):
        x = _layer(p, pre, li, x, training, nb)
    base = x  # [B,128,28,28]
    g = _layer(p, pre, 4, _layer(p, pre, 3, base, training, nb), training, nb)
    return base, F.adaptive_avg_pool2d(g, 1).flatten(1)


def attention_hier_forward(p: Params, image, numerical, training=True, dropout_rate=0.0, new_buffers=None):
    """AttentionHierarchicalCNN.forward — QS/models.py:56-101."""
    base, glob = _hier_trunk(p, image, training, new_buffers)
    qv, sv = _level12(p, base)
    stacked = torch.stack(sv, dim=1)  # [B,16,64]
    s = F.linear(F.relu(F.linear(stacked, p["attention_gate.0.weight"], p["attention_gate.0.bias"])),
                 p["attention_gate.2.weight"], p["attention_gate.2.bias"]).squeeze(-1)
    wts = F.softmax(s, dim=1).unsqueeze(-1)
    attended = torch.sum(stacked * wts, dim=1)
    img = torch.cat([glob] + qv + [attended], dim=1)
    num = _dropout(F.relu(F.linear(numerical, p["numerical_mlp.0.weight"], p["numerical_mlp.0.bias"])), p, training,
                   dropout_rate)
    h = F.relu(F.linear(torch.cat((img, num), dim=1), p["classifier.0.weight"], p["classifier.0.bias"]))
    h = _dropout(h, p, training, dropout_rate)
    return F.linear(h, p["classifier.3.weight"], p["classifier.3.bias"])


def hier_forward(p: Params, image, numerical, training=True, dropout_rate=0.0, new_buffers=None):
    """HierarchicalQuadtreeCNN.forward with the evidently intended slicing (QS/models.py:167-210; the file's
    `w:` / `qw:` bottom-right slices are zero-width and crash — SURVEY.md §0.2 — so there is no reference
    behaviour to pin; the slicing of lines 64-67/75-78 is used)."""
    base, glob = _hier_trunk(p, image, training, new_buffers)
    qv, sv = _level12(p, base)
    img = torch.cat([glob] + qv + sv, dim=1)
    num = _dropout(F.relu(F.linear(numerical, p["numerical_mlp.0.weight"], p["numerical_mlp.0.bias"])), p, training,
                   dropout_rate)
    h = F.relu(F.linear(torch.cat((img, num), dim=1), p["classifier.0.weight"], p["classifier.0.bias"]))
    h = _dropout(h, p, training, dropout_rate)
    return F.linear(h, p["classifier.3.weight"], p["classifier.3.bias"])


def standard_resnet_forward(p: Params, image, numerical=None, training=True, dropout_rate=0.0, new_buffers=None):
    """StandardResNetCNN.forward — resnet/models.py:56-65."""
    pre = "base_cnn."
    x = _stem(p, pre, image, training, new_buffers)
    for li in (1, 2, 3, 4):
        x = _layer(p, pre, li, x, training, new_buffers)
    f = F.adaptive_avg_pool2d(x, 1).flatten(1)
    h = _dropout(F.relu(F.linear(f, p["classifier.0.weight"], p["classifier.0.bias"])), p, training, dropout_rate)
    return F.linear(h, p["classifier.3.weight"], p["classifier.3.bias"])


def _lstm_last(p: Params, x, name="numerical_lstm", layers=2):
    """nn.LSTM(batch_first=True) last time step, gate order i,f,g,o (inter-layer dropout off: parity runs use
    rate 0 / eval). 3dcnn/models.py:144-150, 200-202."""
    seq = x
    for layer in range(layers):
        wih, whh = p[f"{name}.weight_ih_l{layer}"], p[f"{name}.weight_hh_l{layer}"]
        bih, bhh = p[f"{name}.bias_ih_l{layer}"], p[f"{name}.bias_hh_l{layer}"]
        hid = whh.shape[1]
        h = x.new_zeros(x.shape[0], hid)
        c = x.new_zeros(x.shape[0], hid)
        outs = []
        for t in range(seq.shape[1]):
            gates = F.linear(seq[:, t], wih, bih) + F.linear(h, whh, bhh)
            i, f, g, o = gates.chunk(4, dim=1)
            c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
            h = torch.sigmoid(o) * torch.tanh(c)
            outs.append(h)
        seq = torch.stack(outs, dim=1)
    return seq[:, -1]


_POOLS_3D = {"conv3d_block1": (1, 2, 2), "conv3d_block2": (2, 2, 2), "conv3d_block3": (2, 2, 2),
             "conv3d_block4_new": (1, 2, 2), "conv3d_final_features": None}


def quadtree3d_conv_stack(p: Params, clips, training=True, new_buffers=None):
    """Conv3d stack of Quadtree3DCNN — 3dcnn/models.py:107-141, 189-198. clips: [B,T,3,H,W]."""
    x = clips.permute(0, 2, 1, 3, 4)
    for nm, pool in _POOLS_3D.items():
        x = F.conv3d(x, p[nm + ".0.weight"], p[nm + ".0.bias"], padding=1)
        x = F.relu(_bn_apply(p, nm + ".1", x, training, new_buffers))
        if pool is not None:
            x = F.max_pool3d(x, pool, pool)
    return F.adaptive_avg_pool3d(x, 1).flatten(1)


def quadtree3d_forward(p: Params, clips, numerical_seq, training=True, mode="quadtree_3d_fusion", dropout_rate=0.0,
                       new_buffers=None):
    """Quadtree3DCNN.forward — 3dcnn/models.py:184-214."""
    img = quadtree3d_conv_stack(p, clips, training, new_buffers)
    if mode == "quadtree_3d_fusion":
        last = _lstm_last(p, numerical_seq)
        num = _dropout(F.relu(F.linear(last, p["numerical_projection.0.weight"], p["numerical_projection.0.bias"])), p,
                       training, dropout_rate)
        comb = torch.cat((img, num), dim=1)
    else:
        comb = img
    h = _dropout(F.relu(F.linear(comb, p["classifier.0.weight"], p["classifier.0.bias"])), p, training, dropout_rate)
    return F.linear(h, p["classifier.3.weight"], p["classifier.3.bias"])


def cnn_lstm_forward(p: Params, image_sequence, numerical_sequence, training=True, dropout_rate=0.0, new_buffers=None):
    """CnnLstm.forward — cnn+lstm/models.py:59-89: every frame through the (frozen, BatchNorm in train mode) ResNet-18
    up to avgpool, per-step MLP on the pose vector, concat 640 -> 2-layer LSTM -> last step -> classifier."""
    b, t, c, h, w = image_sequence.shape
    pre = "cnn_backbone."
    q = dict(p)  # torchvision-style aliases of the Sequential-indexed names, so the ResNet helpers apply
    for name, idx in _SEQ_INDEX.items():
        for k, v in p.items():
            if k.startswith(f"{pre}{idx}."):
                q[f"{pre}{name}.{k[len(pre) + len(str(idx)) + 1:]}"] = v
    nb_alias: dict = {}
    x = _stem(q, pre, image_sequence.reshape(b * t, c, h, w), training, nb_alias)
    for li in (1, 2, 3, 4):
        x = _layer(q, pre, li, x, training, nb_alias)
    if new_buffers is not None:
        for k, v in nb_alias.items():  # back to the module's own buffer names
            head, rest = k[len(pre):].split(".", 1)
            new_buffers[f"{pre}{_SEQ_INDEX[head]}.{rest}"] = v
    c_out = F.adaptive_avg_pool2d(x, 1).flatten(1).view(b, t, -1)
    n_out = F.linear(F.relu(F.linear(numerical_sequence, p["numerical_mlp.0.weight"], p["numerical_mlp.0.bias"])),
                     p["numerical_mlp.2.weight"], p["numerical_mlp.2.bias"])
    last = _lstm_last(p, torch.cat((c_out, n_out), dim=2), name="lstm")
    hdn = _dropout(F.relu(F.linear(last, p["classifier.0.weight"], p["classifier.0.bias"])), p, training, dropout_rate)
    return F.linear(hdn, p["classifier.3.weight"], p["classifier.3.bias"])


def _r3d_features(p: Params, pre: str, hybrid: bool, clips, training, nb):
    """torchvision r3d_18 up to layer4 (video/resnet.py: BasicStem, BasicBlock with Conv3DSimple) on [B,T,3,H,W] clips."""
    x = clips.permute(0, 2, 1, 3, 4)
    stem = pre + ("0." if hybrid else "stem.")
    x = F.conv3d(x, p[stem + "0.weight"], None, stride=(1, 2, 2), padding=(1, 3, 3))
    x = F.relu(_bn_apply(p, stem + "1", x, training, nb))
    for li in (1, 2, 3, 4):
        for bi in (0, 1):
            b = f"{pre}{li}.{bi}." if hybrid else f"{pre}layer{li}.{bi}."
            stride = 2 if (li > 1 and bi == 0) else 1
            out = F.conv3d(x, p[b + "conv1.0.weight"], None, stride=stride, padding=1)
            out = F.relu(_bn_apply(p, b + "conv1.1", out, training, nb))
            out = F.conv3d(out, p[b + "conv2.0.weight"], None, stride=1, padding=1)
            out = _bn_apply(p, b + "conv2.1", out, training, nb)
            if b + "downsample.0.weight" in p:
                idt = F.conv3d(x, p[b + "downsample.0.weight"], None, stride=stride)
                idt = _bn_apply(p, b + "downsample.1", idt, training, nb)
            else:
                idt = x
            x = F.relu(out + idt)
    return F.adaptive_avg_pool3d(x, 1).flatten(1)


def resnet3d_video_forward(p: Params, clips, numerical_seq=None, training=True, dropout_rate=0.0, new_buffers=None):
    """ResNet3DVideo.forward — 3dcnn/models.py:255-262 (r3d_18 with the 512 -> 256 -> nc head as `fc`)."""
    f = _r3d_features(p, "r3d_model.", False, clips, training, new_buffers)
    h = _dropout(F.relu(F.linear(f, p["r3d_model.fc.0.weight"], p["r3d_model.fc.0.bias"])), p, training, dropout_rate)
    return F.linear(h, p["r3d_model.fc.3.weight"], p["r3d_model.fc.3.bias"])


def hybrid3d_forward(p: Params, clips, numerical_seq, training=True, mode="hybrid_quadtree_3d_fusion", dropout_rate=0.0,
                     new_buffers=None):
    """HybridQuadtree3DCNN.forward — 3dcnn/models.py:335-372."""
    img = _r3d_features(p, "pretrained_image_extractor.", True, clips, training, new_buffers)
    if mode == "hybrid_quadtree_3d_fusion":
        last = _lstm_last(p, numerical_seq)
        num = _dropout(F.relu(F.linear(last, p["numerical_projection.0.weight"], p["numerical_projection.0.bias"])), p,
                       training, dropout_rate)
        comb = torch.cat((img, num), dim=1)
    else:
        comb = img
    h = _dropout(F.relu(F.linear(comb, p["classifier.0.weight"], p["classifier.0.bias"])), p, training, dropout_rate)
    return F.linear(h, p["classifier.3.weight"], p["classifier.3.bias"])


FORWARDS = {
    "resnet3d_video": resnet3d_video_forward,
    "hybrid3d": hybrid3d_forward,
    "cnn_lstm": cnn_lstm_forward,
    "quadtree": quadtree_forward,
    "attention_hierarchical": attention_hier_forward,
    "hierarchical_quadtree": hier_forward,
    "standard_resnet": standard_resnet_forward,
    "quadtree3d": quadtree3d_forward,
}


def loss_and_grads(kind: str, p: Params, inputs, labels, training=True, **kw):
    """CrossEntropyLoss (mean) + autograd, like the training scripts' hot loop (QS/Quadtree_train.py:63-65).
    Returns logits, loss, {name: grad} for float parameters that received one, and the updated BN buffers."""
    leaves = {k: v.clone().requires_grad_(True) for k, v in p.items() if v.is_floating_point() and "running_" not in k}
    full = dict(p)
    full.update(leaves)
    nb: dict = {}
    logits = FORWARDS[kind](full, *inputs, training=training, new_buffers=nb, **kw)
    loss = F.cross_entropy(logits, labels)
    names = list(leaves)
    grads = torch.autograd.grad(loss, [leaves[k] for k in names], allow_unused=True)
    return logits.detach(), loss.detach(), {k: g for k, g in zip(names, grads) if g is not None}, nb


def train_step_cpu(kind: str, p: Params, inputs, labels, lr=1e-4, weight_decay=1e-4, state=None, **kw):
    """One fwd+bwd+Adam step (torch.optim.Adam L2-in-grad semantics, QS/Quadtree_train.py:45) — used as the
    timed CPU baseline."""
    logits, loss, grads, nb = loss_and_grads(kind, p, inputs, labels, training=True, **kw)
    state = {} if state is None else state
    t = state.get("step", 0) + 1
    state["step"] = t
    b1, b2, eps = 0.9, 0.999, 1e-8
    with torch.no_grad():
        for k, g in grads.items():
            g = g + weight_decay * p[k]
            m = state.setdefault("m." + k, torch.zeros_like(g))
            v = state.setdefault("v." + k, torch.zeros_like(g))
            m.mul_(b1).add_(g, alpha=1 - b1)
            v.mul_(b2).addcmul_(g, g, value=1 - b2)
            denom = (v.sqrt() / math.sqrt(1 - b2 ** t)).add_(eps)
            p[k] = p[k] - (lr / (1 - b1 ** t)) * m / denom
        p.update(nb)
    return float(loss), state
