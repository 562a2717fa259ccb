"""Generate tests/golden/*.json by running the UNMODIFIED reference modules (needs /root/reference; run in the
build container only:  python tests/golden/make_golden.py).

For each case the seeded parameters of oracle/quadtree_oracle.make_params are loaded into the reference
nn.Module (`torchvision.models.resnet18` is patched to skip the ImageNet download — there is no network), the
reference runs forward (+ CrossEntropyLoss backward) on the seeded synthetic batch, and a compact digest is
stored: logits, loss, per-parameter gradient norm + 4 sampled elements, and BN running-stat digests.
Fixtures are tiny; weights and inputs are regenerated from seeds by the tests.
"""
import importlib.util
import json
import os
import sys

import torch
import torchvision

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import quadtree_oracle as O  # noqa: E402

REF = "/root/reference"
_r18 = torchvision.models.resnet18
torchvision.models.resnet18 = lambda weights=None, **kw: _r18(weights=None, **kw)
import torchvision.models.video as _video  # noqa: E402
_r3d = _video.r3d_18
_video.r3d_18 = lambda weights=None, **kw: _r3d(weights=None, **kw)  # KINETICS400 download is impossible offline too


def load_ref(relpath, name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, relpath))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def sample_idx(numel):
    return sorted({0, numel // 3, (2 * numel) // 3, numel - 1})


def digest(t):
    f = t.detach().double().flatten()
    return {"norm": float(f.norm()), "idx": sample_idx(f.numel()), "val": [float(f[i]) for i in sample_idx(f.numel())]}


def to_ref_state(kind, p, model):
    """Map oracle parameter names onto the reference module's state_dict keys."""
    sd = model.state_dict()
    out = {}
    alias = {}
    if kind in ("attention_hierarchical", "hierarchical_quadtree"):
        # these classes keep the ResNet only through features_extractor / global_processor (QS/models.py:13-20)
        fe = {"conv1": "features_extractor.0", "bn1": "features_extractor.1", "layer1": "features_extractor.4",
              "layer2": "features_extractor.5", "layer3": "global_processor.0", "layer4": "global_processor.1"}
        for k, v in p.items():
            if k.startswith("base_cnn."):
                rest = k[len("base_cnn."):]
                head = rest.split(".")[0]
                if head in fe:
                    alias[fe[head] + rest[len(head):]] = v
            else:
                alias[k] = v
    else:
        alias = dict(p)
    for k, v in alias.items():
        if k in sd:
            assert sd[k].shape == v.shape, (k, sd[k].shape, v.shape)
            out[k] = v
    missing = [k for k in sd if k not in out and not any(k.startswith(a) for a in ("features_extractor.", "global_processor."))]
    assert not missing or kind in ("attention_hierarchical", "hierarchical_quadtree"), missing[:5]
    return out


def ref_grad_names(kind, name):
    """reference parameter name -> oracle parameter name"""
    if kind in ("attention_hierarchical", "hierarchical_quadtree"):
        back = {"features_extractor.0": "conv1", "features_extractor.1": "bn1", "features_extractor.4": "layer1",
                "features_extractor.5": "layer2", "global_processor.0": "layer3", "global_processor.1": "layer4"}
        for a, b in back.items():
            if name.startswith(a + "."):
                return "base_cnn." + b + name[len(a):]
    return name


def run_case(case):
    kind, batch, seed = case["kind"], case["batch"], case["seed"]
    mode = case.get("mode", "fusion")
    training = case.get("training", True)
    torch.manual_seed(0)
    if kind == "quadtree" and case.get("file") == "resnet":
        mod = load_ref("resnet/models.py", "ref_resnet")
        model = mod.QuadtreeCNN(num_classes=8, dropout_rate=0.0, mode=mode)
    elif kind == "quadtree":
        mod = load_ref("Quadtree_from scratch/models.py", "ref_qs")
        model = mod.QuadtreeCNN(num_classes=8, dropout_rate=0.0)
    elif kind == "attention_hierarchical":
        mod = load_ref("Quadtree_from scratch/models.py", "ref_qs")
        model = mod.AttentionHierarchicalCNN(num_classes=8, dropout_rate=0.0)
    elif kind == "standard_resnet":
        mod = load_ref("resnet/models.py", "ref_resnet")
        model = mod.StandardResNetCNN(num_classes=8, dropout_rate=0.0)
    elif kind == "quadtree3d":
        mod = load_ref("3dcnn/models.py", "ref_3d")
        model = mod.Quadtree3DCNN(num_classes=8, sequence_length=case["seq_len"], dropout_rate=0.0, mode=mode)
    elif kind == "resnet3d_video":
        mod = load_ref("3dcnn/models.py", "ref_3d")
        model = mod.ResNet3DVideo(num_classes=8, dropout_rate=0.0)
    elif kind == "hybrid3d":
        mod = load_ref("3dcnn/models.py", "ref_3d")
        model = mod.HybridQuadtree3DCNN(num_classes=8, sequence_length=case["seq_len"], dropout_rate=0.0, mode=mode)
    elif kind == "cnn_lstm":
        mod = load_ref("cnn+lstm/models.py", "ref_cnnlstm")
        model = mod.CnnLstm(num_classes=8, sequence_length=case["seq_len"], dropout_rate=0.0)
    else:
        raise ValueError(kind)
    p = O.make_params(kind, 8, seed=case["param_seed"], mode=mode)
    sd = to_ref_state(kind, p, model)
    res = model.load_state_dict(sd, strict=False)
    own_missing = [k for k in res.missing_keys if not k.startswith(("features_extractor.", "global_processor."))]
    assert not own_missing and not res.unexpected_keys, (own_missing[:5], res.unexpected_keys[:5])
    if kind in ("quadtree3d", "cnn_lstm", "resnet3d_video", "hybrid3d"):
        images, numerical, labels = O.synthetic_batch(batch, seed, seq_len=case["seq_len"], clip_size=case["clip"])
    else:
        images, numerical, labels = O.synthetic_batch(batch, seed)
    model.train(training)
    out = {"case": case}
    if training:
        logits = model(images, numerical)
        loss = torch.nn.functional.cross_entropy(logits, labels)
        loss.backward()
        out["loss"] = float(loss)
        grads = {}
        seen = set()
        for name, prm in model.named_parameters():
            if prm.grad is None or id(prm) in seen:
                continue
            seen.add(id(prm))
            grads[ref_grad_names(kind, name)] = digest(prm.grad)
        out["grads"] = grads
        bufs = {}
        for name, b in model.named_buffers():
            on = ref_grad_names(kind, name)
            if "running_" in name and on not in bufs:
                bufs[on] = digest(b)
        out["buffers"] = bufs
    else:
        with torch.no_grad():
            logits = model(images, numerical)
    out["logits"] = logits.detach().double().tolist()
    return out


CASES = [
    {"name": "quadtree_train_b2", "kind": "quadtree", "batch": 2, "seed": 1234, "param_seed": 0},
    {"name": "quadtree_eval_b2", "kind": "quadtree", "batch": 2, "seed": 4321, "param_seed": 0, "training": False},
    {"name": "quadtree_frozen_fusion_b2", "kind": "quadtree", "file": "resnet", "mode": "fusion", "batch": 2,
     "seed": 77, "param_seed": 3},
    {"name": "quadtree_image_only_b2", "kind": "quadtree", "file": "resnet", "mode": "image_only", "batch": 2,
     "seed": 78, "param_seed": 4},
    {"name": "quadtree_numerical_only_b3", "kind": "quadtree", "file": "resnet", "mode": "numerical_only", "batch": 3,
     "seed": 79, "param_seed": 5},
    {"name": "attention_hier_train_b2", "kind": "attention_hierarchical", "batch": 2, "seed": 99, "param_seed": 1},
    {"name": "attention_hier_eval_b2", "kind": "attention_hierarchical", "batch": 2, "seed": 98, "param_seed": 1,
     "training": False},
    {"name": "standard_resnet_train_b2", "kind": "standard_resnet", "batch": 2, "seed": 55, "param_seed": 2},
    {"name": "quadtree3d_fusion_b2", "kind": "quadtree3d", "mode": "quadtree_3d_fusion", "batch": 2, "seed": 11,
     "param_seed": 6, "seq_len": 4, "clip": 32},
    {"name": "quadtree3d_image_only_b2", "kind": "quadtree3d", "mode": "quadtree_3d_image_only", "batch": 2, "seed": 12,
     "param_seed": 7, "seq_len": 4, "clip": 32},
    {"name": "cnn_lstm_train_b2", "kind": "cnn_lstm", "batch": 2, "seed": 21, "param_seed": 8, "seq_len": 3, "clip": 64},
    {"name": "resnet3d_video_b2", "kind": "resnet3d_video", "batch": 2, "seed": 31, "param_seed": 9, "seq_len": 8, "clip": 64},
    {"name": "hybrid3d_fusion_b2", "kind": "hybrid3d", "mode": "hybrid_quadtree_3d_fusion", "batch": 2, "seed": 32,
     "param_seed": 10, "seq_len": 8, "clip": 64},
]


if __name__ == "__main__":
    torch.set_num_threads(8)
    only = set(sys.argv[1:])
    for case in CASES:
        if only and case["name"] not in only:
            continue
        res = run_case(case)
        path = os.path.join(HERE, case["name"] + ".json")
        with open(path, "w") as f:
            json.dump(res, f)
        print("wrote", path, "loss", res.get("loss"), "logit0", res["logits"][0][:3])
