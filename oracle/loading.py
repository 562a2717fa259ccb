"""TEST INFRASTRUCTURE: load an oracle parameter dict (oracle/quadtree_oracle.make_params naming) into one of the product's
drop-in modules. Lives beside the oracle because only tests, smoke() and the parity legs of bench.py need it."""
import torch.nn as nn


def load_oracle_params(model: nn.Module, params: dict) -> None:
    """Load a parameter dict in the oracle's naming (oracle/quadtree_oracle.make_params). For QuadtreeCNN /
    StandardResNetCNN the names ARE the reference's state_dict keys; the hierarchical classes hold the ResNet
    only through features_extractor / global_processor, so `base_cnn.*` names are mapped onto those."""
    own = model.state_dict()
    alias = {}
    if not hasattr(model, "base_cnn") and any(k.startswith("base_cnn.") for k in params):
        fe = {"conv1": "features_extractor.0", "bn1": "features_extractor.1", "layer1": "features_extractor.4",
              "layer2": "features_extractor.5", "layer3": "global_processor.0", "layer4": "global_processor.1"}
        for k, v in params.items():
            if k.startswith("base_cnn."):
                rest = k[len("base_cnn."):]
                head = rest.split(".")[0]
                if head in fe:
                    alias[fe[head] + rest[len(head):]] = v
            else:
                alias[k] = v
    else:
        alias = params
    sd = {k: v for k, v in alias.items() if k in own}
    res = model.load_state_dict(sd, strict=False)
    missing = [k for k in res.missing_keys if not k.startswith(("features_extractor.", "global_processor."))]
    if missing and hasattr(model, "base_cnn"):
        raise RuntimeError(f"load_oracle_params: missing {missing[:5]}")
    if not hasattr(model, "base_cnn"):
        really_missing = [k for k in own if k not in sd]
        if really_missing:
            raise RuntimeError(f"load_oracle_params: missing {really_missing[:5]}")
