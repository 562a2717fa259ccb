"""The product's synthetic-input recipe (qtcnn_b200.data) is the oracle's, draw for draw, and the uint8 image helpers
invert / reproduce the reference transform (ToTensor + Normalize, Quadtree_from scratch/dataloader.py:35-36)."""
import torch

from oracle import quadtree_oracle as O
from qtcnn_b200 import data


def test_synthetic_batch_equals_oracle_recipe():
    for kw in (dict(batch=3, seed=5), dict(batch=2, seed=9, seq_len=4, clip_size=16)):
        a = data.synthetic_batch(**kw)
        b = O.synthetic_batch(**kw)
        for x, y in zip(a, b):
            assert x.dtype == y.dtype and torch.equal(x, y)
    images, numerical, labels = data.synthetic_batch(4, 1)
    assert images.shape == (4, 3, 224, 224) and numerical.shape == (4, 47) and labels.dtype == torch.int64
    assert float(numerical[:, 33:43].max()) > 10.0  # un-standardised joint angles in degrees


def test_u8_quantise_normalise_roundtrip():
    images, _, _ = data.synthetic_batch(2, 3)
    u8 = data.quantize_images_u8(images)
    assert u8.dtype == torch.uint8 and u8.shape == images.shape
    back = data.normalize_u8_reference(u8)
    inside = (images * torch.tensor(data.IMAGENET_STD).view(1, 3, 1, 1) + torch.tensor(data.IMAGENET_MEAN).view(1, 3, 1, 1))
    ok = (inside > 0.01) & (inside < 0.99)  # values the byte range can represent
    step = (1.0 / 255.0) / torch.tensor(data.IMAGENET_STD).view(1, 3, 1, 1)
    assert bool(((back - images).abs() <= 0.5 * step + 1e-6)[ok].all())
