"""Synthetic inputs of the benchmark workloads and the device-side input pipeline of the drop-in models.

`synthetic_batch` is the SURVEY.md §8(d) recipe (the same draws as the oracle's copy — `tests/test_data_cpu.py` checks
that they are identical): N(0,1) images (ImageNet-normalised pixels, `Quadtree_from scratch/dataloader.py:35-36`), a
47-float pose vector with the real, un-standardised feature ranges (`img process/1_prepare_still_image_dataset.py:101-113`:
33 visibilities in [0,1], 10 joint angles in degrees, 3 normalised distances, 1 ratio) and integer labels.

`BatchPrefetcher` is the loader side of `images.to(device)` in the reference loop (`Quadtree_train.py:61`): pinned host
batches are copied on a side stream into pre-allocated device staging buffers (double buffered), so the copy of step
i+1 overlaps the kernels of step i. uint8 images (what a JPEG decoder produces) travel as bytes — a quarter of the fp32
PCIe traffic — and are normalised on the device inside the stem's packing kernel (`models.ImageBatchU8`).
"""
from __future__ import annotations

from typing import Iterable, Iterator, Optional, Sequence, Tuple

import torch

IMAGENET_MEAN = (0.485, 0.456, 0.406)  # Quadtree_from scratch/dataloader.py:36
IMAGENET_STD = (0.229, 0.224, 0.225)


def synthetic_batch(batch: int, seed: int = 1234, image_size: int = 224, num_classes: int = 8, seq_len: int = 0,
                    clip_size: int = 112):
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


def quantize_images_u8(images: torch.Tensor, mean: Sequence[float] = IMAGENET_MEAN, std: Sequence[float] = IMAGENET_STD) -> torch.Tensor:
    """Inverse of ToTensor+Normalize: normalised fp32 [.., 3, H, W] -> uint8 pixels (what the JPEG decoder handed the
    reference's transform). Used to build byte-valued synthetic batches for the uint8 input path."""
    shape = [1] * images.dim()
    shape[-3] = 3
    m = torch.tensor(mean, dtype=torch.float32).view(shape)
    s = torch.tensor(std, dtype=torch.float32).view(shape)
    return ((images * s + m) * 255.0).round().clamp_(0, 255).to(torch.uint8)


def normalize_u8_reference(images_u8: torch.Tensor, mean: Sequence[float] = IMAGENET_MEAN, std: Sequence[float] = IMAGENET_STD) -> torch.Tensor:
    """ToTensor + Normalize of the reference transform (dataloader.py:35-36) in fp32: (u8/255 - mean) / std."""
    shape = [1] * images_u8.dim()
    shape[-3] = 3
    m = torch.tensor(mean, dtype=torch.float32, device=images_u8.device).view(shape)
    s = torch.tensor(std, dtype=torch.float32, device=images_u8.device).view(shape)
    return (images_u8.float() / 255.0 - m) / s


class BatchPrefetcher:
    """Iterate over host batches (tuples of CPU tensors, ideally pinned) and yield device tuples whose H2D copies were
    issued one step ahead on a copy stream into reusable staging buffers.

        for images, numerical, labels in BatchPrefetcher(loader, device):
            loss = criterion(model(images, numerical), labels) ...

    The yielded tensors stay valid until the next-but-one batch is requested (two staging slots). The compute stream
    waits on the copy's event; the copy stream waits on the event recorded when the slot's previous consumer step was
    queued, so no `record_stream` bookkeeping and no allocator traffic happen per step."""

    def __init__(self, batches: Iterable[Tuple[torch.Tensor, ...]], device, slots: int = 2):
        self.batches = batches
        self.device = torch.device(device)
        self.slots = max(2, int(slots))
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self._staging = [None] * self.slots
        self._ready = [None] * self.slots
        self._released = [None] * self.slots
        self.h2d_bytes_last = 0

    def _issue(self, slot: int, host: Tuple[torch.Tensor, ...]):
        st = self._staging[slot]
        if st is None or len(st) != len(host) or any(a.shape != b.shape or a.dtype != b.dtype for a, b in zip(st, host)):
            st = tuple(torch.empty(t.shape, dtype=t.dtype, device=self.device) for t in host)
            self._staging[slot] = st
        with torch.cuda.stream(self.copy_stream):
            if self._released[slot] is not None:
                self.copy_stream.wait_event(self._released[slot])  # the step that last read this slot has been queued
            for dst, src in zip(st, host):
                dst.copy_(src, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        self._ready[slot] = ev
        self.h2d_bytes_last = sum(t.numel() * t.element_size() for t in host)

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, ...]]:
        it = iter(self.batches)
        i = 0
        try:
            self._issue(0, next(it))
        except StopIteration:
            return
        while True:
            slot = i % self.slots
            nxt: Optional[Tuple[torch.Tensor, ...]]
            try:
                nxt = next(it)
            except StopIteration:
                nxt = None
            if nxt is not None:
                self._issue((i + 1) % self.slots, nxt)  # travels while step i computes
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(self._ready[slot])
            yield self._staging[slot]
            rel = torch.cuda.Event()
            rel.record(cur)  # everything the consumer queued for this batch precedes the slot's next overwrite
            self._released[slot] = rel
            if nxt is None:
                return
            i += 1
