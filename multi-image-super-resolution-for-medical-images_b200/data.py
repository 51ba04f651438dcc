"""Seeded on-device synthetic triplet generator (replaces the DICOM pipeline, which needs a dataset that is not
available: reference src/ModelDataGenerator.py). Produces what `build_dataloader` yields after collation:
inputs (B,2,H,W) = [prior, posterior] and target (B,1,H,W) = middle slice, fp32, each slice z-scored as in
ModelDataGenerator.py:73-75 `(x - mean) / (std + 1e-6)`.

Volumes are smooth random fields (low-resolution Gaussian noise upsampled bilinearly + a little iid noise) with a
linear drift across 5 slices, so neighbouring slices are correlated like real MRI slices; `distance` 2 gives
(i, i+2 -> i+1) triplets, 4 gives (i, i+4 -> i+2). This is data generation, not the hot path: plain torch ops.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


class SyntheticTripletGenerator:
    def __init__(self, batch_size, height=256, width=256, device="cuda", seed=1234, rank=0, distance=2,
                 coarse=8, noise=0.1):
        if distance not in (2, 4):
            raise ValueError("distance must be 2 or 4")
        self.B, self.H, self.W = batch_size, height, width
        self.device = torch.device(device)
        self.distance, self.coarse, self.noise = distance, coarse, noise
        self.gen = torch.Generator(device=self.device)
        self.gen.manual_seed(seed + rank)

    def _zscore(self, s):
        mean = s.mean(dim=(-2, -1), keepdim=True)
        std = s.std(dim=(-2, -1), keepdim=True)
        return (s - mean) / (std + 1e-6)

    def next(self):
        B, H, W, c = self.B, self.H, self.W, self.coarse
        dev, g = self.device, self.gen
        base = torch.randn(B, 1, c, c, device=dev, generator=g)
        drift = torch.randn(B, 1, c, c, device=dev, generator=g)
        t = torch.linspace(-1.0, 1.0, 5, device=dev).view(1, 5, 1, 1)
        coarse_vol = base + 0.5 * t * drift                          # (B,5,c,c)
        vol = F.interpolate(coarse_vol, size=(H, W), mode="bilinear", align_corners=False)
        vol = vol + self.noise * torch.randn(B, 5, H, W, device=dev, generator=g)
        vol = self._zscore(vol)
        if self.distance == 2:
            i, m, j = 1, 2, 3
        else:
            i, m, j = 0, 2, 4
        inputs = torch.stack([vol[:, i], vol[:, j]], dim=1).contiguous()
        target = vol[:, m:m + 1].contiguous()
        return inputs, target

    def __iter__(self):
        while True:
            yield self.next()


def unpack_batch(batch):
    """Accept both loader contracts of the reference: (inputs, targets) (unet_model.py:174) and
    ((pre, post), mid) (ModelDataGenerator.py:214)."""
    first, targets = batch
    if isinstance(first, (tuple, list)):
        pre, post = first
        first = torch.cat([pre, post], dim=1)
    return first, targets


class DevicePrefetcher:
    """Wraps a host loader: batch i+1 is copied host->device on a side stream while batch i is being consumed, so the
    H2D copy (25 MB per B=32 step) overlaps the train step instead of preceding it. Yields (inputs, targets) on
    `device`. Host tensors should be pinned (the reference's DataLoader uses pin_memory=True,
    ModelDataGenerator.py:276-282); unpinned tensors still work but copy synchronously."""

    def __init__(self, loader, device):
        self.loader, self.device = loader, torch.device(device)

    def __iter__(self):
        if self.device.type != "cuda":
            for batch in self.loader:
                x, y = unpack_batch(batch)
                yield x.to(self.device), y.to(self.device)
            return
        copy_stream = torch.cuda.Stream(device=self.device)
        it = iter(self.loader)
        # two persistent device slots (no allocator traffic in the loop): slot s is refilled on the copy stream only
        # after the consumer's work on it has been enqueued (event `done[s]`)
        bufs, ready, done = [None, None], [torch.cuda.Event(), torch.cuda.Event()], [None, None]

        def fetch(slot):
            batch = next(it, None)
            if batch is None:
                return False
            x, y = unpack_batch(batch)
            b = bufs[slot]
            if b is None or b[0].shape != x.shape or b[1].shape != y.shape or b[0].dtype != x.dtype:
                b = bufs[slot] = (torch.empty(x.shape, dtype=x.dtype, device=self.device),
                                  torch.empty(y.shape, dtype=y.dtype, device=self.device))
                copy_stream.wait_stream(torch.cuda.current_stream(self.device))
            if done[slot] is not None:
                copy_stream.wait_event(done[slot])
            with torch.cuda.stream(copy_stream):
                b[0].copy_(x, non_blocking=True)
                b[1].copy_(y, non_blocking=True)
            ready[slot].record(copy_stream)
            return True

        i = 0
        have = fetch(0)
        while have:
            nxt = fetch((i + 1) % 2)
            stream = torch.cuda.current_stream(self.device)
            stream.wait_event(ready[i % 2])
            yield bufs[i % 2]
            ev = done[i % 2] or torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))  # the consumer's use of this slot is enqueued
            done[i % 2] = ev
            i += 1
            have = nxt
