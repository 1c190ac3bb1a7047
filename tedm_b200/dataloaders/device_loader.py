"""uint8 batches -> pinned host memory -> copy stream -> fp32 tensors on the GPU.

The reference's loaders (dataloaders/JSRT.py:36-45, dataloaders/CXR14.py:35-45) hand fp32 CPU tensors to the
training loop, which then calls `x.to(device)` synchronously (trainers/train_baseline.py:32-33).  Once the
path runs at > 1000 images/s that copy and the fp32 expansion on the host are on the critical path, so:
  * datasets return uint8 planes (a quarter of the bytes over PCIe);
  * the H2D copy of batch i+1 is issued on a side stream while batch i is being consumed;
  * `ToTensor` (u8 / 255) and the label rule ((mask > .5) summed, clipped) run on the device, bit-exactly.
"""
from __future__ import annotations

import csv
from typing import Dict, Iterator, List, Optional, Sequence

import numpy as np
import torch
from torch.utils.data import DataLoader, Dataset, DistributedSampler

from .. import native as N


def read_csv_columns(path, columns: Sequence[str]) -> Dict[str, List[str]]:
    out: Dict[str, List[str]] = {c: [] for c in columns}
    with open(path, newline="") as f:
        for row in csv.DictReader(f):
            for c in columns:
                out[c].append(row[c])
    return out


class DeviceLoader:
    """Iterates like the reference's DataLoader but yields CUDA fp32 tensors: `x` (B, 1, S, S) in [0, 1] for
    unlabelled datasets, `(x, y)` with y (B, 1, S, S) in {0, 1} for labelled ones."""

    def __init__(self, dataset: Dataset, batch_size: int, shuffle: bool, num_workers: int, device="cuda",
                 labelled: bool = False, rank: int = 0, world_size: int = 1, seed: int = 0):
        self.dataset = dataset
        self.labelled = labelled
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("tedm_b200 loaders feed the CUDA path only; there is no CPU fallback")
        self.sampler = (DistributedSampler(dataset, num_replicas=world_size, rank=rank, shuffle=shuffle, seed=seed)
                        if world_size > 1 else None)
        self.loader = DataLoader(dataset, batch_size=batch_size, shuffle=shuffle and self.sampler is None,
                                 sampler=self.sampler, num_workers=num_workers, pin_memory=True)
        self._copy_stream: Optional[torch.cuda.Stream] = None
        self._epoch = 0
        self.h2d_bytes = 0

    def __len__(self) -> int:
        return len(self.loader)

    def _upload(self, batch):
        """Issue the H2D copies of one uint8 batch on the copy stream; returns (device uint8 tensors, event)."""
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        parts = batch if isinstance(batch, (list, tuple)) else (batch,)
        for p in parts:
            if p.dtype != torch.uint8:
                raise TypeError(f"datasets behind DeviceLoader return uint8 planes, got {p.dtype}")
        with torch.cuda.stream(self._copy_stream):
            dev = [p.contiguous().to(self.device, non_blocking=True) for p in parts]
            ev = torch.cuda.Event()
            ev.record(self._copy_stream)
        self.h2d_bytes += sum(p.numel() for p in parts)
        return dev, ev

    def _finish(self, staged):
        dev, ev = staged
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ev)
        for d in dev:
            d.record_stream(cur)
        x = N.u8_to_unit(dev[0])
        if not self.labelled:
            return x
        return x, N.u8_masks_to_label(dev[1])

    def __iter__(self) -> Iterator:
        if self.sampler is not None:
            self.sampler.set_epoch(self._epoch)
        self._epoch += 1
        staged = None
        for batch in self.loader:
            nxt = self._upload(batch)
            if staged is not None:
                yield self._finish(staged)
            staged = nxt
        if staged is not None:
            yield self._finish(staged)


class SyntheticXray(Dataset):
    """Deterministic stand-in for the chest-X-ray datasets (there is no dataset on the build or GPU boxes):
    smooth low-frequency uint8 images and, when labelled, two lung-like elliptical masks (0 / 255)."""

    def __init__(self, n: int, img_size: int = 128, labelled: bool = False, seed: int = 0):
        self.n, self.img_size, self.labelled, self.seed = n, img_size, labelled, seed

    def __len__(self) -> int:
        return self.n

    def __getitem__(self, index: int):
        s = self.img_size
        rng = np.random.Generator(np.random.PCG64(self.seed * 1000003 + index))
        yy, xx = np.mgrid[0:s, 0:s].astype(np.float32) / s
        img = np.zeros((s, s), np.float32)
        for _ in range(4):
            fx, fy, ph = rng.uniform(0.5, 3.0), rng.uniform(0.5, 3.0), rng.uniform(0, 6.28)
            img += rng.uniform(0.2, 1.0) * np.sin(6.28 * (fx * xx + fy * yy) + ph)
        cx = rng.uniform(0.25, 0.35), rng.uniform(0.65, 0.75)
        cy, rx, ry = rng.uniform(0.45, 0.55), rng.uniform(0.10, 0.16), rng.uniform(0.22, 0.32)
        lungs = [(((xx - c) / rx) ** 2 + ((yy - cy) / ry) ** 2) < 1.0 for c in cx]
        img = img - 1.5 * (lungs[0] | lungs[1])
        img = (img - img.min()) / max(float(img.max() - img.min()), 1e-6)
        img = np.clip(img * 255.0 + rng.normal(0, 4.0, (s, s)), 0, 255).astype(np.uint8)
        x = torch.from_numpy(img)[None]
        if not self.labelled:
            return x
        masks = torch.from_numpy(np.stack([l.astype(np.uint8) * 255 for l in lungs]))
        return x, masks


def build_synthetic_dataloaders(img_size: int = 128, batch_size: int = 16, num_workers: int = 0, labelled: bool = False,
                                n_train: int = 256, n_val: int = 32, device="cuda", rank: int = 0, world_size: int = 1,
                                n_labelled_images: Optional[int] = None, seed: int = 0) -> Dict[str, DeviceLoader]:
    if n_labelled_images is not None:
        n_train = n_labelled_images
    mk = lambda n, shuffle, sd: DeviceLoader(SyntheticXray(n, img_size, labelled, sd), batch_size, shuffle, num_workers,
                                             device, labelled, rank, world_size)
    return {"train": mk(n_train, True, seed), "val": mk(n_val, False, seed + 1), "test": mk(n_val, False, seed + 2)}
