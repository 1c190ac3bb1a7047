"""Montgomery County chest-X-ray pairs used as an external test set (reference: dataloaders/Montgomery.py:15-61):
right / left lung masks in two files per image.  uint8 planes until on the GPU, like JSRT."""
from __future__ import annotations

import os
from pathlib import Path
from typing import Sequence, Tuple

import numpy as np
import torch
from torch import Tensor
from torch.utils.data import Dataset

from .device_loader import read_csv_columns


class MonDataset(Dataset):
    def __init__(self, base_path, csv_path, csv_name: str, img_size: int = 128,
                 labels: Sequence[str] = ("right lung", "left lung"), **kwargs) -> None:
        self.labels = tuple(labels)
        cols = read_csv_columns(os.path.join(csv_path, csv_name), ("scan",) + self.labels)
        self.scans = cols["scan"]
        self.mask_files = [cols[l] for l in self.labels]
        self.base_path = Path(base_path)
        self.img_size = img_size

    def _load_u8(self, fname) -> Tensor:
        from PIL import Image
        img = Image.open(self.base_path / fname).convert("L").resize((self.img_size, self.img_size))
        return torch.from_numpy(np.asarray(img, dtype=np.uint8).copy())

    def __getitem__(self, index: int) -> Tuple[Tensor, Tensor]:
        masks = torch.stack([self._load_u8(files[index]) for files in self.mask_files])
        return self._load_u8(self.scans[index])[None], masks

    def __len__(self) -> int:
        return len(self.scans)
