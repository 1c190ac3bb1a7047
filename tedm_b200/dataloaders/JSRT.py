"""JSRT image / lung-mask pairs (reference: dataloaders/JSRT.py:15-94), kept as uint8 until they are on the GPU.

The reference's `__getitem__` returns fp32 (1, S, S) image = ToTensor(PIL 'L' resized) and fp32 label =
sum over the two lung masks of (ToTensor(mask) > .5), made binary when the lungs overlap.  Here the same PIL
decode + resize runs on the host, but the dataset hands out the raw uint8 planes: image (1, S, S) and masks
(K, S, S); `DeviceLoader` finishes the arithmetic on the device (bit-exact: u8 / 255, u8 >= 128)."""
from __future__ import annotations

import os
from pathlib import Path
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch
from torch import Tensor
from torch.utils.data import Dataset, Subset

from .device_loader import DeviceLoader, read_csv_columns

PROJECT_DATA = Path(os.path.realpath(__file__)).parent.parent.parent / "data"


class JSRTDataset(Dataset):
    def __init__(self, base_path, csv_path, csv_name: str, img_size: int = 128,
                 labels: Sequence[str] = ("right lung", "left lung"), **kwargs) -> None:
        cols = read_csv_columns(os.path.join(csv_path, csv_name), ("path", "id"))
        self.paths, self.ids = cols["path"], cols["id"]
        self.base_path = Path(base_path)
        self.labels = tuple(labels)
        self.img_size = img_size

    def _load_u8(self, fname) -> Tensor:
        from PIL import Image
        img = Image.open(self.base_path / fname).convert("L").resize((self.img_size, self.img_size))
        return torch.from_numpy(np.asarray(img, dtype=np.uint8).copy())

    def __getitem__(self, index: int) -> Tuple[Tensor, Tensor]:
        img = self._load_u8(self.paths[index])[None]
        masks = torch.stack([self._load_u8(f"SCR/masks/{item}/{self.ids[index]}.gif") for item in self.labels])
        return img, masks

    def __len__(self) -> int:
        return len(self.paths)


def build_dataloaders(data_dir, img_size: int = 128, batch_size: int = 16, num_workers: int = 1,
                      n_labelled_images: Optional[int] = None, device="cuda", csv_dir=PROJECT_DATA,
                      rank: int = 0, world_size: int = 1, **kwargs) -> Dict[str, DeviceLoader]:
    """Same keys and batching as the reference's build_dataloaders (JSRT.py:18-47); under data parallelism every
    rank draws its own shard of each split."""
    train_ds: Dataset = JSRTDataset(data_dir, csv_dir, "JSRT_train_split.csv", img_size)
    if n_labelled_images is not None:
        train_ds = Subset(train_ds, range(n_labelled_images))
        print(f"Using {n_labelled_images} labelled images")
    val_ds = JSRTDataset(data_dir, csv_dir, "JSRT_val_split.csv", img_size)
    test_ds = JSRTDataset(data_dir, csv_dir, "JSRT_test_split.csv", img_size)
    mk = lambda ds, shuffle: DeviceLoader(ds, batch_size, shuffle, num_workers, device, labelled=True, rank=rank,
                                          world_size=world_size)
    return {"train": mk(train_ds, True), "val": mk(val_ds, False), "test": mk(test_ds, False)}
