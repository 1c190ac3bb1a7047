"""ChestXray-14 images for DDPM pre-training (reference: dataloaders/CXR14.py:21-74), uint8 until on the GPU."""
from __future__ import annotations

import os
from pathlib import Path
from typing import Dict

import numpy as np
import torch
from torch import Tensor
from torch.utils.data import Dataset

from .device_loader import DeviceLoader, read_csv_columns

PROJECT_DATA = Path(os.path.realpath(__file__)).parent.parent.parent / "data"


class CXR14Dataset(Dataset):
    def __init__(self, data_path, csv_path, img_size: int) -> None:
        super().__init__()
        if not os.path.isdir(data_path):
            raise FileNotFoundError(f"CXR14 image directory {data_path} does not exist")
        if not os.path.isfile(csv_path):
            raise FileNotFoundError(f"CXR14 split file {csv_path} does not exist")
        self.data_path = Path(data_path)
        self.files = read_csv_columns(csv_path, ("Image Index",))["Image Index"]
        self.img_size = img_size

    def __len__(self) -> int:
        return len(self.files)

    def __getitem__(self, index: int) -> Tensor:
        from PIL import Image
        img = Image.open(self.data_path / self.files[index]).convert("L").resize((self.img_size, self.img_size))
        return torch.from_numpy(np.asarray(img, dtype=np.uint8).copy())[None]


def build_dataloaders(data_dir, img_size: int = 128, batch_size: int = 16, num_workers: int = 1, device="cuda",
                      csv_dir=PROJECT_DATA, rank: int = 0, world_size: int = 1) -> Dict[str, DeviceLoader]:
    """(CXR14.py:21-47) NB the reference builds train, val and test from the same train_split.csv; kept."""
    mk = lambda shuffle: DeviceLoader(CXR14Dataset(data_dir, Path(csv_dir) / "train_split.csv", img_size), batch_size,
                                      shuffle, num_workers, device, labelled=False, rank=rank, world_size=world_size)
    return {"train": mk(True), "val": mk(False), "test": mk(False)}
