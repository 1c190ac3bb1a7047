"""Helpers the hot path shares with its callers (reference: trainers/utils.py:18-98)."""
from __future__ import annotations

import os
import random
from inspect import isfunction
from typing import Any, Optional, Tuple

import numpy as np
import torch
from torch import Tensor


def seed_everything(seed: int) -> None:  # trainers/utils.py:18-25
    random.seed(seed)
    os.environ["PYTHONHASHSEED"] = str(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed(seed)


def normalize_to_neg_one_to_one(img: Tensor) -> Tensor:  # trainers/utils.py:28-29
    return img * 2 - 1


def unnormalize_to_zero_to_one(img: Tensor) -> Tensor:  # trainers/utils.py:32-33
    return (img + 1) * 0.5


def exists(x: Any) -> bool:
    return x is not None


def default(val: Any, d: Any) -> Any:
    if exists(val):
        return val
    return d() if isfunction(d) else d


def get_index_from_list(vals: Tensor, t: Tensor, x_shape: Tuple[int, ...]) -> Tensor:  # trainers/utils.py:48-59
    return vals.gather(-1, t).reshape(t.shape[0], *((1,) * (len(x_shape) - 1)))


@torch.no_grad()
def sample_images(diffusion_model, T: int, img_size: int, batch: int, channels: int = 1,
                  cond: Optional[Tensor] = None, n_snapshots: int = 8, generator: Optional[torch.Generator] = None):
    """Ancestral sampling driver (reference: sample_plot_image, trainers/utils.py:62-98, without the
    torchvision grid assembly): x_T ~ N(0, I); for t = T-1 .. 0: x <- sample_timestep(x, t).
    Returns (final images in [0,1] space, list of snapshots taken every T/n_snapshots steps)."""
    device = next(diffusion_model.parameters()).device
    img = torch.randn((batch, channels, img_size, img_size), device=device, generator=generator)
    stepsize = max(1, int(T / n_snapshots))
    snaps = []
    for t in range(T - 1, -1, -1):
        img = diffusion_model.sample_timestep(img, t=t, cond=cond)
        if t % stepsize == 0:
            snaps.append(unnormalize_to_zero_to_one(img))
    return unnormalize_to_zero_to_one(img), snaps
