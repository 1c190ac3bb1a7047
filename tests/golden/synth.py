"""Deterministic synthetic weights / inputs shared by make_golden.py (run against the live
reference in the build container) and the tests (run anywhere, incl. the GPU box).

Nothing here depends on torch's RNG stream: every tensor is drawn from a numpy PCG64
generator seeded by crc32(key) ^ seed, so fixtures stay valid wherever numpy is the same.
Scales follow PyTorch's default initialisers (uniform +-1/sqrt(fan_in)) so activations are
realistic; norm gains/biases and BatchNorm running statistics are perturbed away from 1/0 so
that no term of the arithmetic is trivially the identity.
"""
import zlib
from typing import Dict, Sequence, Tuple

import numpy as np
import torch


def _rng(key: str, seed: int) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64((zlib.crc32(key.encode()) ^ (seed * 0x9E3779B1)) & 0xFFFFFFFF))


def synth_tensor(key: str, shape: Tuple[int, ...], seed: int = 0) -> torch.Tensor:
    g = _rng(key, seed)
    leaf = key.rsplit(".", 1)[-1]
    if leaf == "num_batches_tracked":
        return torch.tensor(7, dtype=torch.long)
    if leaf == "running_var":
        a = g.uniform(0.5, 1.5, shape)
    elif leaf == "running_mean":
        a = g.uniform(-0.3, 0.3, shape)
    elif leaf == "g" or (leaf == "weight" and len(shape) == 1):       # LayerNorm gain, GroupNorm/BN weight
        a = 1.0 + g.uniform(-0.2, 0.2, shape)
    elif leaf == "bias":
        a = g.uniform(-0.1, 0.1, shape)
    elif leaf == "weight":
        fan_in = int(np.prod(shape[1:]))
        b = 1.0 / np.sqrt(fan_in)
        a = g.uniform(-b, b, shape)
    else:
        raise KeyError(key)
    return torch.from_numpy(np.asarray(a, dtype=np.float32))


def synth_state_dict(shapes: Dict[str, Tuple[int, ...]], seed: int = 0) -> Dict[str, torch.Tensor]:
    return {k: synth_tensor(k, tuple(s), seed) for k, s in shapes.items()}


def synth_images(b: int, size: int, seed: int = 0, channels: int = 1) -> torch.Tensor:
    """Smooth-ish images in [0,1): low-frequency blobs + pixel noise (what ToTensor() yields in range)."""
    g = _rng(f"images{b}x{size}", seed)
    yy, xx = np.meshgrid(np.linspace(-1, 1, size), np.linspace(-1, 1, size), indexing="ij")
    out = np.empty((b, channels, size, size), np.float32)
    for i in range(b):
        for c in range(channels):
            cx, cy, r = g.uniform(-0.4, 0.4), g.uniform(-0.4, 0.4), g.uniform(0.3, 0.8)
            blob = np.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / (r * r))
            out[i, c] = np.clip(0.7 * blob + 0.3 * g.uniform(0, 1, (size, size)), 0, 0.999)
    return torch.from_numpy(out)


def synth_noise(shape: Sequence[int], seed: int = 0, tag: str = "noise") -> torch.Tensor:
    g = _rng(f"{tag}{tuple(shape)}", seed)
    return torch.from_numpy(g.standard_normal(tuple(shape)).astype(np.float32))


def synth_timesteps(b: int, T: int = 1000, seed: int = 0) -> torch.Tensor:
    g = _rng(f"t{b}", seed)
    return torch.from_numpy(g.integers(0, T, (b,)).astype(np.int64))
