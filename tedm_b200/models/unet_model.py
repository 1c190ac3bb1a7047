"""B200-native DDPM UNet behind the reference's nn.Module surface (models/unet_model.py:246-368).

The class tree below exists to hold parameters under exactly the reference's state_dict keys
(`downs.0.0.block1.proj.weight`, `ups.2.2.fn.fn.to_out.1.g`, ...), in the reference's fp32 OIHW
layout, so checkpoints load both ways and optimisers see ordinary nn.Parameters.  The arithmetic
of `Unet.forward` does not go through these submodules: it is executed by `tedm_b200.engine`, a
schedule of hand-written sm_100a kernels over NHWC bf16 activations (tcgen05 implicit-GEMM convs
with fused bias/residual/GroupNorm statistics, fused GroupNorm+scale/shift+SiLU, fused attention).
There is no PyTorch-op fallback: on a non-CUDA tensor `forward` raises.
"""
from __future__ import annotations

import math
from typing import List, Optional

import torch
from torch import Tensor, nn


def exists(x) -> bool:  # trainers/utils.py:36-38
    return x is not None


def default(val, d):  # trainers/utils.py:41-45
    if exists(val):
        return val
    return d() if callable(d) else d


class Residual(nn.Module):  # unet_model.py:29-36 -- container; fused into the consumer kernels
    def __init__(self, fn: nn.Module):
        super().__init__()
        self.fn = fn


class LayerNorm(nn.Module):  # unet_model.py:52-61
    def __init__(self, dim: int):
        super().__init__()
        self.g = nn.Parameter(torch.ones(1, dim, 1, 1))


class PreNorm(nn.Module):  # unet_model.py:64-73
    def __init__(self, dim: int, fn: nn.Module):
        super().__init__()
        self.fn = fn
        self.norm = LayerNorm(dim)


class SinusoidalPosEmb(nn.Module):  # unet_model.py:76-93 (parameter-free)
    def __init__(self, dim: int):
        super().__init__()
        self.dim = dim

    def frequencies(self, device) -> Tensor:
        half_dim = self.dim // 2
        step = math.log(10000) / (half_dim - 1)
        return torch.exp(torch.arange(half_dim, device=device) * -step)


class Block(nn.Module):  # unet_model.py:119-135
    def __init__(self, dim: int, dim_out: int, groups: int = 8):
        super().__init__()
        self.proj = nn.Conv2d(dim, dim_out, 3, padding=1)
        self.norm = nn.GroupNorm(groups, dim_out)
        self.act = nn.SiLU()


class ResnetBlock(nn.Module):  # unet_model.py:138-175
    def __init__(self, dim: int, dim_out: int, *, time_emb_dim: Optional[int] = None, groups: int = 8):
        super().__init__()
        self.time_mlp = nn.Sequential(nn.SiLU(), nn.Linear(time_emb_dim, dim_out * 2)) if exists(time_emb_dim) else None
        self.block1 = Block(dim, dim_out, groups=groups)
        self.block2 = Block(dim_out, dim_out, groups=groups)
        self.res_conv = nn.Conv2d(dim, dim_out, 1) if dim != dim_out else nn.Identity()


class LinearAttention(nn.Module):  # unet_model.py:178-210
    def __init__(self, dim: int, heads: int = 4, dim_head: int = 32):
        super().__init__()
        self.scale = dim_head ** -0.5
        self.heads = heads
        self.dim_head = dim_head
        hidden_dim = dim_head * heads
        self.to_qkv = nn.Conv2d(dim, hidden_dim * 3, 1, bias=False)
        self.to_out = nn.Sequential(nn.Conv2d(hidden_dim, dim, 1), LayerNorm(dim))


class Attention(nn.Module):  # unet_model.py:213-241
    def __init__(self, dim: int, heads: int = 4, dim_head: int = 32, scale: int = 16):
        super().__init__()
        self.scale = scale
        self.heads = heads
        self.dim_head = dim_head
        hidden_dim = dim_head * heads
        self.to_qkv = nn.Conv2d(dim, hidden_dim * 3, 1, bias=False)
        self.to_out = nn.Conv2d(hidden_dim, dim, 1)


def Upsample(dim: int, dim_out: Optional[int] = None) -> nn.Sequential:  # unet_model.py:39-44
    return nn.Sequential(nn.Upsample(scale_factor=2, mode="nearest"), nn.Conv2d(dim, default(dim_out, dim), 3, padding=1))


def Downsample(dim: int, dim_out: Optional[int] = None) -> nn.Conv2d:  # unet_model.py:47-49
    return nn.Conv2d(dim, default(dim_out, dim), 4, 2, 1)


class Unet(nn.Module):
    """Drop-in for the reference `Unet` (same constructor, attribute names and state_dict)."""

    def __init__(self, dim: int = 64, init_dim: Optional[int] = None, out_dim: Optional[int] = None,
                 dim_mults: List[int] = [1, 2, 4, 8], channels: int = 1, resnet_block_groups: int = 8,
                 learned_variance: bool = False, learned_sinusoidal_cond: bool = False,
                 learned_sinusoidal_dim: int = 16, **kwargs):
        super().__init__()
        if learned_sinusoidal_cond:
            # never enabled by any reference entry point
            raise NotImplementedError("learned_sinusoidal_cond=True is not part of the B200 hot path")
        self.channels = channels
        self.dim = dim
        self.groups = resnet_block_groups
        init_dim = default(init_dim, dim)
        self.init_conv = nn.Conv2d(channels, init_dim, 7, padding=3)
        dims = [init_dim, *[dim * m for m in dim_mults]]
        in_out = list(zip(dims[:-1], dims[1:]))
        time_dim = dim * 4
        self.learned_sinusoidal_cond = learned_sinusoidal_cond
        self.time_mlp = nn.Sequential(SinusoidalPosEmb(dim), nn.Linear(dim, time_dim), nn.GELU(),
                                      nn.Linear(time_dim, time_dim))

        def block(i, o):
            return ResnetBlock(i, o, time_emb_dim=time_dim, groups=resnet_block_groups)

        self.downs = nn.ModuleList([])
        self.ups = nn.ModuleList([])
        for ind, (dim_in, dim_out) in enumerate(in_out):
            is_last = ind >= len(in_out) - 1
            self.downs.append(nn.ModuleList([
                block(dim_in, dim_in), block(dim_in, dim_in),
                Residual(PreNorm(dim_in, LinearAttention(dim_in))),
                Downsample(dim_in, dim_out) if not is_last else nn.Conv2d(dim_in, dim_out, 3, padding=1)]))
        mid_dim = dims[-1]
        self.mid_block1 = block(mid_dim, mid_dim)
        self.mid_attn = Residual(PreNorm(mid_dim, Attention(mid_dim)))
        self.mid_block2 = block(mid_dim, mid_dim)
        for ind, (dim_in, dim_out) in enumerate(reversed(in_out)):
            is_last = ind == len(in_out) - 1
            self.ups.append(nn.ModuleList([
                block(dim_out + dim_in, dim_out), block(dim_out + dim_in, dim_out),
                Residual(PreNorm(dim_out, LinearAttention(dim_out))),
                Upsample(dim_out, dim_in) if not is_last else nn.Conv2d(dim_out, dim_in, 3, padding=1)]))
        self.out_dim = default(out_dim, channels * (1 if not learned_variance else 2))
        self.final_res_block = block(dim * 2, dim)
        self.final_conv = nn.Conv2d(dim, self.out_dim, 1)
        self._engine = None

    # -- execution ------------------------------------------------------------------------------
    @property
    def engine(self):
        if self._engine is None:
            from ..engine import UnetEngine
            self._engine = UnetEngine(self)
        return self._engine

    def forward(self, x: Tensor, timestep: Optional[Tensor] = None, cond: Optional[Tensor] = None) -> Tensor:
        """(B, channels, H, W) fp32, (B,) int64 -> (B, out_dim, H, W) fp32.  `cond` is accepted and ignored,
        as in the reference (unet_model.py:333)."""
        params = tuple(self.parameters())
        if torch.is_grad_enabled() and any(p.requires_grad for p in params):
            # training: forward keeps a tape, backward is the native reverse schedule (tedm_b200/engine.py)
            from ..engine import UnetFunction
            return UnetFunction.apply(self.engine, x, timestep, *params)
        return self.engine.forward(x, timestep)

    def forward_features(self, x: Tensor, timestep: Optional[Tensor], skip_tail: bool = True, time_key=None):
        """Decoder feature maps of `ups[i][2]` (what DatasetDM's hooks capture, datasetDM_model.py:50-53) as
        NHWC bf16 device tensors; with skip_tail the part of the net after the last hooked map is not run."""
        return self.engine.forward(x, timestep, want_features=True, skip_tail=skip_tail, time_key=time_key)
