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


# ---- the submodules are callable on their own -------------------------------------------------
# `Unet.forward` never goes through the `forward` methods below (it runs the fused engine schedule), but
# code that walks the tree does: the reference's contrastive encoders call `self.init_conv(x)`,
# `self.downs[i][j](x, t)` ... directly (models/global_local_cl.py:32-50,74-107).  Each method takes and
# returns the reference's NCHW fp32 tensors and dispatches to the same native kernels the engine uses,
# with an NCHW fp32 <-> NHWC bf16 conversion at its edges.  Inference only: the native backward exists
# for the whole network (engine.UnetFunction), not per submodule.
def _nhwc(x: Tensor) -> Tensor:
    from .. import native as N
    if not x.is_cuda:
        raise RuntimeError("tedm_b200 modules run on CUDA (sm_100a) only; there is no CPU fallback")
    if torch.is_grad_enabled() and x.requires_grad:
        raise RuntimeError("tedm_b200 submodules are forward-only when called on their own; train through Unet.forward")
    return N.nchw_to_nhwc_bf16(x.detach().float().contiguous())


def _nchw(y: Tensor) -> Tensor:
    from .. import native as N
    return N.nhwc_to_nchw_f32(y)


def _f32(p: Tensor) -> Tensor:
    return p.detach().float().contiguous()


class _Derived:
    """bf16 operand copies of a module's own conv weights, rebuilt when the parameter changes."""

    def _derived(self, key: str, w: Tensor, make):
        from ..engine import WeightCache
        if not hasattr(self, "_wc"):
            object.__setattr__(self, "_wc", WeightCache())
        return self._wc.get(key, (w,), make)


class Conv2d(nn.Conv2d, _Derived):
    """nn.Conv2d (same parameters, same state_dict) whose forward runs the native kernels: the 7x7 stem, 1x1 / 3x3 /
    4x4-stride-2 implicit-GEMM convs on tcgen05, and the per-pixel dot of the output conv."""

    def forward(self, x: Tensor) -> Tensor:
        from .. import native as N
        k, st, pad = self.kernel_size, self.stride, self.padding
        cin, cout = self.in_channels, self.out_channels
        bias = _f32(self.bias) if self.bias is not None else None
        if k == (7, 7) and st == (1, 1) and pad == (3, 3) and cout % 8 == 0:
            if not x.is_cuda:
                raise RuntimeError("tedm_b200 modules run on CUDA (sm_100a) only; there is no CPU fallback")
            zero = torch.zeros(cout, device=x.device)
            return _nchw(N.stem_conv7x7(x.detach().float().contiguous(), _f32(self.weight), bias if bias is not None else zero))
        h = _nhwc(x)
        if k == (1, 1) and st == (1, 1) and cout % 64:
            zero = torch.zeros(cout, device=x.device)
            return N.final_conv1x1(h, _f32(self.weight).reshape(cout, -1), bias if bias is not None else zero)
        mode = {((1, 1), (1, 1), (0, 0)): N.MODE_1X1, ((3, 3), (1, 1), (1, 1)): N.MODE_3X3,
                ((4, 4), (2, 2), (1, 1)): N.MODE_4X4S2}.get((k, st, pad))
        if mode is None:
            raise NotImplementedError(f"Conv2d(kernel {k}, stride {st}, padding {pad}) is not part of the UNet")
        w = self._derived("krsc", self.weight, lambda a: N.weight_to_krsc(a.float()))
        return _nchw(N.conv_igemm(h, w.reshape(-1), mode, cout, bias=bias))


class Residual(nn.Module):  # unet_model.py:29-36
    def __init__(self, fn: nn.Module):
        super().__init__()
        self.fn = fn

    def forward(self, x: Tensor, *args, **kwargs) -> Tensor:
        from .. import native as N
        y = self.fn(x, *args, **kwargs)
        return _nchw(N.add_bf16(_nhwc(y), _nhwc(x)))


class LayerNorm(nn.Module):  # unet_model.py:52-61
    def __init__(self, dim: int):
        super().__init__()
        self.g = nn.Parameter(torch.ones(1, dim, 1, 1))

    def forward(self, x: Tensor) -> Tensor:
        from .. import native as N
        return _nchw(N.layernorm(_nhwc(x), _f32(self.g).reshape(-1), eps=1e-5))


class PreNorm(nn.Module):  # unet_model.py:64-73
    def __init__(self, dim: int, fn: nn.Module):
        super().__init__()
        self.fn = fn
        self.norm = LayerNorm(dim)

    def forward(self, x: Tensor) -> Tensor:
        return self.fn(self.norm(x))


class SinusoidalPosEmb(nn.Module):  # unet_model.py:76-93 (parameter-free)
    def __init__(self, dim: int):
        super().__init__()
        self.dim = dim

    def frequencies(self, device) -> Tensor:
        half_dim = self.dim // 2
        step = math.log(10000) / (half_dim - 1)
        return torch.exp(torch.arange(half_dim, device=device) * -step)

    def forward(self, x: Tensor) -> Tensor:
        arg = x[:, None] * self.frequencies(x.device)[None, :]
        return torch.cat((arg.sin(), arg.cos()), dim=-1)


class LearnedSinusoidalPosEmb(nn.Module):  # unet_model.py:96-114
    """Learned-frequency embedding [t, sin(2 pi t w), cos(2 pi t w)]: (B,) -> (B, dim + 1).  No reference entry point turns
    it on (`learned_sinusoidal_cond` defaults to False everywhere), so it stays on the host-side torch path;
    the time MLP output it feeds is consumed by the native kernels like the fixed embedding's."""

    def __init__(self, dim: int):
        super().__init__()
        assert dim % 2 == 0
        self.weights = nn.Parameter(torch.randn(dim // 2))

    def forward(self, x: Tensor) -> Tensor:
        col = x.reshape(-1, 1).to(self.weights.dtype)
        phase = col * self.weights.reshape(1, -1) * (2 * math.pi)
        return torch.cat((col, phase.sin(), phase.cos()), dim=-1)


class Block(nn.Module, _Derived):  # unet_model.py:119-135
    def __init__(self, dim: int, dim_out: int, groups: int = 8):
        super().__init__()
        self.proj = Conv2d(dim, dim_out, 3, padding=1)
        self.norm = nn.GroupNorm(groups, dim_out)
        self.act = nn.SiLU()

    def _run(self, h: Tensor, ss: Optional[Tensor], residual: Optional[Tensor] = None) -> Tensor:
        from .. import native as N
        w = self.proj._derived("krsc", self.proj.weight, lambda a: N.weight_to_krsc(a.float()))
        y, part = N.conv_igemm(h, w.reshape(-1), N.MODE_3X3, self.proj.out_channels, bias=_f32(self.proj.bias),
                               gn_groups=self.norm.num_groups)
        return N.gn_silu(y, part, _f32(self.norm.weight), _f32(self.norm.bias), self.norm.num_groups, eps=self.norm.eps,
                         scale_shift=ss, ss_offset=0, residual=residual)

    def forward(self, x: Tensor, scale_shift=None) -> Tensor:
        ss = None
        if scale_shift is not None:
            scale, shift = scale_shift
            b = x.shape[0]
            ss = torch.cat([scale.reshape(b, -1), shift.reshape(b, -1)], dim=1).float().contiguous()
        return _nchw(self._run(_nhwc(x), ss))


class ResnetBlock(nn.Module):  # unet_model.py:138-175
    def __init__(self, dim: int, dim_out: int, *, time_emb_dim: Optional[int] = None, groups: int = 8):
        super().__init__()
        self.time_mlp = nn.Sequential(nn.SiLU(), nn.Linear(time_emb_dim, dim_out * 2)) if exists(time_emb_dim) else None
        self.block1 = Block(dim, dim_out, groups=groups)
        self.block2 = Block(dim_out, dim_out, groups=groups)
        self.res_conv = Conv2d(dim, dim_out, 1) if dim != dim_out else nn.Identity()

    def forward(self, x: Tensor, time_emb: Optional[Tensor] = None) -> Tensor:
        from .. import native as N
        ss = None
        if exists(self.time_mlp) and exists(time_emb):
            lin = self.time_mlp[1]
            ss = N.time_proj(time_emb.detach().float().contiguous(), _f32(lin.weight), _f32(lin.bias))   # SiLU + Linear
        h0 = _nhwc(x)
        h = self.block1._run(h0, ss)
        res = _nhwc(self.res_conv(x)) if isinstance(self.res_conv, nn.Conv2d) else h0
        return _nchw(self.block2._run(h, None, residual=res))


class LinearAttention(nn.Module):  # unet_model.py:178-210
    def __init__(self, dim: int, heads: int = 4, dim_head: int = 32):
        super().__init__()
        self.scale = dim_head ** -0.5
        self.heads = heads
        self.dim_head = dim_head
        hidden_dim = dim_head * heads
        self.to_qkv = Conv2d(dim, hidden_dim * 3, 1, bias=False)
        self.to_out = nn.Sequential(Conv2d(hidden_dim, dim, 1), LayerNorm(dim))

    def forward(self, x: Tensor) -> Tensor:
        from .. import native as N
        krsc = lambda a: N.weight_to_krsc(a.float())
        qkv = N.conv_igemm(_nhwc(x), self.to_qkv._derived("krsc", self.to_qkv.weight, krsc).reshape(-1), N.MODE_1X1,
                           self.to_qkv.out_channels)
        o = N.linear_attention(qkv, self.heads, self.dim_head, self.scale)
        conv, ln = self.to_out[0], self.to_out[1]
        o = N.conv_igemm(o, conv._derived("krsc", conv.weight, krsc).reshape(-1), N.MODE_1X1, conv.out_channels,
                         bias=_f32(conv.bias))
        return _nchw(N.layernorm(o, _f32(ln.g).reshape(-1), eps=1e-5))


class Attention(nn.Module):  # unet_model.py:213-241
    def __init__(self, dim: int, heads: int = 4, dim_head: int = 32, scale: int = 16):
        super().__init__()
        self.scale = scale
        self.heads = heads
        self.dim_head = dim_head
        hidden_dim = dim_head * heads
        self.to_qkv = Conv2d(dim, hidden_dim * 3, 1, bias=False)
        self.to_out = Conv2d(hidden_dim, dim, 1)

    def forward(self, x: Tensor) -> Tensor:
        from .. import native as N
        krsc = lambda a: N.weight_to_krsc(a.float())
        qkv = N.conv_igemm(_nhwc(x), self.to_qkv._derived("krsc", self.to_qkv.weight, krsc).reshape(-1), N.MODE_1X1,
                           self.to_qkv.out_channels)
        o = N.attention(qkv, self.heads, self.dim_head, float(self.scale))
        return _nchw(N.conv_igemm(o, self.to_out._derived("krsc", self.to_out.weight, krsc).reshape(-1), N.MODE_1X1,
                                  self.to_out.out_channels, bias=_f32(self.to_out.bias)))


def Upsample(dim: int, dim_out: Optional[int] = None) -> nn.Sequential:  # unet_model.py:39-44
    return nn.Sequential(nn.Upsample(scale_factor=2, mode="nearest"), Conv2d(dim, default(dim_out, dim), 3, padding=1))


def Downsample(dim: int, dim_out: Optional[int] = None) -> nn.Conv2d:  # unet_model.py:47-49
    return Conv2d(dim, default(dim_out, dim), 4, 2, 1)


class Unet(nn.Module):
    """Drop-in for the reference `Unet` (same constructor, attribute names and state_dict)."""

    def __init__(self, dim: int = 64, init_dim: Optional[int] = None, out_dim: Optional[int] = None,
                 dim_mults: List[int] = [1, 2, 4, 8], channels: int = 1, resnet_block_groups: int = 8,
                 learned_variance: bool = False, learned_sinusoidal_cond: bool = False,
                 learned_sinusoidal_dim: int = 16, **kwargs):
        super().__init__()
        self.channels = channels
        self.dim = dim
        self.groups = resnet_block_groups
        init_dim = default(init_dim, dim)
        self.init_conv = Conv2d(channels, init_dim, 7, padding=3)
        dims = [init_dim, *[dim * m for m in dim_mults]]
        in_out = list(zip(dims[:-1], dims[1:]))
        time_dim = dim * 4
        self.learned_sinusoidal_cond = learned_sinusoidal_cond
        if learned_sinusoidal_cond:              # unet_model.py:279-285: fourier_dim = learned_sinusoidal_dim + 1
            pos_emb, fourier_dim = LearnedSinusoidalPosEmb(learned_sinusoidal_dim), learned_sinusoidal_dim + 1
        else:
            pos_emb, fourier_dim = SinusoidalPosEmb(dim), dim
        self.time_mlp = nn.Sequential(pos_emb, nn.Linear(fourier_dim, time_dim), nn.GELU(), nn.Linear(time_dim, time_dim))

        def block(i, o):
            return ResnetBlock(i, o, time_emb_dim=time_dim, groups=resnet_block_groups)

        self.downs = nn.ModuleList([])
        self.ups = nn.ModuleList([])
        for ind, (dim_in, dim_out) in enumerate(in_out):
            is_last = ind >= len(in_out) - 1
            self.downs.append(nn.ModuleList([
                block(dim_in, dim_in), block(dim_in, dim_in),
                Residual(PreNorm(dim_in, LinearAttention(dim_in))),
                Downsample(dim_in, dim_out) if not is_last else Conv2d(dim_in, dim_out, 3, padding=1)]))
        mid_dim = dims[-1]
        self.mid_block1 = block(mid_dim, mid_dim)
        self.mid_attn = Residual(PreNorm(mid_dim, Attention(mid_dim)))
        self.mid_block2 = block(mid_dim, mid_dim)
        for ind, (dim_in, dim_out) in enumerate(reversed(in_out)):
            is_last = ind == len(in_out) - 1
            self.ups.append(nn.ModuleList([
                block(dim_out + dim_in, dim_out), block(dim_out + dim_in, dim_out),
                Residual(PreNorm(dim_out, LinearAttention(dim_out))),
                Upsample(dim_out, dim_in) if not is_last else Conv2d(dim_out, dim_in, 3, padding=1)]))
        self.out_dim = default(out_dim, channels * (1 if not learned_variance else 2))
        self.final_res_block = block(dim * 2, dim)
        self.final_conv = Conv2d(dim, self.out_dim, 1)
        self._engine = None
        self._engine32 = None
        self.precision = "bf16"
        self.set_precision(kwargs.get("precision", "bf16"))

    # -- execution ------------------------------------------------------------------------------
    def set_precision(self, precision: str) -> "Unet":
        """"bf16" (default): bf16 storage / fp32 accumulation, forward and backward -- the throughput mode.
        "fp32": the reference's default arithmetic (config.py:15) to 1e-4 -- fp32 storage, split-bf16 tensor-core
        convolutions (tedm_b200/engine_fp32.py); inference only."""
        if precision not in ("bf16", "fp32"):
            raise ValueError(f"precision must be 'bf16' or 'fp32', got {precision!r}")
        self.precision = precision
        return self

    @property
    def engine(self):
        if self.precision == "fp32":
            if self._engine32 is None:
                from ..engine_fp32 import UnetEngineF32
                self._engine32 = UnetEngineF32(self)
            return self._engine32
        if self._engine is None:
            from ..engine import UnetEngine
            self._engine = UnetEngine(self)
        return self._engine

    def forward(self, x: Tensor, timestep: Optional[Tensor] = None, cond: Optional[Tensor] = None) -> Tensor:
        """(B, channels, H, W) fp32, (B,) int64 -> (B, out_dim, H, W) fp32.  `cond` is accepted and ignored,
        as in the reference (unet_model.py:333)."""
        params = tuple(self.parameters())
        if self.precision == "fp32":
            if torch.is_grad_enabled() and any(p.requires_grad for p in params) and self.training:
                raise NotImplementedError("precision='fp32' is an inference mode (call under torch.no_grad() / eval()); "
                                          "training runs in the bf16 mode")
            return self.engine.forward(x, timestep)
        if torch.is_grad_enabled() and any(p.requires_grad for p in params):
            # training: forward keeps a tape, backward is the native reverse schedule (tedm_b200/engine.py)
            from ..engine import UnetFunction
            return UnetFunction.apply(self.engine, x, timestep, *params)
        return self.engine.forward(x, timestep)

    def forward_features(self, x: Tensor, timestep: Optional[Tensor], skip_tail: bool = True, time_key=None):
        """Decoder feature maps of `ups[i][2]` (what DatasetDM's hooks capture, datasetDM_model.py:50-53) as
        NHWC bf16 device tensors; with skip_tail the part of the net after the last hooked map is not run."""
        return self.engine.forward(x, timestep, want_features=True, skip_tail=skip_tail, time_key=time_key)
