"""fp32 precision mode of the UNet forward (BASELINE north star: "1e-4 in fp32 mode"; the reference's default
arithmetic is fp32, config.py:15 `mixed_precision` False).  Inference only.

Same schedule as `tedm_b200.engine.UnetEngine.forward` (the body of the reference's `Unet.forward`,
models/unet_model.py:333-368) with every activation kept in fp32 NHWC.  Convolutions still run on the tcgen05
implicit-GEMM kernel: activations and weights are split into bf16 (hi, lo) pairs and the product is
hi*hi + lo*hi + hi*lo with fp32 accumulation in TMEM (three passes of the K loop over the sources
[x_hi, x_lo, x_hi] against [w_hi, w_hi, w_lo]); measured against the fp32 reference this is 1e-5 on the features
and 2.5e-5 on the UNet output.  GroupNorm / SiLU, LayerNorm, both attention cores, the stem and the output conv are
plain fp32 CUDA kernels with exact exp / division (csrc/fp32_mode.cu).  About 3.5x the tensor-core work and 2x the
activation traffic of the bf16 path: a parity mode, not the throughput mode.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
from torch import Tensor, nn

from . import native as N
from .engine import UnetEngine

_FOLD_ROWS = {0: ((0,), (1, 2)), 1: ((0, 1), (2,))}      # output parity -> 3x3 taps feeding folded tap 0 / 1 (csrc/weights.cu)


def split_weight(w_oihw: Tensor, mode: int) -> Tensor:
    """fp32 OIHW parameter -> bf16 operand of the split convolution: the kernel's layout for `mode` with the channel axis
    replaced by cat(w_hi, w_hi, w_lo), w_hi = bf16(w), w_lo = bf16(w - w_hi)."""
    a = w_oihw.detach().float()
    co, ci = a.shape[0], a.shape[1]
    if mode == N.MODE_UP3X3:                  # nearest-x2 upsample folded into the 3x3: [parity][co][a*2+b][ci], summed in fp32
        k = a.new_zeros(4, co, 4, ci)
        for py in (0, 1):
            for px in (0, 1):
                for aa in (0, 1):
                    for bb in (0, 1):
                        k[py * 2 + px, :, aa * 2 + bb, :] = a[:, :, list(_FOLD_ROWS[py][aa])][:, :, :, list(_FOLD_ROWS[px][bb])].sum(dim=(2, 3))
    else:                                     # KRSC [co][tap][ci]
        k = a.permute(0, 2, 3, 1).reshape(co, -1, ci)
    hi = k.to(torch.bfloat16)
    lo = (k - hi.float()).to(torch.bfloat16)
    return torch.cat([hi, hi, lo], dim=-1).contiguous().reshape(-1)


class UnetEngineF32(UnetEngine):
    def __init__(self, unet: nn.Module, ln_eps: float = 1e-5):
        super().__init__(unet, ln_eps=ln_eps)
        self._splits: Dict[int, Tuple[Tensor, Tuple[Tensor, Tensor]]] = {}

    # -- operands -----------------------------------------------------------------------------------
    def _w3(self, key: str, w: Tensor, mode: int) -> Tensor:
        return self.cache.get(key + ":w3", (w,), lambda a: split_weight(a, mode))

    def _split(self, x: Optional[Tensor]):
        """(hi, lo) of an activation, computed once per forward however many convolutions read it."""
        if x is None:
            return None
        hit = self._splits.get(id(x))
        if hit is None or hit[0] is not x:
            hit = (x, N.f32_split(x))
            self._splits[id(x)] = hit
        return hit[1]

    def _conv(self, key: str, conv: nn.Conv2d, mode: int, x0: Tensor, x1: Optional[Tensor] = None, gn_groups: int = 0):
        bias = self._f32(conv.bias) if conv.bias is not None else None
        return N.f32_conv(self._split(x0), self._w3(key, conv.weight, mode), mode, conv.weight.shape[0], bias=bias,
                          src1=self._split(x1), gn_groups=gn_groups)

    # -- blocks (unet_model.py:119-241) -----------------------------------------------------------------
    def _block32(self, key: str, blk: nn.Module, x0: Tensor, x1: Optional[Tensor], ss, ss_off: int, residual: Optional[Tensor]) -> Tensor:
        h, part = self._conv(key + ".proj", blk.proj, N.MODE_3X3, x0, x1, gn_groups=blk.norm.num_groups)
        return N.f32_gn_silu(h, part, self._f32(blk.norm.weight), self._f32(blk.norm.bias), blk.norm.num_groups, eps=blk.norm.eps,
                             scale_shift=ss, ss_offset=ss_off, residual=residual)

    def _resblock32(self, key: str, rb: nn.Module, x0: Tensor, x1: Optional[Tensor], tproj: Optional[Tensor]) -> Tensor:
        h = self._block32(key + ".block1", rb.block1, x0, x1, tproj, self._ss_offset[id(rb)], None)
        if isinstance(rb.res_conv, nn.Conv2d):
            res = self._conv(key + ".res_conv", rb.res_conv, N.MODE_1X1, x0, x1)
        else:
            if x1 is not None:
                raise RuntimeError("identity residual with a two-source input")
            res = x0
        return self._block32(key + ".block2", rb.block2, h, None, None, 0, res)

    def _linear_attention32(self, key: str, wrap: nn.Module, x: Tensor) -> Tensor:
        pre, att = wrap.fn.norm, wrap.fn.fn
        y = N.f32_layernorm(x, self._f32(pre.g).reshape(-1), eps=self.ln_eps)
        qkv = self._conv(key + ".to_qkv", att.to_qkv, N.MODE_1X1, y)
        o = N.f32_linear_attention(qkv, att.heads, att.dim_head, att.scale)
        o2 = self._conv(key + ".to_out", att.to_out[0], N.MODE_1X1, o)
        return N.f32_layernorm(o2, self._f32(att.to_out[1].g).reshape(-1), eps=self.ln_eps, residual=x)

    def _mid_attention32(self, key: str, wrap: nn.Module, x: Tensor) -> Tensor:
        pre, att = wrap.fn.norm, wrap.fn.fn
        y = N.f32_layernorm(x, self._f32(pre.g).reshape(-1), eps=self.ln_eps)
        qkv = self._conv(key + ".to_qkv", att.to_qkv, N.MODE_1X1, y)
        o = N.f32_attention(qkv, att.heads, att.dim_head, float(att.scale))
        return N.f32_add(self._conv(key + ".to_out", att.to_out, N.MODE_1X1, o), x)

    @staticmethod
    def _fire(mod: nn.Module, ins, out: Tensor) -> None:
        if not mod._forward_hooks:
            return
        xin = [i.permute(0, 3, 1, 2).contiguous() for i in ins]
        args = (torch.cat(xin, dim=1) if len(xin) > 1 else xin[0],)
        o = out.permute(0, 3, 1, 2).contiguous()
        for hook in list(mod._forward_hooks.values()):
            if hook(mod, args, o) is not None:
                raise NotImplementedError("forward hooks that replace a submodule's output are not supported by the fused engine")

    # -- whole network ------------------------------------------------------------------------------------
    def forward(self, x: Tensor, timestep: Optional[Tensor], want_features: bool = False, skip_tail: bool = False,
                tape=None, time_key=None):
        m = self.m
        if tape is not None:
            raise NotImplementedError("precision='fp32' is an inference mode; train in the bf16 mode")
        if not x.is_cuda:
            raise RuntimeError("tedm_b200.Unet runs on CUDA (sm_100a) only; there is no CPU fallback")
        if x.dim() != 4 or x.shape[1] != m.channels:
            raise ValueError(f"expected input (B, {m.channels}, H, W), got {tuple(x.shape)}")
        x = x.detach().float().contiguous()
        self._splits.clear()
        try:
            return self._forward(x, timestep, want_features, skip_tail, time_key)
        finally:
            self._splits.clear()

    def _forward(self, x, timestep, want_features, skip_tail, time_key):
        m = self.m
        tproj = self._time_projection(x, timestep, None, time_key)
        h = N.f32_stem_conv7x7(x, self._f32(m.init_conv.weight), self._f32(m.init_conv.bias))
        stem = h
        skips: List[Tensor] = []
        for i, (b1, b2, attn, down) in enumerate(m.downs):
            k = f"downs.{i}"
            hin = h
            h = self._resblock32(k + ".0", b1, h, None, tproj)
            self._fire(b1, [hin], h)
            skips.append(h)
            hin = h
            h = self._resblock32(k + ".1", b2, h, None, tproj)
            self._fire(b2, [hin], h)
            hin = h
            h = self._linear_attention32(k + ".2", attn, h)
            self._fire(attn, [hin], h)
            skips.append(h)
            h = self._conv(k + ".3", down, N.MODE_4X4S2 if down.kernel_size[0] == 4 else N.MODE_3X3, h)
        hin = h
        h = self._resblock32("mid_block1", m.mid_block1, h, None, tproj)
        self._fire(m.mid_block1, [hin], h)
        hin = h
        h = self._mid_attention32("mid_attn", m.mid_attn, h)
        self._fire(m.mid_attn, [hin], h)
        hin = h
        h = self._resblock32("mid_block2", m.mid_block2, h, None, tproj)
        self._fire(m.mid_block2, [hin], h)
        feats: List[Tensor] = []
        n_up = len(m.ups)
        for i, (b1, b2, attn, up) in enumerate(m.ups):
            k = f"ups.{i}"
            hin, sk = h, skips.pop()
            h = self._resblock32(k + ".0", b1, h, sk, tproj)
            self._fire(b1, [hin, sk], h)
            hin, sk = h, skips.pop()
            h = self._resblock32(k + ".1", b2, h, sk, tproj)
            self._fire(b2, [hin, sk], h)
            hin = h
            h = self._linear_attention32(k + ".2", attn, h)
            self._fire(attn, [hin], h)
            feats.append(h)
            if skip_tail and i == n_up - 1:
                return None, feats
            if isinstance(up, nn.Sequential):          # Upsample: nearest x2 folded into the 3x3 conv (weights summed in fp32)
                h = self._conv(k + ".3.1", up[1], N.MODE_UP3X3, h)
            else:
                h = self._conv(k + ".3", up, N.MODE_3X3, h)
        h = self._resblock32("final_res_block", m.final_res_block, h, stem, tproj)
        out = N.f32_final_conv1x1(h, self._f32(m.final_conv.weight).reshape(m.out_dim, -1), self._f32(m.final_conv.bias))
        return (out, feats) if want_features else out
