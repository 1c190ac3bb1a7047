"""Kernel schedule of the UNet forward on sm_100a (the body of reference `Unet.forward`,
models/unet_model.py:333-368, restated as a sequence of fused native ops).

Data layout in HBM: activations NHWC bf16; conv weights bf16 [Cout][kh][kw][Cin] (a derived cache
of the fp32 OIHW parameters, rebuilt when a parameter's version counter changes); GroupNorm
statistics, time embeddings and everything scalar fp32.  Skip-connection concats are never
materialised (the conv kernel's K loop walks two sources); the ResnetBlock residual add is fused
into the second GroupNorm+SiLU pass; the mid-attention residual into the to_out conv epilogue; the
LinearAttention residual into the output LayerNorm.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
from torch import Tensor, nn

from . import native as N


class WeightCache:
    """Derived (re-laid-out / concatenated) copies of parameters, keyed by name, invalidated by version."""

    def __init__(self):
        self._store: Dict[str, Tuple[tuple, Tensor]] = {}

    def get(self, key: str, params: Tuple[Tensor, ...], make) -> Tensor:
        sig = tuple((p.data_ptr(), p._version, p.device) for p in params)
        hit = self._store.get(key)
        if hit is not None and hit[0] == sig:
            return hit[1]
        with torch.no_grad():
            val = make(*[p.detach() for p in params])
        self._store[key] = (sig, val)
        return val

    def clear(self):
        self._store.clear()


class UnetEngine:
    def __init__(self, unet: nn.Module, fold_upsample: bool = True, ln_eps: float = 1e-5):
        self.m = unet
        self.cache = WeightCache()
        self.fold_upsample = fold_upsample
        self.ln_eps = ln_eps
        # every ResnetBlock in execution order, for the batched time projection
        m = unet
        self._resblocks: List[nn.Module] = []
        for b1, b2, _, _ in m.downs:
            self._resblocks += [b1, b2]
        self._resblocks += [m.mid_block1, m.mid_block2]
        for b1, b2, _, _ in m.ups:
            self._resblocks += [b1, b2]
        self._resblocks.append(m.final_res_block)
        self._ss_offset: Dict[int, int] = {}
        off = 0
        for rb in self._resblocks:
            self._ss_offset[id(rb)] = off
            off += rb.time_mlp[1].weight.shape[0]
        self._ss_total = off

    # -- derived weights ------------------------------------------------------------------------
    def _f32(self, p: Tensor) -> Tensor:
        return p.detach().float().contiguous() if (p.dtype != torch.float32 or not p.is_contiguous()) else p.detach()

    def _krsc(self, key: str, w: Tensor) -> Tensor:
        return self.cache.get(key + ":krsc", (w,), lambda a: N.weight_to_krsc(a.float()))

    def _folded(self, key: str, w: Tensor) -> Tensor:
        return self.cache.get(key + ":fold", (w,), lambda a: N.fold_upsample_weight(a.float()))

    def _time_cat(self) -> Tuple[Tensor, Tensor]:
        ws = tuple(rb.time_mlp[1].weight for rb in self._resblocks)
        bs = tuple(rb.time_mlp[1].bias for rb in self._resblocks)
        w = self.cache.get("time_cat:w", ws, lambda *a: torch.cat([x.float() for x in a], dim=0).contiguous())
        b = self.cache.get("time_cat:b", bs, lambda *a: torch.cat([x.float() for x in a], dim=0).contiguous())
        return w, b

    # -- blocks -----------------------------------------------------------------------------------
    def _block(self, key: str, blk: nn.Module, x0: Tensor, x1: Optional[Tensor], ss, ss_off: int,
               residual: Optional[Tensor]) -> Tensor:
        cout = blk.proj.weight.shape[0]
        h, part = N.conv_igemm(x0, self._krsc(key + ".proj", blk.proj.weight), N.MODE_3X3, cout,
                               bias=self._f32(blk.proj.bias), src1=x1, gn_groups=blk.norm.num_groups)
        return N.gn_silu(h, part, self._f32(blk.norm.weight), self._f32(blk.norm.bias), blk.norm.num_groups,
                         eps=blk.norm.eps, scale_shift=ss, ss_offset=ss_off, residual=residual)

    def _resblock(self, key: str, rb: nn.Module, x0: Tensor, x1: Optional[Tensor], tproj: Optional[Tensor]) -> Tensor:
        ss_off = self._ss_offset[id(rb)]
        h = self._block(key + ".block1", rb.block1, x0, x1, tproj, ss_off, None)
        if isinstance(rb.res_conv, nn.Conv2d):
            cout = rb.res_conv.weight.shape[0]
            res = N.conv_igemm(x0, self._krsc(key + ".res_conv", rb.res_conv.weight), N.MODE_1X1, cout,
                               bias=self._f32(rb.res_conv.bias), src1=x1)
        else:
            if x1 is not None:
                raise RuntimeError("identity residual with a two-source input")
            res = x0
        return self._block(key + ".block2", rb.block2, h, None, None, 0, res)

    def _linear_attention(self, key: str, wrap: nn.Module, x: Tensor) -> Tensor:
        pre, att = wrap.fn.norm, wrap.fn.fn
        c = x.shape[-1]
        y = N.layernorm(x, self._f32(pre.g).reshape(-1), eps=self.ln_eps)
        qkv = N.conv_igemm(y, self._krsc(key + ".to_qkv", att.to_qkv.weight), N.MODE_1X1, att.to_qkv.weight.shape[0])
        o = N.linear_attention(qkv, att.heads, att.dim_head, att.scale)
        o = N.conv_igemm(o, self._krsc(key + ".to_out", att.to_out[0].weight), N.MODE_1X1, c,
                         bias=self._f32(att.to_out[0].bias))
        return N.layernorm(o, self._f32(att.to_out[1].g).reshape(-1), eps=self.ln_eps, residual=x)

    def _mid_attention(self, key: str, wrap: nn.Module, x: Tensor) -> Tensor:
        pre, att = wrap.fn.norm, wrap.fn.fn
        c = x.shape[-1]
        y = N.layernorm(x, self._f32(pre.g).reshape(-1), eps=self.ln_eps)
        qkv = N.conv_igemm(y, self._krsc(key + ".to_qkv", att.to_qkv.weight), N.MODE_1X1, att.to_qkv.weight.shape[0])
        o = N.attention(qkv, att.heads, att.dim_head, float(att.scale))
        return N.conv_igemm(o, self._krsc(key + ".to_out", att.to_out.weight), N.MODE_1X1, c,
                            bias=self._f32(att.to_out.bias), residual=x)

    # -- whole network ----------------------------------------------------------------------------
    def forward(self, x: Tensor, timestep: Optional[Tensor], want_features: bool = False, skip_tail: bool = False):
        m = self.m
        if not x.is_cuda:
            raise RuntimeError("tedm_b200.Unet runs on CUDA (sm_100a) only; there is no CPU fallback")
        if x.dim() != 4 or x.shape[1] != m.channels:
            raise ValueError(f"expected input (B, {m.channels}, H, W), got {tuple(x.shape)}")
        x = x.detach().float().contiguous()
        tproj = None
        if timestep is not None:
            t = timestep.detach().to(device=x.device, dtype=torch.int64).contiguous()
            pe = m.time_mlp[0]
            freq = self.cache.get("freq", (m.time_mlp[1].weight,), lambda w: pe.frequencies(w.device).float().contiguous())
            temb = N.time_embed(t, freq, self._f32(m.time_mlp[1].weight), self._f32(m.time_mlp[1].bias),
                                self._f32(m.time_mlp[3].weight), self._f32(m.time_mlp[3].bias))
            wcat, bcat = self._time_cat()
            tproj = N.time_proj(temb, wcat, bcat)

        h = N.stem_conv7x7(x, self._f32(m.init_conv.weight), self._f32(m.init_conv.bias))
        stem = h
        skips: List[Tensor] = []
        for i, (b1, b2, attn, down) in enumerate(m.downs):
            k = f"downs.{i}"
            h = self._resblock(k + ".0", b1, h, None, tproj)
            skips.append(h)
            h = self._resblock(k + ".1", b2, h, None, tproj)
            h = self._linear_attention(k + ".2", attn, h)
            skips.append(h)
            mode = N.MODE_4X4S2 if down.kernel_size[0] == 4 else N.MODE_3X3
            h = N.conv_igemm(h, self._krsc(k + ".3", down.weight), mode, down.weight.shape[0], bias=self._f32(down.bias))
        h = self._resblock("mid_block1", m.mid_block1, h, None, tproj)
        h = self._mid_attention("mid_attn", m.mid_attn, h)
        h = self._resblock("mid_block2", m.mid_block2, h, None, tproj)
        feats: List[Tensor] = []
        n_up = len(m.ups)
        for i, (b1, b2, attn, up) in enumerate(m.ups):
            k = f"ups.{i}"
            h = self._resblock(k + ".0", b1, h, skips.pop(), tproj)
            h = self._resblock(k + ".1", b2, h, skips.pop(), tproj)
            h = self._linear_attention(k + ".2", attn, h)
            feats.append(h)
            if skip_tail and i == n_up - 1:
                return None, feats
            if isinstance(up, nn.Sequential):          # Upsample: nearest x2 + 3x3 conv
                conv = up[1]
                if self.fold_upsample:
                    h = N.conv_igemm(h, self._folded(k + ".3.1", conv.weight), N.MODE_UP3X3, conv.weight.shape[0],
                                     bias=self._f32(conv.bias))
                else:
                    h = N.conv_igemm(N.upsample2x(h), self._krsc(k + ".3.1", conv.weight), N.MODE_3X3,
                                     conv.weight.shape[0], bias=self._f32(conv.bias))
            else:
                h = N.conv_igemm(h, self._krsc(k + ".3", up.weight), N.MODE_3X3, up.weight.shape[0], bias=self._f32(up.bias))
        h = self._resblock("final_res_block", m.final_res_block, h, stem, tproj)
        out = N.final_conv1x1(h, self._f32(m.final_conv.weight).reshape(m.out_dim, -1), self._f32(m.final_conv.bias))
        return (out, feats) if want_features else out
