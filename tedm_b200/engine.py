"""Kernel schedule of the UNet forward AND backward on sm_100a (the body of reference `Unet.forward`,
models/unet_model.py:333-368, restated as a sequence of fused native ops, plus the hand-scheduled
reverse pass that torch autograd derives for the reference).

Data layout in HBM: activations NHWC bf16; conv weights bf16 [Cout][kh][kw][Cin] (a derived cache
of the fp32 OIHW parameters, rebuilt when a parameter's version counter changes); GroupNorm
statistics, time embeddings and everything scalar fp32.  Skip-connection concats are never
materialised (the conv kernel's K loop walks two sources); the ResnetBlock residual add is fused
into the second GroupNorm+SiLU pass; the mid-attention residual into the to_out conv epilogue; the
LinearAttention residual into the output LayerNorm.

Backward: the data gradient of every convolution runs on the SAME tcgen05 implicit-GEMM kernel with
re-laid-out weights (flipped/transposed 3x3; the stride-2 4x4 becomes a parity conv and vice versa);
the weight gradient is a second tcgen05 kernel contracting over pixels; gradient accumulation at
residual / skip joins rides the conv epilogue's `residual` input.  Parameter gradients are written
into one flat fp32 arena (parameter order), so a data-parallel step needs ONE all-reduce and the
optimiser ONE kernel.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Tuple

import torch
from torch import Tensor, nn

from . import native as N


class WeightCache:
    """Derived (re-laid-out / concatenated) copies of parameters, keyed by name, invalidated by version."""

    def __init__(self):
        self._store: Dict[str, Tuple[tuple, Tensor]] = {}
        self.force = False      # True while capturing a CUDA graph: every derivation must be part of the replay

    def get(self, key: str, params: Tuple[Tensor, ...], make) -> Tensor:
        sig = tuple((p.data_ptr(), p._version, p.device) for p in params)
        hit = self._store.get(key)
        if hit is not None and hit[0] == sig and not self.force:
            return hit[1]
        with torch.no_grad():
            val = make(*[p.detach() for p in params])
        self._store[key] = (sig, val)
        return val

    def clear(self):
        self._store.clear()


class Tape:
    """What one training-mode forward keeps for its backward: saved activations by op key."""

    def __init__(self):
        self.saved: Dict[str, tuple] = {}


class GradArena:
    """One flat fp32 buffer holding every parameter gradient, in `module.parameters()` order."""

    def __init__(self, params: List[nn.Parameter], device):
        self.offsets: Dict[int, Tuple[int, torch.Size]] = {}
        off = 0
        for p in params:
            self.offsets[id(p)] = (off, p.shape)
            off += p.numel()
        self.numel = off
        self.flat = torch.zeros((off + 3) // 4 * 4, device=device, dtype=torch.float32)

    def of(self, p: nn.Parameter) -> Tensor:
        off, shape = self.offsets[id(p)]
        return self.flat[off:off + p.numel()].view(shape)


class UnetEngine:
    def __init__(self, unet: nn.Module, fold_upsample: bool = True, ln_eps: float = 1e-5):
        self.m = unet
        self.cache = WeightCache()
        self.fold_upsample = fold_upsample
        self.ln_eps = ln_eps
        self.fuse_linear_attention = True
        # tcgen05 form of the fused block (csrc/attention_tc.cu) where it applies; TEDM_LINATTN_TC=0 keeps the mma.sync kernels (A/B runs)
        self.linear_attention_tc = os.environ.get("TEDM_LINATTN_TC", "1") != "0"
        # res_conv + GroupNorm + SiLU + add of a ResnetBlock's tail in one kernel (forward of inference and training);
        # TEDM_FUSE_RES=0 keeps the two passes
        self.fuse_res_conv = os.environ.get("TEDM_FUSE_RES", "1") != "0"
        # inference: block1's GroupNorm + SiLU applied to block2's conv input in shared memory instead of a pass of its own
        # (bit-identical).  OFF by default: measured on B200 it removes 0.33 ms of GroupNorm passes per step and adds about as
        # much to the convs (the transform's LDS / STS / MUFU compete with the tensor core for shared memory and issue slots):
        # 8.88-8.96 ms against 8.99-9.08 ms, inside the box-to-box noise; TEDM_FUSE_GN=1 turns it on
        self.fuse_gn_into_conv = os.environ.get("TEDM_FUSE_GN", "0") != "0"
        self.fuse_gn_min_bytes = int(os.environ.get("TEDM_FUSE_GN_MIN_MB", "0")) << 20    # A/B: only tensors at least this large
        self.fuse_gn_ws4 = os.environ.get("TEDM_FUSE_GN_WS4", "1") != "0"                # A/B: the four-row kernel's variant
        # backward: weight gradients run on a side stream next to the data-gradient chain (they only meet in the optimiser);
        # at the low-resolution levels neither kernel fills the 148 SMs on its own
        self.overlap_wgrad = True
        self._side: Optional[torch.cuda.Stream] = None
        self._keep: list = []
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
        self.last_grad_arena: Optional[GradArena] = None
        # data parallelism: an object with early(flat_slice) / late(tensors) that sums gradient regions over ranks WHILE the
        # backward is still running (tedm_b200/train.py: GradReducer); None = the caller reduces afterwards
        self.grad_reducer = None
        # every implicit-GEMM conv weight of the net: key -> (parameter, forward mode); re-laid out in ONE launch
        self._wspec: Dict[str, Tuple[nn.Parameter, int]] = {}
        self._wsig = None
        self._wtable = None
        self._wfwd: Optional[Tensor] = None
        self._wdg: Optional[Tensor] = None
        self._wviews: Dict[str, Tuple[Tensor, Optional[Tensor]]] = {}
        self.force_refresh = False      # set while capturing a CUDA graph: the re-layout must be part of every replay
        self._collect_conv_specs()

    # -- conv weight arena ----------------------------------------------------------------------
    def _collect_conv_specs(self) -> None:
        m, spec = self.m, self._wspec

        def resblock(key, rb):
            spec[key + ".block1.proj"] = (rb.block1.proj.weight, N.MODE_3X3)
            spec[key + ".block2.proj"] = (rb.block2.proj.weight, N.MODE_3X3)
            if isinstance(rb.res_conv, nn.Conv2d):
                spec[key + ".res_conv"] = (rb.res_conv.weight, N.MODE_1X1)

        def linattn(key, wrap):
            att = wrap.fn.fn
            spec[key + ".to_qkv"] = (att.to_qkv.weight, N.MODE_1X1)
            spec[key + ".to_out"] = (att.to_out[0].weight, N.MODE_1X1)

        for i, (b1, b2, attn, down) in enumerate(m.downs):
            k = f"downs.{i}"
            resblock(k + ".0", b1)
            resblock(k + ".1", b2)
            linattn(k + ".2", attn)
            spec[k + ".3"] = (down.weight, N.MODE_4X4S2 if down.kernel_size[0] == 4 else N.MODE_3X3)
        resblock("mid_block1", m.mid_block1)
        spec["mid_attn.to_qkv"] = (m.mid_attn.fn.fn.to_qkv.weight, N.MODE_1X1)
        spec["mid_attn.to_out"] = (m.mid_attn.fn.fn.to_out.weight, N.MODE_1X1)
        resblock("mid_block2", m.mid_block2)
        for i, (b1, b2, attn, up) in enumerate(m.ups):
            k = f"ups.{i}"
            resblock(k + ".0", b1)
            resblock(k + ".1", b2)
            linattn(k + ".2", attn)
            if isinstance(up, nn.Sequential):
                spec[k + ".3.1"] = (up[1].weight, N.MODE_UP3X3)
            else:
                spec[k + ".3"] = (up.weight, N.MODE_3X3)
        resblock("final_res_block", m.final_res_block)

    def _ensure_weights(self, train: bool) -> None:
        """bf16 operand copies of every conv weight (forward layout; + data-gradient layout when training), refreshed
        by one kernel launch whenever any parameter changed."""
        params = [p for p, _ in self._wspec.values()]
        sig = tuple((p.data_ptr(), p._version) for p in params)
        have_dg = self._wdg is not None
        if sig == self._wsig and (have_dg or not train) and not self.force_refresh:
            return
        ptrs = tuple(p.data_ptr() for p in params)
        if self._wtable is None or self._wtable[0] != ptrs or (train and not have_dg):
            dev = params[0].device
            taps_f = {N.MODE_1X1: 1, N.MODE_3X3: 9, N.MODE_4X4S2: 16, N.MODE_UP3X3: 16}
            n_f = sum(p.numel() // {N.MODE_1X1: 1, N.MODE_3X3: 9, N.MODE_4X4S2: 16, N.MODE_UP3X3: 9}[md] * taps_f[md]
                      for p, md in self._wspec.values())
            if self._wfwd is None or self._wfwd.device != dev:
                self._wfwd = torch.empty(n_f, device=dev, dtype=torch.bfloat16)
                self._wdg = None
            if train and self._wdg is None:
                self._wdg = torch.empty(n_f, device=dev, dtype=torch.bfloat16)
            table = (N.WeightEntry * len(self._wspec))()
            off, cta = 0, 0
            self._wviews = {}
            for i, (key, (p, md)) in enumerate(self._wspec.items()):
                if p.dtype != torch.float32 or not p.is_contiguous() or p.device != dev:
                    raise TypeError(f"{key}: conv weights must be contiguous fp32 on one device")
                cout, cin = p.shape[0], p.shape[1]
                if cout % 32 or cin % 32:
                    raise RuntimeError(f"{key}: channel counts ({cout}, {cin}) must be multiples of 32")
                n = cout * cin * taps_f[md]
                fwd = self._wfwd[off:off + n]
                dg = self._wdg[off:off + n] if self._wdg is not None else None
                self._wviews[key] = (fwd, dg)
                table[i] = N.WeightEntry(p.data_ptr(), fwd.data_ptr(), dg.data_ptr() if dg is not None else None, cout, cin,
                                         md, cta)
                off += n
                cta += (cout // 32) * (cin // 32)
            raw = torch.frombuffer(bytearray(bytes(table)), dtype=torch.uint8).to(dev)
            self._wtable = (ptrs, raw, len(self._wspec), cta)
        _, raw, n_entries, total_ctas = self._wtable
        N.prepare_weights(raw, n_entries, total_ctas)
        self._wsig = sig

    def _w(self, key: str) -> Tensor:
        return self._wviews[key][0]

    def _wd(self, key: str) -> Tensor:
        return self._wviews[key][1]

    # -- derived weights ------------------------------------------------------------------------
    def _f32(self, p: Tensor) -> Tensor:
        return p.detach().float().contiguous() if (p.dtype != torch.float32 or not p.is_contiguous()) else p.detach()

    def _krsc(self, key: str, w: Tensor) -> Tensor:
        return self.cache.get(key + ":krsc", (w,), lambda a: N.weight_to_krsc(a.float()))

    def _folded(self, key: str, w: Tensor) -> Tensor:
        return self.cache.get(key + ":fold", (w,), lambda a: N.fold_upsample_weight(a.float()))

    def _dgrad_w(self, key: str, w: Tensor, mode: int) -> Tensor:
        return self.cache.get(key + ":dgrad", (w,), lambda a: N.weight_to_dgrad(a.float(), mode))

    def _time_cat(self) -> Tuple[Tensor, Tensor]:
        ws = tuple(rb.time_mlp[1].weight for rb in self._resblocks)
        bs = tuple(rb.time_mlp[1].bias for rb in self._resblocks)
        w = self.cache.get("time_cat:w", ws, lambda *a: torch.cat([x.float() for x in a], dim=0).contiguous())
        b = self.cache.get("time_cat:b", bs, lambda *a: torch.cat([x.float() for x in a], dim=0).contiguous())
        return w, b

    # -- blocks -----------------------------------------------------------------------------------
    def _block(self, key: str, blk: nn.Module, x0: Tensor, x1: Optional[Tensor], ss, ss_off: int,
               residual: Optional[Tensor], tape: Optional[Tape] = None) -> Tensor:
        cout = blk.proj.weight.shape[0]
        h, part = N.conv_igemm(x0, self._w(key + ".proj"), N.MODE_3X3, cout,
                               bias=self._f32(blk.proj.bias), src1=x1, gn_groups=blk.norm.num_groups)
        if tape is not None:
            tape.saved[key] = (x0, x1, h, part)
        return N.gn_silu(h, part, self._f32(blk.norm.weight), self._f32(blk.norm.bias), blk.norm.num_groups,
                         eps=blk.norm.eps, scale_shift=ss, ss_offset=ss_off, residual=residual)

    def _resblock(self, key: str, rb: nn.Module, x0: Tensor, x1: Optional[Tensor], tproj: Optional[Tensor],
                  tape: Optional[Tape] = None) -> Tensor:
        ss_off = self._ss_offset[id(rb)]
        b2 = rb.block2
        c_mid, c_out = b2.proj.weight.shape[1], b2.proj.weight.shape[0]
        if (tape is None and self.fuse_gn_into_conv and N.conv_src_affine_supported(x0.shape[1], x0.shape[2], c_mid, c_out)
                and x0.shape[0] * x0.shape[1] * x0.shape[2] * c_mid * 2 >= self.fuse_gn_min_bytes
                and (self.fuse_gn_ws4 or not (x0.shape[2] >= 128 and c_out == 64))):
            # inference: block1's GroupNorm + scale/shift + SiLU is applied to block2's halo boxes in shared memory
            b1 = rb.block1
            h1, part1 = N.conv_igemm(x0, self._w(key + ".block1.proj"), N.MODE_3X3, c_mid, bias=self._f32(b1.proj.bias), src1=x1,
                                     gn_groups=b1.norm.num_groups)
            aff1 = N.gn_affine(part1, self._f32(b1.norm.weight), self._f32(b1.norm.bias), b1.norm.num_groups,
                               h1.shape[1] * h1.shape[2], eps=b1.norm.eps, scale_shift=tproj, ss_offset=ss_off)
            h2, part2 = N.conv_igemm(h1, self._w(key + ".block2.proj"), N.MODE_3X3, c_out, bias=self._f32(b2.proj.bias),
                                     gn_groups=b2.norm.num_groups, src0_affine=aff1)
            if isinstance(rb.res_conv, nn.Conv2d) and self.fuse_res_conv:
                aff2 = N.gn_affine(part2, self._f32(b2.norm.weight), self._f32(b2.norm.bias), b2.norm.num_groups,
                                   h2.shape[1] * h2.shape[2], eps=b2.norm.eps)
                return N.conv_igemm(x0, self._w(key + ".res_conv"), N.MODE_1X1, c_out, bias=self._f32(rb.res_conv.bias), src1=x1,
                                    residual=h2, residual_affine=aff2)
            if isinstance(rb.res_conv, nn.Conv2d):
                res = N.conv_igemm(x0, self._w(key + ".res_conv"), N.MODE_1X1, c_out, bias=self._f32(rb.res_conv.bias), src1=x1)
            else:
                res = x0
            return N.gn_silu(h2, part2, self._f32(b2.norm.weight), self._f32(b2.norm.bias), b2.norm.num_groups, eps=b2.norm.eps,
                             residual=res)
        h = self._block(key + ".block1", rb.block1, x0, x1, tproj, ss_off, None, tape)
        if isinstance(rb.res_conv, nn.Conv2d) and self.fuse_res_conv and h.shape[1] * h.shape[2] >= 128:
            # block2's GroupNorm + SiLU and the residual add happen in the epilogue of the 1x1 res_conv, which reads block2's
            # raw conv output as its residual -- res_conv's own output is never written.  Training too: the backward needs
            # block2's input, raw output and statistics (saved below), never res_conv's output
            blk, cout = rb.block2, rb.res_conv.weight.shape[0]
            h2, part = N.conv_igemm(h, self._w(key + ".block2.proj"), N.MODE_3X3, cout, bias=self._f32(blk.proj.bias),
                                    gn_groups=blk.norm.num_groups)
            if tape is not None:
                tape.saved[key + ".block2"] = (h, None, h2, part)
            aff = N.gn_affine(part, self._f32(blk.norm.weight), self._f32(blk.norm.bias), blk.norm.num_groups,
                              h2.shape[1] * h2.shape[2], eps=blk.norm.eps)
            return N.conv_igemm(x0, self._w(key + ".res_conv"), N.MODE_1X1, cout, bias=self._f32(rb.res_conv.bias), src1=x1,
                                residual=h2, residual_affine=aff)
        if isinstance(rb.res_conv, nn.Conv2d):
            cout = rb.res_conv.weight.shape[0]
            res = N.conv_igemm(x0, self._w(key + ".res_conv"), N.MODE_1X1, cout,
                               bias=self._f32(rb.res_conv.bias), src1=x1)
        else:
            if x1 is not None:
                raise RuntimeError("identity residual with a two-source input")
            res = x0
        return self._block(key + ".block2", rb.block2, h, None, None, 0, res, tape)

    def _linear_attention(self, key: str, wrap: nn.Module, x: Tensor, tape: Optional[Tape] = None) -> Tensor:
        pre, att = wrap.fn.norm, wrap.fn.fn
        c = x.shape[-1]
        n_pix = x.shape[1] * x.shape[2]
        if tape is None and self.fuse_linear_attention and self.linear_attention_tc and N.linear_attention_tc_supported(
                n_pix, c, att.heads, att.dim_head):
            # the same block with every GEMM on tcgen05; its softmax over pixels uses a weight-only bound of the k logits
            # instead of a running maximum, so blocks with extreme projection weights stay on the mma.sync kernels below
            wg, shift_log2, bound = self.cache.get(key + ":tcw", (att.to_qkv.weight, pre.g),
                                                   lambda w, g: N.linear_attention_tc_weights(w, g, att.heads, att.dim_head))
            if bound <= N.LINATTN_TC_MAX_SHIFT:
                return N.linear_attention_block_tc(x, wg, shift_log2, self._w(key + ".to_out"), self._f32(att.to_out[0].bias),
                                                   self._f32(att.to_out[1].g).reshape(-1), att.heads, att.dim_head, att.scale,
                                                   self.ln_eps)
        if (tape is None and self.fuse_linear_attention
                and N.linear_attention_fused_supported(n_pix, c, att.heads, att.dim_head)):
            # inference at the high-resolution levels: the whole block in 3 launches, q/k/v never leave the SM
            return N.linear_attention_block_fused(x, self._w(key + ".to_qkv"), self._f32(pre.g).reshape(-1),
                                                  self._w(key + ".to_out"), self._f32(att.to_out[0].bias),
                                                  self._f32(att.to_out[1].g).reshape(-1), att.heads, att.dim_head,
                                                  att.scale, self.ln_eps)
        y = N.layernorm(x, self._f32(pre.g).reshape(-1), eps=self.ln_eps)
        qkv = N.conv_igemm(y, self._w(key + ".to_qkv"), N.MODE_1X1, att.to_qkv.weight.shape[0])
        if tape is not None:
            o, ws = N.linear_attention(qkv, att.heads, att.dim_head, att.scale, want_workspace=True)
        else:
            o = N.linear_attention(qkv, att.heads, att.dim_head, att.scale)
        o2 = N.conv_igemm(o, self._w(key + ".to_out"), N.MODE_1X1, c,
                          bias=self._f32(att.to_out[0].bias))
        if tape is not None:
            tape.saved[key] = (x, y, qkv, o, o2, ws)
        return N.layernorm(o2, self._f32(att.to_out[1].g).reshape(-1), eps=self.ln_eps, residual=x)

    def _mid_attention(self, key: str, wrap: nn.Module, x: Tensor, tape: Optional[Tape] = None) -> Tensor:
        pre, att = wrap.fn.norm, wrap.fn.fn
        c = x.shape[-1]
        y = N.layernorm(x, self._f32(pre.g).reshape(-1), eps=self.ln_eps)
        qkv = N.conv_igemm(y, self._w(key + ".to_qkv"), N.MODE_1X1, att.to_qkv.weight.shape[0])
        o = N.attention(qkv, att.heads, att.dim_head, float(att.scale))
        if tape is not None:
            tape.saved[key] = (x, y, qkv, o)
        return N.conv_igemm(o, self._w(key + ".to_out"), N.MODE_1X1, c,
                            bias=self._f32(att.to_out.bias), residual=x)

    def _time_projection(self, x: Tensor, timestep: Optional[Tensor], tape: Optional[Tape], time_key) -> Optional[Tensor]:
        """(B, sum of 2*Cout over the ResnetBlocks) fp32: every block's (scale, shift) from ONE batched projection of the
        time embedding (unet_model.py:76-93, 287-292, 151, 163-166); None without a timestep."""
        m = self.m
        tproj = None
        if timestep is not None and getattr(m, "learned_sinusoidal_cond", False):
            # learned-frequency embedding (unet_model.py:96-114): never enabled by a reference entry point; its (B, 17) ->
            # (B, 256) MLP stays on the host-side torch path, everything downstream is native
            if tape is not None:
                raise NotImplementedError("training with learned_sinusoidal_cond=True is not part of the B200 hot path")
            with torch.no_grad():
                temb = m.time_mlp(timestep.detach().to(x.device))
            tproj = N.time_proj(temb.float().contiguous(), *self._time_cat())
        elif timestep is not None:
            t = timestep.detach().to(device=x.device, dtype=torch.int64).contiguous()
            pe = m.time_mlp[0]
            freq = self.cache.get("freq", (m.time_mlp[1].weight,), lambda w: pe.frequencies(w.device).float().contiguous())
            tw = (self._f32(m.time_mlp[1].weight), self._f32(m.time_mlp[1].bias),
                  self._f32(m.time_mlp[3].weight), self._f32(m.time_mlp[3].bias))
            if tape is not None:
                emb, hid, temb = N.time_embed_train(t, freq, *tw)
                tape.saved["time"] = (emb, hid, temb)
                wcat, bcat = self._time_cat()
                tproj = N.time_proj(temb, wcat, bcat)
            elif time_key is not None:
                # the caller vouches that `timestep` holds the same values whenever it passes this key (TEDM's fixed step
                # list): the whole time path is a function of the weights alone and is computed once per weight version
                tparams = (m.time_mlp[1].weight, m.time_mlp[1].bias, m.time_mlp[3].weight, m.time_mlp[3].bias,
                           *(rb.time_mlp[1].weight for rb in self._resblocks), *(rb.time_mlp[1].bias for rb in self._resblocks))
                tproj = self.cache.get(f"tproj:{time_key}", tparams,
                                       lambda *_: N.time_proj(N.time_embed(t, freq, *tw), *self._time_cat()))
            else:
                wcat, bcat = self._time_cat()
                tproj = N.time_proj(N.time_embed(t, freq, *tw), wcat, bcat)

        return tproj

    @staticmethod
    def _fire(mod: nn.Module, ins, out: Tensor) -> None:
        """Forward hooks registered on a submodule (the reference's DatasetDM hooks `ups[i][2]`, datasetDM_model.py:50-53)
        see what they would see in the reference: NCHW fp32 input(s) and output.  The engine does not go through the
        submodules' own forward, so it calls the hooks itself; a hook may observe, not replace, the output."""
        if not mod._forward_hooks:
            return
        xin = [N.nhwc_to_nchw_f32(i) for i in ins]                 # a decoder block's input is cat(h, skip)
        args = (torch.cat(xin, dim=1) if len(xin) > 1 else xin[0],)
        o = N.nhwc_to_nchw_f32(out)
        for hook in list(mod._forward_hooks.values()):
            if hook(mod, args, o) is not None:
                raise NotImplementedError("forward hooks that replace a submodule's output are not supported by the fused engine")

    # -- whole network ----------------------------------------------------------------------------
    def forward(self, x: Tensor, timestep: Optional[Tensor], want_features: bool = False, skip_tail: bool = False,
                tape: Optional[Tape] = None, time_key=None):
        m = self.m
        if not x.is_cuda:
            raise RuntimeError("tedm_b200.Unet runs on CUDA (sm_100a) only; there is no CPU fallback")
        if x.dim() != 4 or x.shape[1] != m.channels:
            raise ValueError(f"expected input (B, {m.channels}, H, W), got {tuple(x.shape)}")
        x = x.detach().float().contiguous()
        self._ensure_weights(train=tape is not None)
        tproj = self._time_projection(x, timestep, tape, time_key)

        h = N.stem_conv7x7(x, self._f32(m.init_conv.weight), self._f32(m.init_conv.bias))
        stem = h
        skips: List[Tensor] = []
        for i, (b1, b2, attn, down) in enumerate(m.downs):
            k = f"downs.{i}"
            hin = h
            h = self._resblock(k + ".0", b1, h, None, tproj, tape)
            self._fire(b1, [hin], h)
            skips.append(h)
            hin = h
            h = self._resblock(k + ".1", b2, h, None, tproj, tape)
            self._fire(b2, [hin], h)
            hin = h
            h = self._linear_attention(k + ".2", attn, h, tape)
            self._fire(attn, [hin], h)
            skips.append(h)
            mode = N.MODE_4X4S2 if down.kernel_size[0] == 4 else N.MODE_3X3
            if tape is not None:
                tape.saved[k + ".3"] = (h,)
            h = N.conv_igemm(h, self._w(k + ".3"), mode, down.weight.shape[0], bias=self._f32(down.bias))
        hin = h
        h = self._resblock("mid_block1", m.mid_block1, h, None, tproj, tape)
        self._fire(m.mid_block1, [hin], h)
        hin = h
        h = self._mid_attention("mid_attn", m.mid_attn, h, tape)
        self._fire(m.mid_attn, [hin], h)
        hin = h
        h = self._resblock("mid_block2", m.mid_block2, h, None, tproj, tape)
        self._fire(m.mid_block2, [hin], h)
        feats: List[Tensor] = []
        n_up = len(m.ups)
        for i, (b1, b2, attn, up) in enumerate(m.ups):
            k = f"ups.{i}"
            hin, sk = h, skips.pop()
            h = self._resblock(k + ".0", b1, h, sk, tproj, tape)
            self._fire(b1, [hin, sk], h)
            hin, sk = h, skips.pop()
            h = self._resblock(k + ".1", b2, h, sk, tproj, tape)
            self._fire(b2, [hin, sk], h)
            hin = h
            h = self._linear_attention(k + ".2", attn, h, tape)
            self._fire(attn, [hin], h)
            feats.append(h)
            if skip_tail and i == n_up - 1:
                return None, feats
            if tape is not None:
                tape.saved[k + ".3"] = (h,)
            if isinstance(up, nn.Sequential):          # Upsample: nearest x2 + 3x3 conv
                conv = up[1]
                if self.fold_upsample:
                    h = N.conv_igemm(h, self._w(k + ".3.1"), N.MODE_UP3X3, conv.weight.shape[0],
                                     bias=self._f32(conv.bias))
                else:
                    h = N.conv_igemm(N.upsample2x(h), self._krsc(k + ".3.1", conv.weight), N.MODE_3X3,
                                     conv.weight.shape[0], bias=self._f32(conv.bias))
            else:
                h = N.conv_igemm(h, self._w(k + ".3"), N.MODE_3X3, up.weight.shape[0], bias=self._f32(up.bias))
        h = self._resblock("final_res_block", m.final_res_block, h, stem, tproj, tape)
        out = N.final_conv1x1(h, self._f32(m.final_conv.weight).reshape(m.out_dim, -1), self._f32(m.final_conv.bias))
        if tape is not None:
            tape.saved["io"] = (x, tproj, h)
        return (out, feats) if want_features else out

    # =============================================================================================
    # backward
    # =============================================================================================
    def _conv_bwd(self, key: str, conv: nn.Conv2d, mode: int, x0: Tensor, x1: Optional[Tensor], dy: Tensor, G: GradArena,
                  add0: Optional[Tensor] = None, add1: Optional[Tensor] = None, bias_grad: bool = True,
                  need_dx: bool = True):
        """Weight (+bias) gradient of conv(x0[, x1]) and its data gradient(s); add0/add1 are gradients already
        flowing into x0/x1 from elsewhere (fused through the conv epilogue's residual input)."""
        cout = conv.weight.shape[0]
        c0 = x0.shape[-1]
        c1 = x1.shape[-1] if x1 is not None else 0
        if self.overlap_wgrad:
            if self._side is None:
                self._side = torch.cuda.Stream(device=dy.device)
            self._side.wait_stream(torch.cuda.current_stream())          # dy, x0 and the zeroed arena are ready
            with torch.cuda.stream(self._side):
                N.conv_wgrad(x0, dy, mode, src1=x1, grad_oihw=G.of(conv.weight))
                if bias_grad and conv.bias is not None:
                    N.bias_grad(dy, G.of(conv.bias))
            self._keep.append((x0, x1, dy))       # freed only after the join: the allocator must not recycle them early
        else:
            N.conv_wgrad(x0, dy, mode, src1=x1, grad_oihw=G.of(conv.weight))
            if bias_grad and conv.bias is not None:
                N.bias_grad(dy, G.of(conv.bias))
        if not need_dx:
            return None, None
        wd = self._wd(key)
        run_mode = {N.MODE_1X1: N.MODE_1X1, N.MODE_3X3: N.MODE_3X3, N.MODE_4X4S2: N.MODE_UP3X3,
                    N.MODE_UP3X3: N.MODE_4X4S2}[mode]
        taps = {N.MODE_1X1: 1, N.MODE_3X3: 9, N.MODE_4X4S2: 16, N.MODE_UP3X3: 16}[mode]
        if c1 == 0:
            return N.conv_igemm(dy, wd, run_mode, c0, residual=add0), None
        if mode not in (N.MODE_1X1, N.MODE_3X3):
            raise RuntimeError("two-source convolutions are 1x1 or 3x3")
        # one launch computes both gradients: N tiles over c0 + c1 channels, 64-channel sub-tiles routed to two tensors
        return N.conv_igemm(dy, wd, run_mode, c0 + c1, residual=add0, split=c0, residual2=add1)

    def _block_bwd(self, key: str, blk: nn.Module, dy: Tensor, tape: Tape, G: GradArena, ss, ss_off: int, dss,
                   add0: Optional[Tensor] = None, add1: Optional[Tensor] = None):
        x0, x1, h, part = tape.saved[key]
        dh = N.gn_silu_bwd(h, dy, part, self._f32(blk.norm.weight), self._f32(blk.norm.bias), blk.norm.num_groups,
                           G.of(blk.norm.weight), G.of(blk.norm.bias), G.of(blk.proj.bias), eps=blk.norm.eps,
                           scale_shift=ss, ss_offset=ss_off, dscale_shift=dss if ss is not None else None)
        return self._conv_bwd(key + ".proj", blk.proj, N.MODE_3X3, x0, x1, dh, G, add0, add1, bias_grad=False)

    def _resblock_bwd(self, key: str, rb: nn.Module, dout: Tensor, tape: Tape, G: GradArena, tproj, dss):
        """out = block2(block1(x)) + res(x)  ->  (d x0, d x1)."""
        ss_off = self._ss_offset[id(rb)]
        da1, _ = self._block_bwd(key + ".block2", rb.block2, dout, tape, G, None, 0, None)
        x0, x1 = tape.saved[key + ".block1"][:2]
        if isinstance(rb.res_conv, nn.Conv2d):
            r0, r1 = self._conv_bwd(key + ".res_conv", rb.res_conv, N.MODE_1X1, x0, x1, dout, G)
        else:
            r0, r1 = dout, None
        return self._block_bwd(key + ".block1", rb.block1, da1, tape, G, tproj, ss_off, dss, add0=r0, add1=r1)

    def _linear_attention_bwd(self, key: str, wrap: nn.Module, dout: Tensor, tape: Tape, G: GradArena) -> Tensor:
        pre, att = wrap.fn.norm, wrap.fn.fn
        x, y, qkv, o, o2, ws = tape.saved[key]
        do2 = N.layernorm_bwd(o2, self._f32(att.to_out[1].g).reshape(-1), dout, G.of(att.to_out[1].g).view(-1), eps=self.ln_eps)
        do, _ = self._conv_bwd(key + ".to_out", att.to_out[0], N.MODE_1X1, o, None, do2, G)
        dqkv = N.linear_attention_bwd(qkv, do, ws, att.heads, att.dim_head, att.scale)
        dy, _ = self._conv_bwd(key + ".to_qkv", att.to_qkv, N.MODE_1X1, y, None, dqkv, G)
        return N.layernorm_bwd(x, self._f32(pre.g).reshape(-1), dy, G.of(pre.g).view(-1), eps=self.ln_eps, add=dout)

    def _mid_attention_bwd(self, key: str, wrap: nn.Module, dout: Tensor, tape: Tape, G: GradArena) -> Tensor:
        pre, att = wrap.fn.norm, wrap.fn.fn
        x, y, qkv, o = tape.saved[key]
        do, _ = self._conv_bwd(key + ".to_out", att.to_out, N.MODE_1X1, o, None, dout, G)
        dqkv = N.attention_bwd(qkv, do, att.heads, att.dim_head, float(att.scale), o=o)
        dy, _ = self._conv_bwd(key + ".to_qkv", att.to_qkv, N.MODE_1X1, y, None, dqkv, G)
        return N.layernorm_bwd(x, self._f32(pre.g).reshape(-1), dy, G.of(pre.g).view(-1), eps=self.ln_eps, add=dout)

    def backward(self, tape: Tape, dout: Tensor) -> GradArena:
        """Given d loss / d output (B, out_dim, H, W) fp32, fill a fresh gradient arena (parameter order)."""
        m = self.m
        params = list(m.parameters())
        G = GradArena(params, dout.device)
        x, tproj, hfinal = tape.saved["io"]
        dss = torch.empty_like(tproj) if tproj is not None else None
        dout = dout.detach().float().contiguous()
        dh = N.final_conv1x1_bwd(hfinal, self._f32(m.final_conv.weight).reshape(m.out_dim, -1), dout,
                                 G.of(m.final_conv.weight).view(m.out_dim, -1), G.of(m.final_conv.bias))
        dh, dstem_skip = self._resblock_bwd("final_res_block", m.final_res_block, dh, tape, G, tproj, dss)
        dskips: List[Tensor] = []
        n_up = len(m.ups)
        for i in reversed(range(n_up)):
            b1, b2, attn, up = m.ups[i]
            k = f"ups.{i}"
            (hin,) = tape.saved[k + ".3"]
            if isinstance(up, nn.Sequential):
                if not self.fold_upsample:
                    raise RuntimeError("training runs with the folded upsample conv")
                dh, _ = self._conv_bwd(k + ".3.1", up[1], N.MODE_UP3X3, hin, None, dh, G)
            else:
                dh, _ = self._conv_bwd(k + ".3", up, N.MODE_3X3, hin, None, dh, G)
            dh = self._linear_attention_bwd(k + ".2", attn, dh, tape, G)
            dh, ds2 = self._resblock_bwd(k + ".1", b2, dh, tape, G, tproj, dss)
            dh, ds1 = self._resblock_bwd(k + ".0", b1, dh, tape, G, tproj, dss)
            # forward popped the skips in the order (b1: after-attention map, b2: after-block1 map)
            dskips.append(ds2)
            dskips.append(ds1)
        dh, _ = self._resblock_bwd("mid_block2", m.mid_block2, dh, tape, G, tproj, dss)
        dh = self._mid_attention_bwd("mid_attn", m.mid_attn, dh, tape, G)
        dh, _ = self._resblock_bwd("mid_block1", m.mid_block1, dh, tape, G, tproj, dss)
        tail_lo = None
        if self.grad_reducer is not None:
            # every parameter from `ups` to the end of the arena (decoder, mid blocks, output convs) has its gradient now,
            # except the per-block time projections, which are reduced separately below: start summing that region over
            # the ranks while the encoder's backward runs.  Weight gradients live on the side stream, so the reducer's
            # stream has to wait for both.
            tail_lo = G.offsets[id(next(m.ups.parameters()))][0]
            self.grad_reducer.early(G.flat[tail_lo:], self._side if self._keep else None)
        level_hi = tail_lo
        for i in reversed(range(len(m.downs))):
            b1, b2, attn, down = m.downs[i]
            k = f"downs.{i}"
            (hin,) = tape.saved[k + ".3"]
            d_attn_skip = dskips.pop()      # gradient of the after-attention map from the decoder
            d_b1_skip = dskips.pop()        # gradient of the after-block1 map from the decoder
            mode = N.MODE_4X4S2 if down.kernel_size[0] == 4 else N.MODE_3X3
            dh, _ = self._conv_bwd(k + ".3", down, mode, hin, None, dh, G, add0=d_attn_skip)
            dh = self._linear_attention_bwd(k + ".2", attn, dh, tape, G)
            dh, _ = self._resblock_bwd(k + ".1", b2, dh, tape, G, tproj, dss)
            dh = N.add_bf16(dh, d_b1_skip)
            dh, _ = self._resblock_bwd(k + ".0", b1, dh, tape, G, tproj, dss)
            if self.grad_reducer is not None:
                # this encoder level's slice of the arena is complete (its time projections aside): the deepest level,
                # which holds most of the encoder's parameters, finishes first
                lo = G.offsets[id(next(m.downs[i].parameters()))][0]
                self.grad_reducer.early(G.flat[lo:level_hi], self._side if self._keep else None)
                level_hi = lo
        dstem = N.add_bf16(dh, dstem_skip)
        N.stem_conv7x7_wgrad(x, dstem, G.of(m.init_conv.weight), G.of(m.init_conv.bias))
        tgrads = self._time_bwd(tape, dss, G) if tproj is not None else None
        if self._side is not None and self._keep:
            torch.cuda.current_stream().wait_stream(self._side)            # join: every weight gradient has landed
        self._keep.clear()
        if self.grad_reducer is not None:
            # what is left: the stem and the time MLP at the head of the arena + the concatenated time-projection gradients;
            # then (after every region has been summed) the time projections are scattered to their owners in the arena
            late = [G.flat[:level_hi]] + ([tgrads[0].view(-1), tgrads[1]] if tgrads is not None else [])
            self.grad_reducer.late(late)
        if tgrads is not None:
            self._time_scatter(tgrads, G)
        self.last_grad_arena = G
        return G

    def _time_bwd(self, tape: Tape, dss: Tensor, G: GradArena):
        """Backward of the time MLPs; returns the gradients of the CONCATENATED per-block projections (weight, bias) for
        `_time_scatter` -- kept apart so that a data-parallel reducer can sum them before they are scattered."""
        m = self.m
        emb, hid, temb = tape.saved["time"]
        wcat, _ = self._time_cat()
        total, tdim = wcat.shape
        dwcat = torch.zeros(total, tdim, device=dss.device, dtype=torch.float32)
        dbcat = torch.zeros(total, device=dss.device, dtype=torch.float32)
        d3 = N.linear_bwd(dss, None, N.ACT_NONE, temb, N.ACT_SILU, wcat, dwcat, dbcat)
        d2 = N.linear_bwd(d3, temb, N.ACT_SILU, hid, N.ACT_GELU, self._f32(m.time_mlp[3].weight),
                          G.of(m.time_mlp[3].weight), G.of(m.time_mlp[3].bias))
        N.linear_bwd(d2, hid, N.ACT_GELU, emb, N.ACT_NONE, self._f32(m.time_mlp[1].weight),
                     G.of(m.time_mlp[1].weight), G.of(m.time_mlp[1].bias), want_dx=False)
        return dwcat, dbcat

    def _time_scatter(self, tgrads, G: GradArena) -> None:
        dwcat, dbcat = tgrads
        off = 0
        for rb in self._resblocks:      # scatter the concatenated projection gradient back to its owners
            lin = rb.time_mlp[1]
            n = lin.weight.shape[0]
            G.of(lin.weight).copy_(dwcat[off:off + n])
            G.of(lin.bias).copy_(dbcat[off:off + n])
            off += n


class UnetFunction(torch.autograd.Function):
    """Autograd node for the whole UNet: forward and backward are both native kernel schedules."""

    @staticmethod
    def forward(ctx, engine: UnetEngine, x: Tensor, timestep: Optional[Tensor], *params: Tensor) -> Tensor:
        tape = Tape()
        out = engine.forward(x, timestep, tape=tape)
        ctx.engine, ctx.tape, ctx.params = engine, tape, params
        return out

    @staticmethod
    def backward(ctx, dout: Tensor):
        G = ctx.engine.backward(ctx.tape, dout)
        ctx.tape = None
        grads = tuple(G.of(p) if p.requires_grad else None for p in ctx.params)
        return (None, None, None) + grads
