"""ctypes binding of libtedm_b200.so -- the C ABI declared in include/tedm_b200.h.

PyTorch owns every tensor; this module only passes raw device pointers, sizes and the current
CUDA stream.  There is no fallback: if the library is missing or a call fails, a RuntimeError
with tedm_last_error() is raised.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TEDM_B200_LIB") or os.path.join(_HERE, "lib", "libtedm_b200.so")   # override: A/B builds

_p, _i, _f, _i64 = C.c_void_p, C.c_int, C.c_float, C.c_int64


class ConvArgs(C.Structure):  # tedm_conv_args
    _fields_ = [("src0", _p), ("src1", _p), ("weight", _p), ("bias", _p), ("residual", _p), ("out", _p),
                ("gn_partial", _p), ("batch", _i), ("height", _i), ("width", _i), ("c0", _i), ("c1", _i),
                ("cout", _i), ("mode", _i), ("gn_groups", _i), ("out_dtype", _i), ("src0_image_stride", _i64),
                ("src1_image_stride", _i64), ("out_image_stride", _i64), ("split", _i), ("out2", _p), ("residual2", _p),
                ("n_extra", _i), ("extra_src", _p * 4), ("extra_c", _i * 4), ("extra_image_stride", _i64 * 4),
                ("extra_center", _i * 4), ("residual_affine", _p), ("src0_affine", _p)]


class WeightEntry(C.Structure):  # tedm_weight_entry
    _fields_ = [("w", _p), ("fwd", _p), ("dgrad", _p), ("cout", _i), ("cin", _i), ("mode", _i), ("cta_begin", _i)]


class HeadArgs(C.Structure):  # tedm_head_args
    _fields_ = [("g", _p * 4), ("g_dtype", _i), ("shift", _i * 4), ("n_levels", _i), ("n_sum", _i), ("n_img", _i), ("height", _i),
                ("width", _i), ("c1", _i), ("c2", _i), ("b1", _p), ("bn1_a", _p), ("bn1_c", _p), ("w2", _p),
                ("b2", _p), ("bn2_a", _p), ("bn2_c", _p), ("w3", _p), ("b3", _f), ("logits", _p), ("f_full", _p), ("w1_full", _p), ("c_full", _i), ("exact", _i)]


# name -> (restype, argtypes); must list every symbol include/tedm_b200.h declares
SIGNATURES = {
    "tedm_version": (_i, []),
    "tedm_last_error": (C.c_char_p, []),
    "tedm_q_sample": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p]),
    "tedm_l1_loss": (_i, [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _p]),
    "tedm_sampler_step": (_i, [_p, _p, _p, _p, _p, _p, _f, _f, _f, _f, _f, _i, _f, _i, _i, _p]),
    "tedm_sampler_step_dev": (_i, [_p, _p, _p, _p, _p, _p, _p, _i, _f, _i, _i, _p]),
    "tedm_time_embed": (_i, [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _p]),
    "tedm_time_proj": (_i, [_p, _p, _p, _p, _i, _i, _i, _p]),
    "tedm_stem_conv7x7": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _p]),
    "tedm_conv_igemm_fwd": (_i, [C.POINTER(ConvArgs), _p]),
    "tedm_conv_igemm_wgrad": (_i, [C.POINTER(ConvArgs), _p, _p, _i, _p, _p]),
    "tedm_conv_igemm_wgrad_workspace": (_i64, []),
    "tedm_prepare_weights": (_i, [_p, _i, _i, _p]),
    "tedm_conv_gn_parts": (_i, [_i, _i]),
    "tedm_conv_src_affine_supported": (_i, [_i, _i, _i, _i]),
    "tedm_gn_affine": (_i, [_p, _i, _p, _p, _p, _i, _i, _p, _i, _i, _i, _i, _f, _p]),
    "tedm_conv_set_tile_n": (_i, [_i]),
    "tedm_conv_set_ws": (_i, [_i]),
    "tedm_conv_set_halo": (_i, [_i]),
    "tedm_conv_set_wgrad_halo": (_i, [_i]),
    "tedm_conv_set_deterministic": (_i, [_i]),
    "tedm_conv_set_cta_pairs": (_i, [_i]),
    "tedm_weight_to_krsc": (_i, [_p, _p, _i, _i, _i, _i, _p]),
    "tedm_fold_upsample_weight": (_i, [_p, _p, _i, _i, _p]),
    "tedm_gn_silu_fwd": (_i, [_p, _p, _i, _p, _p, _p, _i, _i, _p, _p, _i, _i, _i, _i, _f, _p]),
    "tedm_layernorm_fwd": (_i, [_p, _p, _p, _p, _i64, _i, _f, _p]),
    "tedm_linear_attention_workspace": (_i64, [_i, _i, _i, _i]),
    "tedm_linear_attention_fwd": (_i, [_p, _p, _p, _i, _i, _i, _i, _f, _p]),
    "tedm_attention_fwd": (_i, [_p, _p, _i, _i, _i, _i, _f, _p]),
    "tedm_linear_attention_fused_supported": (_i, [_i, _i, _i, _i]),
    "tedm_linear_attention_fused_workspace": (_i64, [_i, _i]),
    "tedm_linear_attention_fused_fwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _f, _f, _p]),
    "tedm_linear_attention_tc_supported": (_i, [_i, _i, _i, _i]),
    "tedm_linear_attention_tc_workspace": (_i64, [_i, _i, _i]),
    "tedm_linear_attention_tc_fwd": (_i, [_p, _p, _f, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _f, _f, _p]),
    "tedm_upsample2x": (_i, [_p, _p, _i, _i, _i, _i, _p]),
    "tedm_final_conv1x1": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _p]),
    "tedm_nchw_f32_to_nhwc_bf16": (_i, [_p, _p, _i, _i, _i, _p]),
    "tedm_nhwc_bf16_to_nchw_f32": (_i, [_p, _p, _i, _i, _i, _p]),
    "tedm_head_infer": (_i, [C.POINTER(HeadArgs), _p]),
    "tedm_head_train_z1": (_i, [C.POINTER(HeadArgs), _p, _p, _p]),
    "tedm_bn_finalize": (_i, [_p, C.c_double, _p, _p, _f, _f, _p, _p, _p, _i, _p]),
    "tedm_head_fold_w2": (_i, [_p, _p, _p, _p, _p, _p, _p]),
    "tedm_head_z2_stats": (_i, [_p, _p, _i64, _p]),
    "tedm_head_train_tail": (_i, [_i, _p, _p, _p, _p, _p, _p, _p, _p, C.c_double, _i64, _p]),
    "tedm_head_bn1_bwd": (_i, [_i, _p, _p, _p, _p, _p, C.POINTER(_p), C.POINTER(_i), _i, _i, _i, _i, C.c_double, _p]),
    "tedm_head_param_grads": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "tedm_ensemble_mask": (_i, [_p, _p, _p, _i, _i, _i, _p]),
    "tedm_weight_to_dgrad": (_i, [_p, _p, _i, _i, _i, _p]),
    "tedm_wgrad_to_oihw": (_i, [_p, _p, _i, _i, _i, _p]),
    "tedm_gn_silu_bwd": (_i, [_p, _p, _p, _i, _p, _p, _p, _i, _i, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _f, _p]),
    "tedm_layernorm_bwd": (_i, [_p, _p, _p, _p, _p, _p, _i64, _i, _f, _p]),
    "tedm_bias_grad": (_i, [_p, _p, _i64, _i, _p]),
    "tedm_add_bf16": (_i, [_p, _p, _p, _i64, _p]),
    "tedm_final_conv1x1_bwd": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p]),
    "tedm_stem_conv7x7_wgrad": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _p]),
    "tedm_time_embed_train": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _p]),
    "tedm_linear_bwd": (_i, [_p, _p, _i, _p, _i, _p, _p, _p, _p, _i, _i, _i, _p]),
    "tedm_linear_attention_bwd_workspace": (_i64, [_i, _i, _i, _i]),
    "tedm_linear_attention_bwd": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _f, _p]),
    "tedm_attention_bwd": (_i, [_p, _p, _p, _i, _i, _i, _i, _f, _p]),
    "tedm_attention_bwd_flash_workspace": (_i64, [_i, _i, _i]),
    "tedm_attention_bwd_flash": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _f, _p]),
    "tedm_adam_step": (_i, [_p, _p, _p, _p, _i64, _f, _f, _f, _f, _f, _i, _p, _f, _p]),
    "tedm_bce_logits": (_i, [_p, _p, _p, _p, _p, _p, C.c_longlong, C.c_longlong, _i, _f, _p]),
    "tedm_bce_workspace_floats": (_i, [C.c_longlong]),
    "tedm_seg_metrics": (_i, [_p, _i, _p, _p, C.c_longlong, C.c_longlong, _i, _p]),
    "tedm_u8_to_unit": (_i, [_p, _p, C.c_longlong, _p]),
    "tedm_u8_masks_to_label": (_i, [_p, _p, C.c_longlong, C.c_longlong, _i, _p]),
    "tedm_debug_umma_probe": (_i, [_p, _p, C.POINTER(_i), C.POINTER(_i), _i, _p, _p]),
    "tedm_f32_split": (_i, [_p, _p, _p, _i64, _p]),
    "tedm_f32_stem_conv7x7": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _p]),
    "tedm_f32_gn_silu": (_i, [_p, _p, _i, _p, _p, _p, _i, _i, _p, _p, _i, _i, _i, _i, _f, _p]),
    "tedm_f32_layernorm": (_i, [_p, _p, _p, _p, _i64, _i, _f, _p]),
    "tedm_f32_linear_attention_workspace": (_i64, [_i, _i, _i]),
    "tedm_f32_linear_attention": (_i, [_p, _p, _p, _i, _i, _i, _i, _f, _p]),
    "tedm_f32_attention": (_i, [_p, _p, _p, _i, _i, _i, _i, _f, _p]),
    "tedm_f32_final_conv1x1": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _p]),
    "tedm_f32_add": (_i, [_p, _p, _p, _i64, _p]),
}

_lib: Optional[C.CDLL] = None
launches = 0  # kernels-launching C calls made so far (bench.py reads this for its gpu_launches claim)
conv_flops = 0  # executed conv MACs*2 so far (minimal form: the folded upsample conv counts 4 taps, not 9)
conv_timer = None  # bench.py instrumentation: callable(flops) -> context manager bracketing one conv launch
elem_timer = None  # bench.py instrumentation: callable(kernel name, algorithmic bytes) -> context manager (memory-bound kernels)


def load() -> C.CDLL:
    """dlopen the library and bind every declared symbol (fails loudly if one is missing)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} not found: build it with `python -m tedm_b200._build` "
                               "(tedm_b200 has no non-CUDA fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the .so does not export it
            fn.restype, fn.argtypes = res, args
        _lib = lib
        if os.environ.get("TEDM_DETERMINISTIC", "0") == "1":
            lib.tedm_conv_set_deterministic(1)
        if os.environ.get("TEDM_CTA_PAIRS", "1") != "1":        # A/B runs: 0 = one CTA per conv tile everywhere, 2 = pairs wherever possible
            lib.tedm_conv_set_cta_pairs(int(os.environ["TEDM_CTA_PAIRS"]))
        if os.environ.get("TEDM_HALO", "1") != "1":             # A/B runs: 0 = one activation box per tap in every 3x3 conv
            lib.tedm_conv_set_halo(int(os.environ["TEDM_HALO"]))
        if os.environ.get("TEDM_WS", "1") != "1":               # A/B runs: 0 = no weight-stationary tiles, 2 = single-row tiles only
            lib.tedm_conv_set_ws(int(os.environ["TEDM_WS"]))
    return _lib


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().tedm_last_error().decode(errors="replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


def _ptr(t: Optional[torch.Tensor], dtype=None, name="tensor") -> Optional[int]:
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (tedm_b200 has no CPU path)")
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")
    return t.data_ptr()


def _call(name: str, *args) -> None:
    global launches
    launches += 1
    _check(getattr(load(), name)(*args), name)


# ------------------------------------------------------------------------------------------------
# DDPM arithmetic
# ------------------------------------------------------------------------------------------------
def q_sample(x0: torch.Tensor, noise: torch.Tensor, t: torch.Tensor, sqrt_ac: torch.Tensor,
             sqrt_1m_ac: torch.Tensor, normalize: bool = False, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    b = x0.shape[0]
    chw = x0.numel() // b if b else int(torch.Size(x0.shape[1:]).numel())
    if noise.shape != x0.shape or t.shape != (b,):
        raise ValueError("q_sample: shape mismatch")
    out = torch.empty_like(x0) if out is None else out
    _call("tedm_q_sample", _ptr(x0, torch.float32, "x0"), _ptr(noise, torch.float32, "noise"),
          _ptr(t, torch.int64, "t"), _ptr(sqrt_ac, torch.float32), _ptr(sqrt_1m_ac, torch.float32),
          _ptr(out, torch.float32, "out"), b, chw, sqrt_ac.numel(), int(normalize), _stream())
    return out


def l1_loss(pred, target, t, p2_weight, want_grad: bool = False):
    b, chw = pred.shape[0], pred[0].numel()
    per_img = torch.empty(b, device=pred.device, dtype=torch.float32)
    loss = torch.empty(1, device=pred.device, dtype=torch.float32)
    grad = torch.empty_like(pred) if want_grad else None
    _call("tedm_l1_loss", _ptr(pred, torch.float32, "pred"), _ptr(target, torch.float32, "target"),
          _ptr(t, torch.int64, "t"), _ptr(p2_weight, torch.float32), _ptr(per_img), _ptr(loss), _ptr(grad), b, chw,
          p2_weight.numel(), _stream())
    return loss[0], per_img, grad


def sampler_step(x_t, eps, z, c_recip: float, c_recipm1: float, coef1: float, coef2: float, sigma: float,
                 k_lo: int, q_weight: float, want_x0: bool = False):
    b, chw = x_t.shape[0], x_t[0].numel()
    out = torch.empty_like(x_t)
    x0h = torch.empty_like(x_t) if want_x0 else None
    s = torch.empty(b, device=x_t.device, dtype=torch.float32)
    _call("tedm_sampler_step", _ptr(x_t, torch.float32, "x_t"), _ptr(eps, torch.float32, "eps"),
          _ptr(z, torch.float32, "z"), _ptr(out), _ptr(x0h), _ptr(s), c_recip, c_recipm1, coef1, coef2, sigma,
          k_lo, q_weight, b, chw, _stream())
    return out, x0h, s


def sampler_step_dev(x_t, eps, z, coefs, k_lo: int, q_weight: float, out=None):
    """sampler_step with the step's five schedule values in the device tensor `coefs` (graph-replayable)."""
    b, chw = x_t.shape[0], x_t[0].numel()
    if out is None:
        out = torch.empty_like(x_t)
    _call("tedm_sampler_step_dev", _ptr(x_t, torch.float32, "x_t"), _ptr(eps, torch.float32, "eps"), _ptr(z, torch.float32, "z"),
          _ptr(out), None, None, _ptr(coefs, torch.float32, "coefs"), k_lo, q_weight, b, chw, _stream())
    return out


# ------------------------------------------------------------------------------------------------
# UNet pieces (activations: NHWC bf16 tensors of shape (B, H, W, C))
# ------------------------------------------------------------------------------------------------
def time_embed(t, freq, w1, b1, w2, b2):
    b, dim, tdim = t.shape[0], w1.shape[1], w1.shape[0]
    out = torch.empty(b, tdim, device=t.device, dtype=torch.float32)
    _call("tedm_time_embed", _ptr(t, torch.int64, "t"), _ptr(freq, torch.float32), _ptr(w1, torch.float32),
          _ptr(b1, torch.float32), _ptr(w2, torch.float32), _ptr(b2, torch.float32), _ptr(out), b, dim, tdim, _stream())
    return out


def time_proj(temb, w_cat, b_cat):
    b, tdim, total = temb.shape[0], temb.shape[1], w_cat.shape[0]
    out = torch.empty(b, total, device=temb.device, dtype=torch.float32)
    _call("tedm_time_proj", _ptr(temb, torch.float32), _ptr(w_cat, torch.float32), _ptr(b_cat, torch.float32),
          _ptr(out), b, tdim, total, _stream())
    return out


def stem_conv7x7(x, weight, bias):
    b, cin, h, w = x.shape
    cout = weight.shape[0]
    out = torch.empty(b, h, w, cout, device=x.device, dtype=torch.bfloat16)
    _call("tedm_stem_conv7x7", _ptr(x, torch.float32, "x"), _ptr(weight, torch.float32), _ptr(bias, torch.float32),
          _ptr(out), b, cin, h, w, cout, _stream())
    return out


def weight_to_krsc(w: torch.Tensor) -> torch.Tensor:
    cout, cin, kh, kw = w.shape
    out = torch.empty(cout, kh, kw, cin, device=w.device, dtype=torch.bfloat16)
    _call("tedm_weight_to_krsc", _ptr(w.contiguous(), torch.float32, "weight"), _ptr(out), cout, cin, kh, kw, _stream())
    return out


def fold_upsample_weight(w: torch.Tensor) -> torch.Tensor:
    cout, cin, kh, kw = w.shape
    if (kh, kw) != (3, 3):
        raise ValueError("fold_upsample_weight expects a 3x3 kernel")
    out = torch.empty(4, cout, 2, 2, cin, device=w.device, dtype=torch.bfloat16)
    _call("tedm_fold_upsample_weight", _ptr(w.contiguous(), torch.float32, "weight"), _ptr(out), cout, cin, _stream())
    return out


MODE_1X1, MODE_3X3, MODE_4X4S2, MODE_UP3X3 = 0, 1, 2, 3


def set_deterministic(enable: bool = True) -> None:
    """Bit-reproducible convolution weight gradients (the generic weight-gradient kernel then adds its split-K slices in
    slice order instead of arrival order).  Also switched on by TEDM_DETERMINISTIC=1 in the environment."""
    load().tedm_conv_set_deterministic(int(bool(enable)))


def set_cta_pairs(mode: int = 1) -> None:
    """0: one CTA per conv tile; 1 (default): CTA pairs (tcgen05 cta_group::2) where they pay; 2: wherever the geometry allows."""
    load().tedm_conv_set_cta_pairs(int(mode))


def conv_gn_parts(oh: int, ow: int) -> int:
    return load().tedm_conv_gn_parts(oh, ow)


def _nhwc(t: Optional[torch.Tensor], name: str, dtype=torch.bfloat16):
    """(pointer, image stride in elements) of an NHWC tensor whose batch axis may be strided."""
    if t is None:
        return None, 0
    if not t.is_cuda or t.dtype != dtype or t.dim() != 4:
        raise TypeError(f"{name} must be a 4-D CUDA {dtype} tensor (B, H, W, C)")
    b, h, w, c = t.shape
    if t.stride(3) != 1 or t.stride(2) != c or t.stride(1) != w * c:
        raise ValueError(f"{name} must be NHWC-contiguous within each image")
    return t.data_ptr(), t.stride(0)


def conv_igemm(src0: torch.Tensor, weight: torch.Tensor, mode: int, cout: int, bias=None, src1=None, residual=None,
               gn_groups: int = 0, out: Optional[torch.Tensor] = None, out_dtype=torch.bfloat16, split: int = 0,
               residual2=None, extra: Sequence = (), residual_affine=None, src0_affine=None):
    """Returns out (B, Ho, Wo, cout) bf16 (or fp32) [, gn_partial (B, parts, groups, 2) fp32 if gn_groups > 0].
    src0/src1/out may be batch-strided views (e.g. x[s::S]); residual must share out's strides.
    split > 0: returns (out[..., :split], out2[..., split:]) as two dense tensors (+ residual / residual2).
    extra: up to four more A sources [(tensor, centre_only)], walked after src0/src1 inside every tap (centre_only: a
    1x1 branch folded into the centre tap of a 3x3).
    residual_affine: (B, cout, 2) fp32 from gn_affine(): the residual enters as SiLU(GroupNorm(residual)).
    src0_affine: (B, c0, 2) fp32 from gn_affine(): the conv reads SiLU(GroupNorm(src0)) (conv_src_affine_supported())."""
    b, h, w, c0 = src0.shape
    c1 = src1.shape[3] if src1 is not None else 0
    if src1 is not None and src1.shape[:3] != src0.shape[:3]:
        raise ValueError("conv_igemm: src0/src1 extent mismatch")
    oh, ow = (h // 2, w // 2) if mode == MODE_4X4S2 else ((2 * h, 2 * w) if mode == MODE_UP3X3 else (h, w))
    taps = {MODE_1X1: 1, MODE_3X3: 9, MODE_4X4S2: 16, MODE_UP3X3: 16}[mode]
    if len(extra) > 4:
        raise ValueError("conv_igemm: at most four extra sources")
    k_total = taps * (c0 + c1) + sum((1 if ctr else taps) * t.shape[3] for t, ctr in extra)
    if weight.numel() != cout * k_total:
        raise ValueError(f"conv_igemm: weight has {weight.numel()} elements, expected {cout * k_total}")
    out2 = None
    if split:
        if out is not None or gn_groups or out_dtype != torch.bfloat16:
            raise ValueError("conv_igemm: split output is bf16, freshly allocated, without GroupNorm statistics")
        out = torch.empty(b, oh, ow, split, device=src0.device, dtype=torch.bfloat16)
        out2 = torch.empty(b, oh, ow, cout - split, device=src0.device, dtype=torch.bfloat16)
        if residual2 is not None and residual2.shape != out2.shape:
            raise ValueError("conv_igemm: residual2 must have out2's shape")
    elif out is None:
        out = torch.empty(b, oh, ow, cout, device=src0.device, dtype=out_dtype)
    elif tuple(out.shape) != (b, oh, ow, cout):
        raise ValueError(f"conv_igemm: out has shape {tuple(out.shape)}, expected {(b, oh, ow, cout)}")
    gnp = None
    if gn_groups:
        gnp = torch.empty(b, conv_gn_parts(oh, ow), gn_groups, 2, device=src0.device, dtype=torch.float32)
    p0, s0 = _nhwc(src0, "src0")
    p1, s1 = _nhwc(src1, "src1")
    po, so = _nhwc(out, "out", out.dtype)
    if out.dtype not in (torch.bfloat16, torch.float32):
        raise TypeError("conv_igemm: out must be bf16 or fp32")
    pr, sr = _nhwc(residual, "residual")
    if residual is not None and (sr != so or residual.shape != out.shape):
        raise ValueError("conv_igemm: residual must have out's shape and strides")
    a = ConvArgs(p0, p1, _ptr(weight, torch.bfloat16, "weight"), _ptr(bias, torch.float32, "bias"), pr, po, _ptr(gnp),
                 b, h, w, c0, c1, cout, mode, gn_groups, 1 if out.dtype == torch.float32 else 0, s0, s1,
                 0 if split else so, split, _ptr(out2), _ptr(residual2, torch.bfloat16, "residual2"))
    if residual_affine is not None:
        if residual is None or tuple(residual_affine.shape) != (b, cout, 2):
            raise ValueError("conv_igemm: residual_affine is (B, cout, 2) and needs a residual")
        a.residual_affine = _ptr(residual_affine, torch.float32, "residual_affine")
    if src0_affine is not None:
        if tuple(src0_affine.shape) != (b, c0, 2):
            raise ValueError("conv_igemm: src0_affine is (B, c0, 2)")
        a.src0_affine = _ptr(src0_affine, torch.float32, "src0_affine")
    a.n_extra = len(extra)
    for i, (t, ctr) in enumerate(extra):
        if t.shape[:3] != src0.shape[:3]:
            raise ValueError("conv_igemm: extra source extent mismatch")
        a.extra_src[i], a.extra_image_stride[i] = _nhwc(t, f"extra[{i}]")
        a.extra_c[i], a.extra_center[i] = t.shape[3], int(bool(ctr))
    global conv_flops
    flops = 2 * b * (h * w if mode == MODE_UP3X3 else oh * ow) * cout * k_total
    conv_flops += flops
    if conv_timer is not None:
        with conv_timer(flops, (mode, b, h, w, c0, c1, cout, bool(gn_groups), residual is not None)):
            _call("tedm_conv_igemm_fwd", C.byref(a), _stream())
    else:
        _call("tedm_conv_igemm_fwd", C.byref(a), _stream())
    if split:
        return out, out2
    return (out, gnp) if gn_groups else out


_wgrad_ws = {}


def _wgrad_workspace(device) -> torch.Tensor:
    """Per-device scratch for the split-K partial tiles of the weight-gradient kernels (stream-ordered reuse: every
    weight gradient of a backward pass is issued on ONE stream)."""
    key = (device.type, device.index)
    if key not in _wgrad_ws:
        # zeros: the tail of the workspace holds the generic kernel's split-K turn counters (zero at first use, self-resetting)
        _wgrad_ws[key] = torch.zeros(load().tedm_conv_igemm_wgrad_workspace(), device=device, dtype=torch.float32)
    return _wgrad_ws[key]


def conv_wgrad(src0: torch.Tensor, dy: torch.Tensor, mode: int, src1=None, grad_oihw: Optional[torch.Tensor] = None):
    """Weight gradient of conv_igemm(src0[, src1]) given the NHWC bf16 output gradient: returns fp32 (cout, taps, c0+c1),
    or, when grad_oihw (the fp32 OIHW parameter gradient) is given, accumulates into it and returns it."""
    b, h, w, c0 = src0.shape
    c1 = src1.shape[3] if src1 is not None else 0
    cout = dy.shape[3]
    taps = {MODE_1X1: 1, MODE_3X3: 9, MODE_4X4S2: 16, MODE_UP3X3: 16}[mode]
    oh, ow = (h // 2, w // 2) if mode == MODE_4X4S2 else ((2 * h, 2 * w) if mode == MODE_UP3X3 else (h, w))
    if tuple(dy.shape) != (b, oh, ow, cout):
        raise ValueError(f"conv_wgrad: dy has shape {tuple(dy.shape)}, expected {(b, oh, ow, cout)}")
    if grad_oihw is None:
        dw = torch.empty(cout, taps, c0 + c1, device=src0.device, dtype=torch.float32)
    else:
        dw = grad_oihw
        khw = 9 if mode == MODE_UP3X3 else taps
        if dw.numel() != cout * (c0 + c1) * khw:
            raise ValueError("conv_wgrad: grad_oihw has the wrong size")
    p0, s0 = _nhwc(src0, "src0")
    p1, s1 = _nhwc(src1, "src1")
    pd, sd = _nhwc(dy, "dy")
    a = ConvArgs(p0, p1, None, None, None, None, None, b, h, w, c0, c1, cout, mode, 0, 0, s0, s1, sd)
    ws = _wgrad_workspace(src0.device)       # split-K partials / turn counters: slices are summed in a fixed order (deterministic)
    if conv_timer is not None:
        flops = 2 * b * (h * w if mode == MODE_UP3X3 else oh * ow) * cout * taps * (c0 + c1)
        with conv_timer(flops, ("wgrad", mode, b, h, w, c0, c1, cout)):
            _call("tedm_conv_igemm_wgrad", C.byref(a), pd, _ptr(dw, torch.float32, "dw"), 0 if grad_oihw is None else 1,
                  _ptr(ws), _stream())
        return dw
    _call("tedm_conv_igemm_wgrad", C.byref(a), pd, _ptr(dw, torch.float32, "dw"), 0 if grad_oihw is None else 1, _ptr(ws),
          _stream())
    return dw


def gn_silu(x, gn_partial, gamma, beta, groups: int, eps: float = 1e-5, scale_shift=None, ss_offset: int = 0,
            residual=None):
    b, h, w, c = x.shape
    out = torch.empty_like(x)
    args = ("tedm_gn_silu_fwd", _ptr(x, torch.bfloat16, "x"), _ptr(gn_partial, torch.float32), gn_partial.shape[1],
            _ptr(gamma, torch.float32), _ptr(beta, torch.float32), _ptr(scale_shift, torch.float32),
            scale_shift.shape[1] if scale_shift is not None else 0, ss_offset, _ptr(residual, torch.bfloat16, "residual"),
            _ptr(out), b, h * w, c, groups, eps, _stream())
    if elem_timer is not None:
        with elem_timer("gn_silu_kernel", x.numel() * 2 * (3 if residual is not None else 2)):
            _call(*args)
    else:
        _call(*args)
    return out


def conv_src_affine_supported(h: int, w: int, c0: int, cout: int) -> bool:
    return bool(load().tedm_conv_src_affine_supported(h, w, c0, cout))


def gn_affine(gn_partial, gamma, beta, groups: int, hw: int, eps: float = 1e-5, scale_shift=None, ss_offset: int = 0):
    """(B, C, 2) fp32: the halved per-(image, channel) affine (a / 2, b / 2) of gn_silu, for conv_igemm(residual_affine=...)."""
    b, c = gn_partial.shape[0], gamma.numel()
    out = torch.empty(b, c, 2, device=gn_partial.device, dtype=torch.float32)
    _call("tedm_gn_affine", _ptr(gn_partial, torch.float32), gn_partial.shape[1], _ptr(gamma, torch.float32),
          _ptr(beta, torch.float32), _ptr(scale_shift, torch.float32), scale_shift.shape[1] if scale_shift is not None else 0,
          ss_offset, _ptr(out), b, hw, c, groups, eps, _stream())
    return out


def layernorm(x, g, eps: float = 1e-5, residual=None):
    c = x.shape[-1]
    out = torch.empty_like(x)
    _call("tedm_layernorm_fwd", _ptr(x, torch.bfloat16, "x"), _ptr(g, torch.float32), _ptr(residual, torch.bfloat16),
          _ptr(out), x.numel() // c, c, eps, _stream())
    return out


def linear_attention(qkv, heads: int = 4, dim_head: int = 32, scale: Optional[float] = None, want_workspace: bool = False):
    b, h, w, c3 = qkv.shape
    n = h * w
    ws_n = load().tedm_linear_attention_workspace(b, n, heads, dim_head)
    if ws_n < 0:
        raise RuntimeError("linear_attention: unsupported configuration")
    ws = torch.empty(ws_n, device=qkv.device, dtype=torch.float32)
    out = torch.empty(b, h, w, heads * dim_head, device=qkv.device, dtype=torch.bfloat16)
    _call("tedm_linear_attention_fwd", _ptr(qkv, torch.bfloat16, "qkv"), _ptr(out), _ptr(ws), b, n, heads, dim_head,
          float(dim_head ** -0.5 if scale is None else scale), _stream())
    return (out, ws) if want_workspace else out


def linear_attention_fused_supported(n: int, channels: int, heads: int = 4, dim_head: int = 32) -> bool:
    return bool(load().tedm_linear_attention_fused_supported(n, channels, heads, dim_head))


def linear_attention_block_fused(x, wqkv, g_pre, wout, b_out, g_out, heads: int = 4, dim_head: int = 32,
                                 scale: Optional[float] = None, eps: float = 1e-5):
    """Residual(PreNorm(LinearAttention)) inference forward in 3 launches; x NHWC bf16 -> same shape."""
    if x.dim() != 4 or not x.is_contiguous():
        raise ValueError("x must be a contiguous NHWC tensor")
    b, h, w, c = x.shape
    scale = dim_head ** -0.5 if scale is None else scale
    out = torch.empty_like(x)
    ws = torch.empty(load().tedm_linear_attention_fused_workspace(b, h * w), device=x.device, dtype=torch.float32)
    _call("tedm_linear_attention_fused_fwd", _ptr(x, torch.bfloat16, "x"), _ptr(wqkv, torch.bfloat16, "wqkv"),
          _ptr(g_pre, torch.float32, "g_pre"), _ptr(wout, torch.bfloat16, "wout"), _ptr(b_out, torch.float32, "b_out"),
          _ptr(g_out, torch.float32, "g_out"), _ptr(out), _ptr(ws), b, h * w, c, heads, dim_head, float(scale), float(eps),
          _stream())
    return out


LINATTN_TC_MAX_SHIFT = 40.0      # exp(-2 * 40) is still a normal fp32 / bf16 number


def linear_attention_tc_supported(n: int, channels: int, heads: int = 4, dim_head: int = 32) -> bool:
    return bool(load().tedm_linear_attention_tc_supported(n, channels, heads, dim_head))


def linear_attention_tc_weights(wqkv: torch.Tensor, g_pre: torch.Tensor, heads: int = 4, dim_head: int = 32):
    """Operands of the tcgen05 LinearAttention block derived from the to_qkv weight (fp32 or bf16, [384][C] or OIHW) and the
    pre-norm gain: (wqkv_g bf16 [384][C] = W * g, log2(e) * bound, bound) where bound >= |q|, |k| for every pixel:
    q = w . (LN(x) * g) with ||LN(x)||_2 <= sqrt(C), hence |q_r| <= ||w_r * g||_2 sqrt(C) (+ 2 % for the bf16 rounding)."""
    hid = heads * dim_head
    c = wqkv.numel() // (3 * hid)
    wf = wqkv.reshape(3 * hid, c).float() * g_pre.reshape(1, c).float()
    bound = float((wf[:2 * hid].to(torch.bfloat16).float().norm(dim=1).max() * (c ** 0.5) * 1.02 + 1e-3).item())
    wf[:hid] *= 1.4426950408889634                      # the q rows carry log2(e): the kernel's softmax over d is exp2(q')
    return wf.to(torch.bfloat16).contiguous(), bound * 1.4426950408889634, bound


def linear_attention_block_tc(x, wqkv_g, shift_log2: float, wout, b_out, g_out, heads: int = 4, dim_head: int = 32,
                              scale: Optional[float] = None, eps: float = 1e-5, want_workspace: bool = False):
    """Residual(PreNorm(LinearAttention)) inference forward on tcgen05 (csrc/attention_tc.cu); x NHWC bf16 -> same shape.
    wqkv_g / shift_log2 come from linear_attention_tc_weights."""
    if x.dim() != 4 or not x.is_contiguous():
        raise ValueError("x must be a contiguous NHWC tensor")
    b, h, w, c = x.shape
    scale = dim_head ** -0.5 if scale is None else scale
    out = torch.empty_like(x)
    ws = torch.empty(load().tedm_linear_attention_tc_workspace(b, h * w, c), device=x.device, dtype=torch.float32)
    _call("tedm_linear_attention_tc_fwd", _ptr(x, torch.bfloat16, "x"), _ptr(wqkv_g, torch.bfloat16, "wqkv_g"), float(shift_log2),
          _ptr(wout, torch.bfloat16, "wout"), _ptr(b_out, torch.float32, "b_out"), _ptr(g_out, torch.float32, "g_out"), _ptr(out),
          _ptr(ws), b, h * w, c, heads, dim_head, float(scale), float(eps), _stream())
    return (out, ws) if want_workspace else out


def attention(qkv, heads: int = 4, dim_head: int = 32, scale: float = 16.0):
    b, h, w, c3 = qkv.shape
    out = torch.empty(b, h, w, heads * dim_head, device=qkv.device, dtype=torch.bfloat16)
    _call("tedm_attention_fwd", _ptr(qkv, torch.bfloat16, "qkv"), _ptr(out), b, h * w, heads, dim_head, float(scale),
          _stream())
    return out


def upsample2x(x):
    b, h, w, c = x.shape
    out = torch.empty(b, 2 * h, 2 * w, c, device=x.device, dtype=torch.bfloat16)
    _call("tedm_upsample2x", _ptr(x, torch.bfloat16, "x"), _ptr(out), b, h, w, c, _stream())
    return out


def final_conv1x1(x, weight, bias):
    b, h, w, c = x.shape
    od = weight.shape[0]
    out = torch.empty(b, od, h, w, device=x.device, dtype=torch.float32)
    _call("tedm_final_conv1x1", _ptr(x, torch.bfloat16, "x"), _ptr(weight, torch.float32), _ptr(bias, torch.float32),
          _ptr(out), b, h * w, c, od, _stream())
    return out


def nchw_to_nhwc_bf16(x):
    b, c, h, w = x.shape
    out = torch.empty(b, h, w, c, device=x.device, dtype=torch.bfloat16)
    _call("tedm_nchw_f32_to_nhwc_bf16", _ptr(x, torch.float32, "x"), _ptr(out), b, c, h * w, _stream())
    return out


def nhwc_to_nchw_f32(x):
    b, h, w, c = x.shape
    out = torch.empty(b, c, h, w, device=x.device, dtype=torch.float32)
    _call("tedm_nhwc_bf16_to_nchw_f32", _ptr(x, torch.bfloat16, "x"), _ptr(out), b, c, h * w, _stream())
    return out


# ------------------------------------------------------------------------------------------------
# training step: backward pieces (parameter gradients are ACCUMULATED into the fp32 tensors passed in)
# ------------------------------------------------------------------------------------------------
_TAPS = {MODE_1X1: 1, MODE_3X3: 9, MODE_4X4S2: 16, MODE_UP3X3: 16}


def weight_to_dgrad(w: torch.Tensor, mode: int) -> torch.Tensor:
    """bf16 operand of the data-gradient conv for a forward conv of `mode` with OIHW weight w."""
    cout, cin = w.shape[0], w.shape[1]
    out = torch.empty(cout * cin * _TAPS[mode], device=w.device, dtype=torch.bfloat16)
    _call("tedm_weight_to_dgrad", _ptr(w.contiguous(), torch.float32, "weight"), _ptr(out), cout, cin, mode, _stream())
    return out


def prepare_weights(table_dev: torch.Tensor, n_entries: int, total_ctas: int) -> None:
    _call("tedm_prepare_weights", _ptr(table_dev, torch.uint8, "table"), n_entries, total_ctas, _stream())


def wgrad_to_oihw(dw: torch.Tensor, grad: torch.Tensor, mode: int) -> None:
    cout, cin = grad.shape[0], grad.shape[1]
    _call("tedm_wgrad_to_oihw", _ptr(dw, torch.float32, "dw"), _ptr(grad, torch.float32, "grad"), cout, cin, mode, _stream())


def gn_silu_bwd(x, dy, gn_partial, gamma, beta, groups: int, dgamma, dbeta, dbias=None, eps: float = 1e-5, scale_shift=None,
                ss_offset: int = 0, dscale_shift=None):
    b, h, w, c = x.shape
    dx = torch.empty_like(x)
    ws = torch.empty(3 * b * c, device=x.device, dtype=torch.float32)
    _call("tedm_gn_silu_bwd", _ptr(x, torch.bfloat16, "x"), _ptr(dy, torch.bfloat16, "dy"), _ptr(gn_partial, torch.float32),
          gn_partial.shape[1], _ptr(gamma, torch.float32), _ptr(beta, torch.float32), _ptr(scale_shift, torch.float32),
          scale_shift.shape[1] if scale_shift is not None else 0, ss_offset, _ptr(dx), _ptr(ws), _ptr(dgamma, torch.float32),
          _ptr(dbeta, torch.float32), _ptr(dbias, torch.float32), _ptr(dscale_shift, torch.float32), b, h * w, c, groups, eps,
          _stream())
    return dx


def layernorm_bwd(x, g, dy, dg, eps: float = 1e-5, add=None):
    c = x.shape[-1]
    dx = torch.empty_like(x)
    _call("tedm_layernorm_bwd", _ptr(x, torch.bfloat16, "x"), _ptr(g, torch.float32), _ptr(dy, torch.bfloat16, "dy"),
          _ptr(add, torch.bfloat16, "add"), _ptr(dx), _ptr(dg, torch.float32, "dg"), x.numel() // c, c, eps, _stream())
    return dx


def bias_grad(dy, dbias) -> None:
    c = dy.shape[-1]
    _call("tedm_bias_grad", _ptr(dy, torch.bfloat16, "dy"), _ptr(dbias, torch.float32, "dbias"), dy.numel() // c, c, _stream())


def add_bf16(a, b, out=None):
    out = torch.empty_like(a) if out is None else out
    _call("tedm_add_bf16", _ptr(a, torch.bfloat16, "a"), _ptr(b, torch.bfloat16, "b"), _ptr(out, torch.bfloat16), a.numel(),
          _stream())
    return out


def final_conv1x1_bwd(h, weight, dout, dweight, dbias):
    b, hh, ww, c = h.shape
    od = weight.shape[0]
    dh = torch.empty_like(h)
    _call("tedm_final_conv1x1_bwd", _ptr(h, torch.bfloat16, "h"), _ptr(weight, torch.float32), _ptr(dout, torch.float32, "dout"),
          _ptr(dh), _ptr(dweight, torch.float32), _ptr(dbias, torch.float32), b, hh * ww, c, od, _stream())
    return dh


def stem_conv7x7_wgrad(x, dy, dweight, dbias) -> None:
    b, cin, h, w = x.shape
    _call("tedm_stem_conv7x7_wgrad", _ptr(x, torch.float32, "x"), _ptr(dy, torch.bfloat16, "dy"), _ptr(dweight, torch.float32),
          _ptr(dbias, torch.float32), b, cin, h, w, dy.shape[-1], _stream())


def time_embed_train(t, freq, w1, b1, w2, b2):
    b, dim, tdim = t.shape[0], w1.shape[1], w1.shape[0]
    emb = torch.empty(b, dim, device=t.device, dtype=torch.float32)
    hid = torch.empty(b, tdim, device=t.device, dtype=torch.float32)
    out = torch.empty(b, tdim, device=t.device, dtype=torch.float32)
    _call("tedm_time_embed_train", _ptr(t, torch.int64, "t"), _ptr(freq, torch.float32), _ptr(w1, torch.float32),
          _ptr(b1, torch.float32), _ptr(w2, torch.float32), _ptr(b2, torch.float32), _ptr(emb), _ptr(hid), _ptr(out), b, dim,
          tdim, _stream())
    return emb, hid, out


ACT_NONE, ACT_SILU, ACT_GELU = 0, 1, 2


def linear_bwd(dy_raw, y_pre, act_y: int, x, act_x: int, w, dw, db, want_dx: bool = True):
    b, n_out = dy_raw.shape
    n_in = x.shape[1]
    dx = torch.empty(b, n_in, device=x.device, dtype=torch.float32) if want_dx else None
    _call("tedm_linear_bwd", _ptr(dy_raw, torch.float32, "dy_raw"), _ptr(y_pre, torch.float32), act_y, _ptr(x, torch.float32, "x"),
          act_x, _ptr(w, torch.float32, "w"), _ptr(dw, torch.float32, "dw"), _ptr(db, torch.float32), _ptr(dx), b, n_out, n_in,
          _stream())
    return dx


def linear_attention_bwd(qkv, dout, fwd_ws, heads: int = 4, dim_head: int = 32, scale: Optional[float] = None):
    b, h, w, c3 = qkv.shape
    n = h * w
    ws_n = load().tedm_linear_attention_bwd_workspace(b, n, heads, dim_head)
    if ws_n < 0:
        raise RuntimeError("linear_attention_bwd: unsupported configuration")
    ws = torch.empty(ws_n, device=qkv.device, dtype=torch.float32)
    dqkv = torch.empty_like(qkv)
    _call("tedm_linear_attention_bwd", _ptr(qkv, torch.bfloat16, "qkv"), _ptr(dout, torch.bfloat16, "dout"),
          _ptr(fwd_ws, torch.float32, "fwd_ws"), _ptr(dqkv), _ptr(ws), b, n, heads, dim_head,
          float(dim_head ** -0.5 if scale is None else scale), _stream())
    return dqkv


def attention_bwd(qkv, dout, heads: int = 4, dim_head: int = 32, scale: float = 16.0, o=None):
    """Mid-attention backward.  With the forward output `o` and n >= 64 tokens: the tensor-core flash form (any n);
    otherwise the one-CTA-per-(image, head) kernel (n <= 256)."""
    b, h, w, c3 = qkv.shape
    dqkv = torch.empty_like(qkv)
    if o is not None and h * w >= 64:
        ws = torch.empty(load().tedm_attention_bwd_flash_workspace(b, h * w, heads), device=qkv.device, dtype=torch.float32)
        _call("tedm_attention_bwd_flash", _ptr(qkv, torch.bfloat16, "qkv"), _ptr(o, torch.bfloat16, "o"),
              _ptr(dout, torch.bfloat16, "dout"), _ptr(dqkv), _ptr(ws), b, h * w, heads, dim_head, float(scale), _stream())
        return dqkv
    _call("tedm_attention_bwd", _ptr(qkv, torch.bfloat16, "qkv"), _ptr(dout, torch.bfloat16, "dout"), _ptr(dqkv), b, h * w, heads,
          dim_head, float(scale), _stream())
    return dqkv


def adam_step(param, grad, exp_avg, exp_avg_sq, lr: float, beta1: float, beta2: float, eps: float, weight_decay: float,
              step: int, grad_scale: float = 1.0, step_counter: Optional[torch.Tensor] = None) -> None:
    n = param.numel()
    _call("tedm_adam_step", _ptr(param, torch.float32, "param"), _ptr(grad, torch.float32, "grad"),
          _ptr(exp_avg, torch.float32), _ptr(exp_avg_sq, torch.float32), n, lr, beta1, beta2, eps, weight_decay, step,
          _ptr(step_counter, torch.int32, "step_counter"), grad_scale, _stream())


# ------------------------------------------------------------------------------------------------
# head
# ------------------------------------------------------------------------------------------------
def head_infer(g_maps: Sequence[torch.Tensor], shifts: Sequence[int], n_sum: int, n_img: int, height: int, width: int,
               b1, bn1_a, bn1_c, w2, b2, bn2_a, bn2_c, w3, b3: float, f_full=None, w1_full=None, exact: bool = False):
    """g_maps: layer-1 outputs per level (fp32 / bf16 NHWC); optionally the full-resolution level as its bf16 feature map
    `f_full` + weight slice `w1_full` (its layer 1 is then fused into the tail kernel)."""
    a = HeadArgs()
    gdt = g_maps[0].dtype if g_maps else torch.float32
    a.f_full, a.w1_full = _ptr(f_full, torch.bfloat16, "f_full"), _ptr(w1_full, torch.bfloat16, "w1_full")
    a.c_full = f_full.shape[-1] if f_full is not None else 0
    a.exact = int(exact)        # fp32 precision mode: the plain-fp32 tail kernel (fp32 g maps) instead of the tensor-core one
    if gdt not in (torch.bfloat16, torch.float32):
        raise TypeError("head_infer: g maps must be bf16 or fp32")
    a.g_dtype = 1 if gdt == torch.float32 else 0
    for l, (g, s) in enumerate(zip(g_maps, shifts)):
        a.g[l] = _ptr(g, gdt, f"g[{l}]")
        a.shift[l] = s
    a.n_levels, a.n_sum, a.n_img, a.height, a.width = len(g_maps), n_sum, n_img, height, width
    a.c1, a.c2 = b1.numel(), b2.numel()
    a.b1, a.bn1_a, a.bn1_c = _ptr(b1, torch.float32), _ptr(bn1_a, torch.float32), _ptr(bn1_c, torch.float32)
    a.w2, a.b2, a.bn2_a, a.bn2_c = (_ptr(w2, torch.float32), _ptr(b2, torch.float32), _ptr(bn2_a, torch.float32),
                                    _ptr(bn2_c, torch.float32))
    a.w3, a.b3 = _ptr(w3, torch.float32), float(b3)
    logits = torch.empty(n_img, 1, height, width, device=b1.device, dtype=torch.float32)
    a.logits = _ptr(logits)
    _call("tedm_head_infer", C.byref(a), _stream())
    return logits


# ---- head training passes (see include/tedm_b200.h "head training") ----
def head_train_z1(g_maps: Sequence[torch.Tensor], shifts: Sequence[int], n_sum: int, n_img: int, height: int, width: int, b1):
    a = HeadArgs()
    a.g_dtype = 1
    for l, (g, s) in enumerate(zip(g_maps, shifts)):
        a.g[l] = _ptr(g, torch.float32, f"g[{l}]")
        a.shift[l] = s
    a.n_levels, a.n_sum, a.n_img, a.height, a.width, a.c1, a.c2 = len(g_maps), n_sum, n_img, height, width, b1.numel(), 32
    a.b1 = _ptr(b1, torch.float32, "b1")
    a1 = torch.empty(n_img, height, width, b1.numel(), device=b1.device, dtype=torch.bfloat16)
    sums = torch.zeros(2, b1.numel(), device=b1.device, dtype=torch.float32)
    _call("tedm_head_train_z1", C.byref(a), _ptr(a1), _ptr(sums), _stream())
    return a1, sums


def bn_finalize(sums, count: int, gamma, beta, eps: float, momentum: float, running_mean=None, running_var=None):
    c = gamma.numel()
    stats = torch.empty(4, c, device=gamma.device, dtype=torch.float32)
    _call("tedm_bn_finalize", _ptr(sums, torch.float32), float(count), _ptr(gamma, torch.float32), _ptr(beta, torch.float32),
          eps, momentum, _ptr(running_mean, torch.float32), _ptr(running_var, torch.float32), _ptr(stats), c, _stream())
    return stats


def head_fold_w2(w2, b2, stats1):
    dev = w2.device
    w2f = torch.empty(64, 128, device=dev, dtype=torch.bfloat16)
    b2f = torch.empty(64, device=dev, dtype=torch.float32)
    w2t = torch.empty(128, 64, device=dev, dtype=torch.bfloat16)
    _call("tedm_head_fold_w2", _ptr(w2, torch.float32, "w2"), _ptr(b2, torch.float32), _ptr(stats1, torch.float32), _ptr(w2f),
          _ptr(b2f), _ptr(w2t), _stream())
    return w2f, b2f, w2t


def head_z2_stats(z2):
    sums = torch.zeros(2, 32, device=z2.device, dtype=torch.float32)
    _call("tedm_head_z2_stats", _ptr(z2, torch.float32, "z2"), _ptr(sums), z2.numel() // 64, _stream())
    return sums


def head_train_tail(mode: int, z2, stats2, w3, b3=None, dlogit=None, S=None, count: int = 1):
    npix = z2.numel() // 64
    logits = dz2 = None
    if mode == 0:
        logits = torch.empty(z2.shape[0], 1, z2.shape[1], z2.shape[2], device=z2.device, dtype=torch.float32)
    if mode == 2:
        dz2 = torch.empty(z2.shape, device=z2.device, dtype=torch.bfloat16)
    _call("tedm_head_train_tail", mode, _ptr(z2, torch.float32, "z2"), _ptr(stats2, torch.float32), _ptr(w3, torch.float32),
          _ptr(b3, torch.float32), _ptr(dlogit, torch.float32, "dlogit"), _ptr(logits), _ptr(S, torch.float32), _ptr(dz2),
          float(count), npix, _stream())
    return logits if mode == 0 else dz2


def head_bn1_bwd(mode: int, dh1, a1, stats1, T, db1=None, shifts: Sequence[int] = (), count: int = 1):
    n, h, w, c = a1.shape
    pooled = []
    n_levels = len(shifts) if mode == 1 else 0
    ptrs = (_p * max(1, n_levels))()
    shs = (_i * max(1, n_levels))()
    for l in range(n_levels):
        d = torch.empty(n, h >> shifts[l], w >> shifts[l], c, device=a1.device, dtype=torch.bfloat16)
        pooled.append(d)
        ptrs[l] = d.data_ptr()
        shs[l] = shifts[l]
    _call("tedm_head_bn1_bwd", mode, _ptr(dh1, torch.float32, "dh1"), _ptr(a1, torch.bfloat16, "a1"), _ptr(stats1, torch.float32),
          _ptr(T, torch.float32), _ptr(db1, torch.float32), ptrs, shs, n_levels, n, h, w, float(count), _stream())
    return pooled


def head_param_grads(dw2f, stats1, S, T, dw2, db2, dg1, dbt1, dg2, dbt2, dw3, db3) -> None:
    _call("tedm_head_param_grads", _ptr(dw2f, torch.float32), _ptr(stats1, torch.float32), _ptr(S, torch.float32),
          _ptr(T, torch.float32), _ptr(dw2, torch.float32), _ptr(db2, torch.float32), _ptr(dg1, torch.float32),
          _ptr(dbt1, torch.float32), _ptr(dg2, torch.float32), _ptr(dbt2, torch.float32), _ptr(dw3, torch.float32),
          _ptr(db3, torch.float32), _stream())


def ensemble_mask(logits: torch.Tensor, n_steps: int):
    bs, c, h, w = logits.shape
    b = bs // n_steps
    prob = torch.empty(b, c, h, w, device=logits.device, dtype=torch.float32)
    mask = torch.empty(b, c, h, w, device=logits.device, dtype=torch.uint8)
    _call("tedm_ensemble_mask", _ptr(logits, torch.float32, "logits"), _ptr(prob), _ptr(mask), b, n_steps, c * h * w,
          _stream())
    return mask.bool(), prob


# ------------------------------------------------------------------------------------------------
# supervised segmentation: loss, metrics, input transport
# ------------------------------------------------------------------------------------------------
def _rows(logits: torch.Tensor, target: torch.Tensor):
    """(n_rows, row_len, target_repeat) of an NCHW logit tensor against its (possibly un-repeated) label tensor."""
    if logits.dim() != 4 or target.dim() != 4 or logits.shape[1:] != target.shape[1:]:
        raise ValueError(f"logits {tuple(logits.shape)} and target {tuple(target.shape)} must be NCHW with equal C, H, W")
    if target.shape[0] == 0 or logits.shape[0] % target.shape[0]:
        raise ValueError(f"logit batch {logits.shape[0]} is not a multiple of the label batch {target.shape[0]}")
    rep = logits.shape[0] // target.shape[0]
    if rep > 1 and logits.shape[1] != 1:
        raise ValueError("label repetition over timesteps needs single-channel logits")
    return logits.shape[0] * logits.shape[1], logits.shape[2] * logits.shape[3], rep


def bce_logits(logits: torch.Tensor, target: torch.Tensor, want_grad: bool = False, grad_scale: float = 1.0):
    """-> (loss scalar, per-(b,c) mean (B, C), grad or None).  target may hold B / S images: row r uses label r // S."""
    n_rows, row_len, rep = _rows(logits, target)
    dev = logits.device
    row_mean = torch.empty(logits.shape[0], logits.shape[1], device=dev, dtype=torch.float32)
    loss = torch.empty((), device=dev, dtype=torch.float32)
    grad = torch.empty_like(logits) if want_grad else None
    ws = torch.empty(load().tedm_bce_workspace_floats(n_rows), device=dev, dtype=torch.float32)
    _call("tedm_bce_logits", _ptr(logits, torch.float32, "logits"), _ptr(target, torch.float32, "target"), _ptr(row_mean),
          _ptr(loss), _ptr(grad), _ptr(ws), n_rows, row_len, rep, float(grad_scale), _stream())
    return loss, row_mean, grad


def seg_metrics(pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """pred: bool / uint8 mask or fp32 logits (thresholded as sigmoid > .5), NCHW.  -> fp32 (B, C, 8):
    dice, precision, recall, TP, FP, FN, sum(pred), sum(target)."""
    n_rows, row_len, rep = _rows(pred, target)
    if pred.dtype == torch.bool:
        pred = pred.view(torch.uint8)
    if pred.dtype not in (torch.uint8, torch.float32):
        raise TypeError(f"pred must be bool, uint8 or float32, got {pred.dtype}")
    out = torch.empty(pred.shape[0], pred.shape[1], 8, device=pred.device, dtype=torch.float32)
    _call("tedm_seg_metrics", _ptr(pred, name="pred"), int(pred.dtype == torch.float32), _ptr(target, torch.float32, "target"),
          _ptr(out), n_rows, row_len, rep, _stream())
    return out


def u8_to_unit(src: torch.Tensor) -> torch.Tensor:
    """uint8 image tensor -> fp32 in [0, 1] (ToTensor), same shape."""
    dst = torch.empty(src.shape, device=src.device, dtype=torch.float32)
    _call("tedm_u8_to_unit", _ptr(src, torch.uint8, "src"), _ptr(dst), src.numel(), _stream())
    return dst


def u8_masks_to_label(src: torch.Tensor) -> torch.Tensor:
    """uint8 (B, K, H, W) structure masks -> fp32 (B, 1, H, W) label = min(sum_k (mask_k / 255 > .5), 1)."""
    b, k, h, w = src.shape
    dst = torch.empty(b, 1, h, w, device=src.device, dtype=torch.float32)
    _call("tedm_u8_masks_to_label", _ptr(src, torch.uint8, "src"), _ptr(dst), b, h * w, k, _stream())
    return dst


# ------------------------------------------------------------------------------------------------
# fp32 precision mode (activations: NHWC fp32 tensors of shape (B, H, W, C)); see csrc/fp32_mode.cu
# ------------------------------------------------------------------------------------------------
F32 = torch.float32


def f32_split(x: torch.Tensor):
    """fp32 -> (hi, lo) bf16 with hi + lo == x to 2^-18 relative."""
    hi = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
    lo = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
    _call("tedm_f32_split", _ptr(x, F32, "x"), _ptr(hi), _ptr(lo), x.numel(), _stream())
    return hi, lo


def f32_conv(src0, weight3, mode: int, cout: int, bias=None, src1=None, gn_groups: int = 0, out=None):
    """fp32-grade convolution of fp32 NHWC input(s) on the tcgen05 kernel: operands as bf16 (hi, lo) pairs, product
    hi*hi + lo*hi + hi*lo.  src0 / src1: (hi, lo) pairs from f32_split; weight3: bf16 [cout][taps][3 * (c0 + c1)] =
    cat(w_hi, w_hi, w_lo) along the channel axis.  Returns fp32 out [, GroupNorm partials]."""
    h0, l0 = src0
    if src1 is None:
        return conv_igemm(h0, weight3, mode, cout, bias=bias, gn_groups=gn_groups, out=out, out_dtype=F32,
                          extra=[(l0, False), (h0, False)])
    h1, l1 = src1
    return conv_igemm(h0, weight3, mode, cout, bias=bias, src1=h1, gn_groups=gn_groups, out=out, out_dtype=F32,
                      extra=[(l0, False), (l1, False), (h0, False), (h1, False)])


def f32_stem_conv7x7(x, weight, bias):
    b, cin, h, w = x.shape
    cout = weight.shape[0]
    out = torch.empty(b, h, w, cout, device=x.device, dtype=F32)
    _call("tedm_f32_stem_conv7x7", _ptr(x, F32, "x"), _ptr(weight, F32), _ptr(bias, F32), _ptr(out), b, cin, h, w, cout, _stream())
    return out


def f32_gn_silu(x, gn_partial, gamma, beta, groups: int, eps: float = 1e-5, scale_shift=None, ss_offset: int = 0, residual=None):
    b, h, w, c = x.shape
    out = torch.empty_like(x)
    _call("tedm_f32_gn_silu", _ptr(x, F32, "x"), _ptr(gn_partial, F32), gn_partial.shape[1], _ptr(gamma, F32), _ptr(beta, F32),
          _ptr(scale_shift, F32), scale_shift.shape[1] if scale_shift is not None else 0, ss_offset, _ptr(residual, F32, "residual"),
          _ptr(out), b, h * w, c, groups, eps, _stream())
    return out


def f32_layernorm(x, g, eps: float = 1e-5, residual=None):
    c = x.shape[-1]
    out = torch.empty_like(x)
    _call("tedm_f32_layernorm", _ptr(x, F32, "x"), _ptr(g, F32), _ptr(residual, F32, "residual"), _ptr(out), x.numel() // c, c, eps,
          _stream())
    return out


def f32_linear_attention(qkv, heads: int = 4, dim_head: int = 32, scale: Optional[float] = None):
    b, h, w, c3 = qkv.shape
    n = h * w
    ws = torch.empty(load().tedm_f32_linear_attention_workspace(b, n, heads), device=qkv.device, dtype=F32)
    out = torch.empty(b, h, w, heads * dim_head, device=qkv.device, dtype=F32)
    _call("tedm_f32_linear_attention", _ptr(qkv, F32, "qkv"), _ptr(out), _ptr(ws), b, n, heads, dim_head,
          float(dim_head ** -0.5 if scale is None else scale), _stream())
    return out


def f32_attention(qkv, heads: int = 4, dim_head: int = 32, scale: float = 16.0):
    b, h, w, c3 = qkv.shape
    out = torch.empty(b, h, w, heads * dim_head, device=qkv.device, dtype=F32)
    rnorm = torch.empty(b, 2 * heads * dim_head, device=qkv.device, dtype=F32)
    _call("tedm_f32_attention", _ptr(qkv, F32, "qkv"), _ptr(out), _ptr(rnorm), b, h * w, heads, dim_head, float(scale), _stream())
    return out


def f32_final_conv1x1(x, weight, bias):
    b, h, w, c = x.shape
    od = weight.shape[0]
    out = torch.empty(b, od, h, w, device=x.device, dtype=F32)
    _call("tedm_f32_final_conv1x1", _ptr(x, F32, "x"), _ptr(weight, F32), _ptr(bias, F32), _ptr(out), b, h * w, c, od, _stream())
    return out


def f32_add(a, b):
    out = torch.empty_like(a)
    _call("tedm_f32_add", _ptr(a, F32, "a"), _ptr(b, F32, "b"), _ptr(out), a.numel(), _stream())
    return out


def umma_probe(A, Bm, shifts: Sequence[int], base_offsets: Sequence[int]):
    n = len(shifts)
    out = torch.zeros(n, 128, 64, device=A.device, dtype=torch.float32)
    sh = (C.c_int * n)(*shifts)
    bo = (C.c_int * n)(*base_offsets)
    _call("tedm_debug_umma_probe", _ptr(A, torch.bfloat16), _ptr(Bm, torch.bfloat16), sh, bo, n, _ptr(out), _stream())
    return out
