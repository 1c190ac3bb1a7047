"""CPU oracle for the TEDM hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A functional, fp32, torch-on-CPU restatement of the reference algorithm for the
data-parallel hot path of mmr12/TEDM.  Nothing in ``tedm_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs
(``cpu_baseline`` / ``--impl reference``) do.

Parity pin
----------
The reference ships no golden vectors or tests.  The pin is
therefore: outputs of the *live* reference (``/root/reference`` imported in the
build container) on synthetic weights/inputs, committed as fixtures in
``tests/golden/*.npz`` by ``tests/golden/make_golden.py``.  ``tests/test_oracle.py``
checks this file against those fixtures (and, when ``/root/reference`` is present,
against the reference directly).

Everything here works on a flat ``state_dict`` that uses the reference's own
parameter names (``model.downs.0.0.block1.proj.weight`` ...), so a reference
checkpoint can be fed in unchanged.  Each function cites the reference lines it
follows.  An optional ``store`` callable emulates the bf16 storage points of the
CUDA path (identity == pure fp32 oracle).
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
StateDict = Dict[str, Tensor]


def _ident(x: Tensor) -> Tensor:
    return x


def bf16_store(x: Tensor) -> Tensor:
    """Round-trip through bf16: what the CUDA path does at every activation store."""
    return x.to(torch.bfloat16).to(torch.float32)


# --------------------------------------------------------------------------- #
# (a1, a2) noise schedule                    models/diffusion_model.py:16-47,82-115
# --------------------------------------------------------------------------- #
def beta_schedule(kind: str, T: int) -> Tensor:
    if kind == "linear":  # diffusion_model.py:16-29
        k = 1000 / T
        return torch.linspace(k * 1e-4, k * 0.02, T, dtype=torch.float32)
    if kind == "cosine":  # diffusion_model.py:32-47
        grid = torch.linspace(0, T, T + 1, dtype=torch.float32)
        abar = torch.cos(((grid / T) + 0.008) / 1.008 * math.pi * 0.5) ** 2
        abar = abar / abar[0]
        return torch.clip(1 - (abar[1:] / abar[:-1]), 0, 0.999)
    raise ValueError(f"unknown beta schedule {kind}")


SCHEDULE_KEYS = (
    "sqrt_alphas_cumprod", "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod",
    "sqrt_one_minus_alphas_cumprod", "posterior_variance", "posterior_log_variance_clipped",
    "posterior_mean_coef1", "posterior_mean_coef2", "p2_loss_weight",
)


def schedule_tables(kind: str = "cosine", T: int = 1000, p2_gamma: float = 0.0,
                    p2_k: float = 1.0) -> Dict[str, Tensor]:
    """The nine fp32 buffers of DiffusionModel.__init__ (diffusion_model.py:82-115)."""
    beta = beta_schedule(kind, T)
    alpha = 1.0 - beta
    abar = torch.cumprod(alpha, dim=0)
    abar_prev = F.pad(abar[:-1], (1, 0), value=1.0)
    post_var = beta * (1.0 - abar_prev) / (1.0 - abar)
    return {
        "sqrt_alphas_cumprod": torch.sqrt(abar),
        "sqrt_recip_alphas_cumprod": torch.sqrt(1.0 / abar),
        "sqrt_recipm1_alphas_cumprod": torch.sqrt(1.0 / abar - 1),
        "sqrt_one_minus_alphas_cumprod": torch.sqrt(1.0 - abar),
        "posterior_variance": post_var,
        "posterior_log_variance_clipped": torch.log(post_var.clamp(min=1e-20)),
        "posterior_mean_coef1": beta * torch.sqrt(abar_prev) / (1.0 - abar),
        "posterior_mean_coef2": (1.0 - abar_prev) * torch.sqrt(alpha) / (1.0 - abar),
        "p2_loss_weight": (p2_k + abar / (1 - abar)) ** (-p2_gamma),
    }


def gather_t(table: Tensor, t: Tensor) -> Tensor:
    """trainers/utils.py:48-59 -- table[t] broadcast as (B,1,1,1)."""
    return table[t.long()].reshape(-1, 1, 1, 1)


# --------------------------------------------------------------------------- #
# (a4, a5) q_sample                          diffusion_model.py:158-203, utils.py:28-33
# --------------------------------------------------------------------------- #
def q_sample(tables: Dict[str, Tensor], x0: Tensor, t: Tensor, noise: Tensor,
             normalize: bool = False) -> Tensor:
    if normalize:
        x0 = x0 * 2 - 1
    return gather_t(tables["sqrt_alphas_cumprod"], t) * x0 + \
        gather_t(tables["sqrt_one_minus_alphas_cumprod"], t) * noise


# --------------------------------------------------------------------------- #
# UNet parameter inventory                   models/unet_model.py:246-331
# --------------------------------------------------------------------------- #
def unet_param_shapes(dim: int = 64, dim_mults: Sequence[int] = (1, 2, 4, 8), channels: int = 1,
                      out_dim: Optional[int] = None, init_dim: Optional[int] = None,
                      prefix: str = "") -> Dict[str, Tuple[int, ...]]:
    """Key -> shape of Unet.state_dict() (learned_sinusoidal_cond=False), in module order."""
    init_dim = init_dim or dim
    widths = [init_dim] + [dim * m for m in dim_mults]
    pairs = list(zip(widths[:-1], widths[1:]))
    tdim = dim * 4
    out: Dict[str, Tuple[int, ...]] = {}

    def conv(name, cin, cout, k, bias=True):
        out[f"{prefix}{name}.weight"] = (cout, cin, k, k)
        if bias:
            out[f"{prefix}{name}.bias"] = (cout,)

    def resblock(name, cin, cout):
        out[f"{prefix}{name}.time_mlp.1.weight"] = (2 * cout, tdim)
        out[f"{prefix}{name}.time_mlp.1.bias"] = (2 * cout,)
        for blk, ci in (("block1", cin), ("block2", cout)):
            conv(f"{name}.{blk}.proj", ci, cout, 3)
            out[f"{prefix}{name}.{blk}.norm.weight"] = (cout,)
            out[f"{prefix}{name}.{blk}.norm.bias"] = (cout,)
        if cin != cout:
            conv(f"{name}.res_conv", cin, cout, 1)

    def linattn(name, c):
        conv(f"{name}.fn.fn.to_qkv", c, 384, 1, bias=False)
        conv(f"{name}.fn.fn.to_out.0", 128, c, 1)
        out[f"{prefix}{name}.fn.fn.to_out.1.g"] = (1, c, 1, 1)
        out[f"{prefix}{name}.fn.norm.g"] = (1, c, 1, 1)

    conv("init_conv", channels, init_dim, 7)
    out[f"{prefix}time_mlp.1.weight"] = (tdim, dim)
    out[f"{prefix}time_mlp.1.bias"] = (tdim,)
    out[f"{prefix}time_mlp.3.weight"] = (tdim, tdim)
    out[f"{prefix}time_mlp.3.bias"] = (tdim,)
    for i, (ci, co) in enumerate(pairs):
        last = i == len(pairs) - 1
        resblock(f"downs.{i}.0", ci, ci)
        resblock(f"downs.{i}.1", ci, ci)
        linattn(f"downs.{i}.2", ci)
        conv(f"downs.{i}.3", ci, co, 3 if last else 4)
    mid = widths[-1]
    resblock("mid_block1", mid, mid)
    conv("mid_attn.fn.fn.to_qkv", mid, 384, 1, bias=False)
    conv("mid_attn.fn.fn.to_out", 128, mid, 1)
    out[f"{prefix}mid_attn.fn.norm.g"] = (1, mid, 1, 1)
    resblock("mid_block2", mid, mid)
    for i, (ci, co) in enumerate(reversed(pairs)):
        last = i == len(pairs) - 1
        resblock(f"ups.{i}.0", co + ci, co)
        resblock(f"ups.{i}.1", co + ci, co)
        linattn(f"ups.{i}.2", co)
        conv(f"ups.{i}.3" if last else f"ups.{i}.3.1", co, ci, 3)
    resblock("final_res_block", dim * 2, dim)
    conv("final_conv", dim, out_dim or channels, 1)
    return out


def head_param_shapes(n_steps: int, shared: bool, prefix: str = "classifier.") -> Dict[str, Tuple[int, ...]]:
    """datasetDM_model.py:57-64 (LEDM: indices 0,2,3,5,6) / train_datasetDM.py:30-42 (TEDM: 1,3,4,6,7)."""
    o = 1 if shared else 0
    cin = 960 if shared else 960 * n_steps
    shp: Dict[str, Tuple[int, ...]] = {}
    for idx, (ci, co) in zip((0, 3, 6), ((cin, 128), (128, 32), (32, 1))):
        shp[f"{prefix}{idx + o}.weight"] = (co, ci, 1, 1)
        shp[f"{prefix}{idx + o}.bias"] = (co,)
    for idx, c in ((2, 128), (5, 32)):
        for nm in ("weight", "bias", "running_mean", "running_var"):
            shp[f"{prefix}{idx + o}.{nm}"] = (c,)
        shp[f"{prefix}{idx + o}.num_batches_tracked"] = ()
    return shp


# --------------------------------------------------------------------------- #
# UNet forward                               models/unet_model.py:52-368
# --------------------------------------------------------------------------- #
def time_embedding(sd: StateDict, p: str, t: Tensor, dim: int) -> Tensor:
    """SinusoidalPosEmb + time_mlp (unet_model.py:76-93, 287-292)."""
    half = dim // 2
    freq = torch.exp(torch.arange(half, device=t.device) * -(math.log(10000) / (half - 1)))
    ang = t[:, None] * freq[None, :]
    e = torch.cat((ang.sin(), ang.cos()), dim=-1)
    e = F.linear(e, sd[p + "time_mlp.1.weight"], sd[p + "time_mlp.1.bias"])
    e = F.gelu(e)
    return F.linear(e, sd[p + "time_mlp.3.weight"], sd[p + "time_mlp.3.bias"])


def _chan_layernorm(x: Tensor, g: Tensor, eps: float) -> Tensor:
    """unet_model.py:52-61: per-pixel norm over channels, biased variance, gain only."""
    mu = x.mean(dim=1, keepdim=True)
    var = x.var(dim=1, unbiased=False, keepdim=True)
    return (x - mu) * torch.rsqrt(var + eps) * g


def _block(sd: StateDict, p: str, x: Tensor, groups: int, ss, st) -> Tensor:
    """Block (unet_model.py:119-135): conv3x3 -> GroupNorm -> [*(scale+1)+shift] -> SiLU."""
    y = st(F.conv2d(x, sd[p + "proj.weight"], sd[p + "proj.bias"], padding=1))
    y = F.group_norm(y, groups, sd[p + "norm.weight"], sd[p + "norm.bias"], eps=1e-5)
    if ss is not None:
        y = y * (ss[0] + 1) + ss[1]
    return F.silu(y)


def _resblock(sd: StateDict, p: str, x: Tensor, temb: Optional[Tensor], groups: int, st) -> Tensor:
    """ResnetBlock (unet_model.py:138-175)."""
    ss = None
    if temb is not None:
        e = F.linear(F.silu(temb), sd[p + "time_mlp.1.weight"], sd[p + "time_mlp.1.bias"])
        ss = e[:, :, None, None].chunk(2, dim=1)
    h = st(_block(sd, p + "block1.", x, groups, ss, st))
    h = _block(sd, p + "block2.", h, groups, None, st)
    if (p + "res_conv.weight") in sd:
        x = st(F.conv2d(x, sd[p + "res_conv.weight"], sd[p + "res_conv.bias"]))
    return st(h + x)


def _linear_attention(sd: StateDict, p: str, x: Tensor, ln_eps: float, st, heads: int = 4) -> Tensor:
    """Residual(PreNorm(LinearAttention)) (unet_model.py:29-36, 64-73, 178-210)."""
    b, c, hh, ww = x.shape
    n = hh * ww
    y = st(_chan_layernorm(x, sd[p + "fn.norm.g"], ln_eps))
    qkv = st(F.conv2d(y, sd[p + "fn.fn.to_qkv.weight"]))
    q, k, v = (z.reshape(b, heads, -1, n) for z in qkv.chunk(3, dim=1))
    q = q.softmax(dim=-2) * (q.shape[2] ** -0.5)
    k = k.softmax(dim=-1)
    v = v / n
    ctx = torch.einsum("bhdn,bhen->bhde", k, v)
    o = torch.einsum("bhde,bhdn->bhen", ctx, q).reshape(b, -1, hh, ww)
    o = st(o)
    o = st(F.conv2d(o, sd[p + "fn.fn.to_out.0.weight"], sd[p + "fn.fn.to_out.0.bias"]))
    o = _chan_layernorm(o, sd[p + "fn.fn.to_out.1.g"], ln_eps)
    return st(o + x)


def _mid_attention(sd: StateDict, p: str, x: Tensor, ln_eps: float, st, heads: int = 4,
                   scale: float = 16.0) -> Tensor:
    """Residual(PreNorm(Attention)) (unet_model.py:213-241); q,k are L2-normalised over n."""
    b, c, hh, ww = x.shape
    n = hh * ww
    y = st(_chan_layernorm(x, sd[p + "fn.norm.g"], ln_eps))
    qkv = st(F.conv2d(y, sd[p + "fn.fn.to_qkv.weight"]))
    q, k, v = (z.reshape(b, heads, -1, n) for z in qkv.chunk(3, dim=1))
    q = F.normalize(q, dim=-1)
    k = F.normalize(k, dim=-1)
    sim = torch.einsum("bhdi,bhdj->bhij", q, k) * scale
    att = sim.softmax(dim=-1)
    o = torch.einsum("bhij,bhdj->bhid", att, v)          # (b, h, n, d)
    o = o.permute(0, 1, 3, 2).reshape(b, -1, hh, ww)
    o = st(o)
    o = F.conv2d(o, sd[p + "fn.fn.to_out.weight"], sd[p + "fn.fn.to_out.bias"])
    return st(o + x)


def unet_forward(sd: StateDict, x: Tensor, t: Optional[Tensor], *, prefix: str = "",
                 dim: int = 64, groups: int = 8, store: Callable[[Tensor], Tensor] = _ident,
                 ln_eps: float = 1e-5, want_features: bool = False,
                 skip_tail: bool = False):
    """Unet.forward (unet_model.py:333-368).  Returns out, or (out, [4 decoder feature maps]).

    Features are the outputs of ``ups[i][2]`` -- what DatasetDM's hooks capture
    (datasetDM_model.py:50-53).  ``skip_tail`` stops after the last hooked map
    (out is None): the part of the net extract_features never uses.
    """
    p, st = prefix, store
    n_levels = sum(1 for k in sd if k.startswith(p + "downs.") and k.endswith(".3.weight"))
    temb = time_embedding(sd, p, t, dim) if t is not None else None
    x = st(F.conv2d(x, sd[p + "init_conv.weight"], sd[p + "init_conv.bias"], padding=3))
    stem = x
    skips: List[Tensor] = []
    for i in range(n_levels):
        q = f"{p}downs.{i}."
        x = _resblock(sd, q + "0.", x, temb, groups, st)
        skips.append(x)
        x = _resblock(sd, q + "1.", x, temb, groups, st)
        x = _linear_attention(sd, q + "2.", x, ln_eps, st)
        skips.append(x)
        w = sd[q + "3.weight"]
        x = st(F.conv2d(x, w, sd[q + "3.bias"], stride=2 if w.shape[-1] == 4 else 1, padding=1))
    x = _resblock(sd, p + "mid_block1.", x, temb, groups, st)
    x = _mid_attention(sd, p + "mid_attn.", x, ln_eps, st)
    x = _resblock(sd, p + "mid_block2.", x, temb, groups, st)
    feats: List[Tensor] = []
    for i in range(n_levels):
        q = f"{p}ups.{i}."
        x = _resblock(sd, q + "0.", torch.cat((x, skips.pop()), dim=1), temb, groups, st)
        x = _resblock(sd, q + "1.", torch.cat((x, skips.pop()), dim=1), temb, groups, st)
        x = _linear_attention(sd, q + "2.", x, ln_eps, st)
        feats.append(x)
        if skip_tail and i == n_levels - 1:
            return None, feats
        if (q + "3.1.weight") in sd:      # Upsample: nearest x2 then conv3x3 (unet_model.py:39-44)
            x = F.interpolate(x, scale_factor=2, mode="nearest")
            x = st(F.conv2d(x, sd[q + "3.1.weight"], sd[q + "3.1.bias"], padding=1))
        else:
            x = st(F.conv2d(x, sd[q + "3.weight"], sd[q + "3.bias"], padding=1))
    x = _resblock(sd, p + "final_res_block.", torch.cat((x, stem), dim=1), temb, groups, st)
    out = F.conv2d(x, sd[p + "final_conv.weight"], sd[p + "final_conv.bias"])
    return (out, feats) if want_features else out


# --------------------------------------------------------------------------- #
# (a6-a8) DDPM training loss                 diffusion_model.py:120-174
# --------------------------------------------------------------------------- #
def ddpm_loss(sd: StateDict, x0: Tensor, t: Tensor, noise: Tensor, *, normalize: bool = True,
              objective: str = "pred_noise", store=_ident, **kw) -> Tensor:
    xin = x0 * 2 - 1 if normalize else x0
    x_t = q_sample(sd, xin, t, noise)
    pred = unet_forward(sd, x_t, t, prefix="model.", store=store, **kw)
    if objective == "pred_noise":
        target = noise
    elif objective == "pred_x_0":
        target = x0          # NB: the reference uses the *un-normalised* x_0 (diffusion_model.py:134)
    else:
        raise ValueError(f"unknown objective {objective}")
    per_img = (pred - target).abs().flatten(1).mean(dim=1)
    return (per_img * sd["p2_loss_weight"][t.long()]).mean()


def l1_p2_loss(pred: Tensor, target: Tensor, w_t: Tensor) -> Tensor:
    """diffusion_model.py:138-143 on already computed tensors."""
    return ((pred - target).abs().flatten(1).mean(dim=1) * w_t).mean()


# --------------------------------------------------------------------------- #
# (a9) one ancestral sampling step           diffusion_model.py:205-286
# --------------------------------------------------------------------------- #
def sampler_update(tb: Dict[str, Tensor], x_t: Tensor, eps: Tensor, t: int, z: Optional[Tensor],
                   pct: float = 0.995) -> Tuple[Tensor, Tensor]:
    """Everything in sample_timestep after the UNet call.  Returns (x_{t-1}, clipped x0_hat)."""
    tt = torch.full((x_t.shape[0],), t, dtype=torch.long, device=x_t.device)
    x0h = gather_t(tb["sqrt_recip_alphas_cumprod"], tt) * x_t - \
        gather_t(tb["sqrt_recipm1_alphas_cumprod"], tt) * eps
    s = torch.quantile(x0h.flatten(1).abs(), pct, dim=1)
    s = torch.max(s, torch.tensor(1.0, device=s.device))[:, None, None, None]
    x0h = torch.clip(x0h, -s, s) / s
    mean = gather_t(tb["posterior_mean_coef1"], tt) * x0h + gather_t(tb["posterior_mean_coef2"], tt) * x_t
    logvar = gather_t(tb["posterior_log_variance_clipped"], tt)
    if t > 0 and z is not None:
        return mean + (0.5 * logvar).exp() * z, x0h
    return mean, x0h


def sample_timestep(sd: StateDict, x_t: Tensor, t: int, z: Optional[Tensor], store=_ident, **kw) -> Tensor:
    tt = torch.full((x_t.shape[0],), t, dtype=torch.long, device=x_t.device)
    eps = unet_forward(sd, x_t, tt, prefix="model.", store=store, **kw)
    return sampler_update(sd, x_t, eps, t, z)[0]


# --------------------------------------------------------------------------- #
# (a19-a22) DatasetDM features, head, ensemble
# --------------------------------------------------------------------------- #
def extract_feature_maps(sd: StateDict, x0: Tensor, steps: Sequence[int], noises: Sequence[Tensor],
                         store=_ident, **kw) -> List[List[Tensor]]:
    """datasetDM_model.py:67-83 without the upsample/concat: [step][level] native-resolution maps.
    NB: x0 is *not* rescaled to [-1,1] here (datasetDM_model.py:76)."""
    maps = []
    for s, nz in zip(steps, noises):
        t = torch.full((x0.shape[0],), int(s), dtype=torch.long, device=x0.device)
        x_t = q_sample(sd, x0, t, nz)
        _, f = unet_forward(sd, x_t, t, prefix="diffusion_model.model." if
                            any(k.startswith("diffusion_model.") for k in sd) else "model.",
                            store=store, want_features=True, **kw)
        maps.append(f)
    return maps


def concat_features(maps: List[List[Tensor]], size: int) -> Tensor:
    """datasetDM_model.py:80-83: nearest-resize every map to size^2, concat [step0: l0..l3, step1: ...]."""
    return torch.cat([F.interpolate(f, size=[size, size]) for lv in maps for f in lv], dim=1)


def head_forward(sd: StateDict, feats: Tensor, n_steps: int, shared: bool, training: bool = False,
                 prefix: str = "classifier.") -> Tensor:
    """Per-pixel MLP: conv1x1 -> ReLU -> BatchNorm, twice, then conv1x1.
    shared=True is the TEDM head (train_datasetDM.py:30-42): 'b (step act) h w -> (b step) act h w'."""
    o = 1 if shared else 0
    x = feats
    if shared:
        b, c, h, w = x.shape
        x = x.reshape(b, n_steps, c // n_steps, h, w).reshape(b * n_steps, c // n_steps, h, w)
    for ci, bi in ((0, 2), (3, 5)):
        x = F.relu(F.conv2d(x, sd[f"{prefix}{ci + o}.weight"], sd[f"{prefix}{ci + o}.bias"]))
        x = F.batch_norm(x, sd[f"{prefix}{bi + o}.running_mean"], sd[f"{prefix}{bi + o}.running_var"],
                         sd[f"{prefix}{bi + o}.weight"], sd[f"{prefix}{bi + o}.bias"],
                         training=training, momentum=0.0 if training else 0.1, eps=1e-5)
    return F.conv2d(x, sd[f"{prefix}{6 + o}.weight"], sd[f"{prefix}{6 + o}.bias"])


def ensemble_mask(logits: Tensor, n_steps: int) -> Tuple[Tensor, Tensor]:
    """testing_shared_weights.py:113,120,133-138 / app.py:79: sigmoid -> mean over step -> > .5."""
    pr = torch.sigmoid(logits)
    pr = pr.reshape(-1, n_steps, *pr.shape[1:]).mean(dim=1)
    return pr > 0.5, pr


def tedm_segment(sd: StateDict, x0: Tensor, steps: Sequence[int], noises: Sequence[Tensor],
                 shared: bool = True, store=_ident, **kw):
    """End-to-end TEDM/LEDM inference: returns (logits, mask, mean prob)."""
    maps = extract_feature_maps(sd, x0, steps, noises, store=store, **kw)
    logits = head_forward(sd, concat_features(maps, x0.shape[-1]), len(steps), shared)
    if shared:
        mask, pr = ensemble_mask(logits, len(steps))
    else:
        pr = torch.sigmoid(logits)
        mask = pr > 0.5
    return logits, mask, pr


def bce_with_logits_loss(logits: Tensor, y: Tensor) -> Tensor:
    """train_baseline.py:44-45."""
    return F.binary_cross_entropy_with_logits(logits, y, reduction="none").mean(dim=(2, 3)).mean()


def bce_rows(logits: Tensor, y: Tensor, n_steps: int = 1) -> Tensor:
    """Per-(b, c) mean BCE with the TEDM label repetition (train_baseline.py:30-31,44): (B*S, C)."""
    if n_steps > 1:
        y = y.repeat_interleave(n_steps, dim=0)          # repeat(y, 'b c h w -> (b step) c h w')
    return F.binary_cross_entropy_with_logits(logits, y, reduction="none").mean(dim=(2, 3))


def seg_metrics(y_hat: Tensor, y: Tensor, n_steps: int = 1) -> Dict[str, Tensor]:
    """dice / precision / recall per (b, c) (train_baseline.py:146-161), on float labels exactly as written there:
    logical_and treats any nonzero label as True, `1 - x` is the negative class, and the dice denominator adds the
    label VALUES (not their count)."""
    y_hat = y_hat.bool()
    y = y.float()
    if n_steps > 1:
        y = y.repeat_interleave(n_steps, dim=0)
    red = lambda a: a.sum(dim=(2, 3))
    tp = red(torch.logical_and(y, y_hat))
    fp = red(torch.logical_and(1 - y, y_hat))
    fn = red(torch.logical_and(y, ~y_hat))
    return {"dice": 2 * tp / (red(y_hat) + red(y)), "precision": tp / (tp + fp), "recall": tp / (tp + fn)}


def to_tensor_u8(img_u8: Tensor) -> Tensor:
    """torchvision ToTensor on an 8-bit 'L' image (dataloaders/CXR14.py:67-70, JSRT.py:62-65): u8 -> fp32 / 255."""
    return img_u8.to(torch.float32).div(255)


def jsrt_label(masks_u8: Tensor) -> Tensor:
    """(dataloaders/JSRT.py:67-82) masks (K, H, W) uint8 -> (1, H, W): sum_k (ToTensor(mask_k) > .5), made binary
    when the structures overlap."""
    label = torch.stack([(to_tensor_u8(m)[None] > .5).float() for m in masks_u8]).sum(0)
    if (label > 1).sum() > 0:
        label = (label > .5).float()
    return label
