"""DDPM process around the B200 UNet, behind the reference's `DiffusionModel` surface
(models/diffusion_model.py:50-301): same constructor (a config Namespace), the same nine fp32
schedule buffers in the state_dict, the same method names.  The schedule tables are produced on
the host by the reference's own fp32 torch-op sequence (bit-exact by construction); q_sample,
the L1/p2 loss and the whole post-UNet part of a reverse step are single fused CUDA kernels.
"""
from __future__ import annotations

import math
from argparse import Namespace
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F
from torch import Tensor, nn

from .. import native as N
from .unet_model import Unet

SCHEDULE_BUFFERS = ("sqrt_alphas_cumprod", "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod",
                    "sqrt_one_minus_alphas_cumprod", "posterior_variance", "posterior_log_variance_clipped",
                    "posterior_mean_coef1", "posterior_mean_coef2", "p2_loss_weight")


def linear_beta_schedule(timesteps: int, start: float = 0.0001, end: float = 0.02) -> Tensor:  # :16-29
    k = 1000 / timesteps
    return torch.linspace(k * start, k * end, timesteps, dtype=torch.float32)


def cosine_beta_schedule(timesteps: int, s: float = 0.008) -> Tensor:  # :32-47
    grid = torch.linspace(0, timesteps, timesteps + 1, dtype=torch.float32)
    abar = torch.cos(((grid / timesteps) + s) / (1 + s) * math.pi * 0.5) ** 2
    abar = abar / abar[0]
    return torch.clip(1 - (abar[1:] / abar[:-1]), 0, 0.999)


def make_schedule(kind: str, timesteps: int, p2_gamma: float, p2_k: float) -> Dict[str, Tensor]:
    """The nine buffers of DiffusionModel.__init__ (:75-115), in registration order."""
    if kind == "linear":
        betas = linear_beta_schedule(timesteps)
    elif kind == "cosine":
        betas = cosine_beta_schedule(timesteps)
    else:
        raise ValueError(f"unknown beta schedule {kind}")
    alphas = 1. - betas
    abar = torch.cumprod(alphas, axis=0)
    abar_prev = F.pad(abar[:-1], (1, 0), value=1.)
    post_var = betas * (1. - abar_prev) / (1. - abar)
    return {
        "sqrt_alphas_cumprod": torch.sqrt(abar),
        "sqrt_recip_alphas_cumprod": torch.sqrt(1. / abar),
        "sqrt_recipm1_alphas_cumprod": torch.sqrt(1. / abar - 1),
        "sqrt_one_minus_alphas_cumprod": torch.sqrt(1. - abar),
        "posterior_variance": post_var,
        "posterior_log_variance_clipped": torch.log(post_var.clamp(min=1e-20)),
        "posterior_mean_coef1": betas * torch.sqrt(abar_prev) / (1. - abar),
        "posterior_mean_coef2": (1. - abar_prev) * torch.sqrt(alphas) / (1. - abar),
        "p2_loss_weight": (p2_k + abar / (1 - abar)) ** (-p2_gamma),
    }


class DiffusionModel(nn.Module):
    def __init__(self, config: Namespace):
        super().__init__()
        self.config = config
        dim: int = self.default("dim", 64)
        dim_mults: List[int] = self.default("dim_mults", [1, 2, 4, 8])
        channels: int = self.default("channels", 1)
        timesteps: int = self.default("timesteps", 1000)
        beta_schedule: str = self.default("beta_schedule", "cosine")
        objective: str = self.default("objective", "pred_noise")
        p2_gamma: float = self.default("p2_loss_weight_gamma", 0.)
        p2_k: float = self.default("p2_loss_weight_k", 1.)
        self.timesteps = timesteps
        self.objective = objective
        self.dynamic_threshold_percentile: float = self.default("dynamic_threshold_percentile", 0.995)
        self.model = Unet(dim, dim_mults=dim_mults, channels=channels, precision=self.default("precision", "bf16"))
        for name, table in make_schedule(beta_schedule, timesteps, p2_gamma, p2_k).items():
            self.register_buffer(name, table)
        self._host_tables: Optional[Tuple[tuple, Dict[str, Tensor]]] = None

    def default(self, val, d):  # :117-118
        return vars(self.config)[val] if val in self.config else d

    def set_precision(self, precision: str) -> "DiffusionModel":
        """'bf16' | 'fp32' (see Unet.set_precision); the DDPM arithmetic around the UNet is fp32 in both."""
        self.model.set_precision(precision)
        return self

    # -- helpers ----------------------------------------------------------------------------------
    def _host(self, name: str) -> Tensor:
        """CPU copies of the schedule buffers (scalar arguments of the sampler kernel come from here,
        so a reverse step needs no device->host sync)."""
        sig = tuple((getattr(self, k).data_ptr(), getattr(self, k)._version) for k in SCHEDULE_BUFFERS)
        if self._host_tables is None or self._host_tables[0] != sig:
            self._host_tables = (sig, {k: getattr(self, k).detach().cpu() for k in SCHEDULE_BUFFERS})
        return self._host_tables[1][name]

    @staticmethod
    def _prep(x: Tensor) -> Tensor:
        if not x.is_cuda:
            raise RuntimeError("tedm_b200.DiffusionModel runs on CUDA (sm_100a) only; there is no CPU fallback")
        return x.detach().float().contiguous()

    def _lincomb(self, table_a: Tensor, x: Tensor, table_b: Tensor, y: Tensor, t: Tensor, normalize: bool = False) -> Tensor:
        """table_a[t]*x + table_b[t]*y per image, unfused mul/mul/add like the reference."""
        t = t.to(device=x.device, dtype=torch.int64).contiguous()
        return N.q_sample(self._prep(x), self._prep(y), t, table_a, table_b, normalize=normalize)

    # -- reference surface ------------------------------------------------------------------------
    def forward_diffusion_model(self, x_0: Tensor, t: Tensor, noise: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
        """q_sample (:176-203): x_t = sqrt(abar_t) x_0 + sqrt(1-abar_t) noise."""
        if noise is None:
            noise = torch.randn_like(x_0)
        x_t = self._lincomb(self.sqrt_alphas_cumprod, x_0, self.sqrt_one_minus_alphas_cumprod, noise, t)
        return x_t, noise

    def forward(self, x_0: Tensor, t: Tensor, cond: Optional[Tensor] = None, noise: Optional[Tensor] = None):
        """(:158-174) rescale to [-1,1] (fused into the q_sample kernel), noise, predict."""
        if noise is None:
            noise = torch.randn_like(x_0)
        x_t = self._lincomb(self.sqrt_alphas_cumprod, x_0, self.sqrt_one_minus_alphas_cumprod, noise, t,
                            normalize=bool(self.config.normalize))
        return self.model(x_t, t, cond), noise

    def train_step(self, x_0: Tensor, cond: Optional[Tensor] = None, t: Optional[Tensor] = None,
                   noise: Optional[Tensor] = None) -> Tensor:
        """(:120-143) L1 between prediction and target, per-image mean, p2 weight, batch mean."""
        n, device = x_0.shape[0], x_0.device
        if t is not None:
            t = t.long().to(device)
        else:
            t = torch.randint(0, self.timesteps, (n,), device=device).long()
        if self.objective not in ("pred_noise", "pred_x_0"):
            raise ValueError(f"unknown objective {self.objective}")
        model_out, noise = self(x_0, t, cond=cond, noise=noise)
        target = noise if self.objective == "pred_noise" else x_0
        return self._loss(model_out, self._prep(target), t)

    def _loss(self, model_out: Tensor, target: Tensor, t: Tensor) -> Tensor:
        if model_out.requires_grad:
            from ..autograd import L1P2Loss
            return L1P2Loss.apply(model_out, target, t, self.p2_loss_weight)
        return N.l1_loss(model_out.contiguous(), target, t.contiguous(), self.p2_loss_weight)[0]

    def val_step(self, x_0: Tensor, cond: Optional[Tensor] = None, t_steps: Optional[int] = None) -> Tensor:
        """(:145-156) mean train_step loss over an evenly spaced timestep grid.  The grid is batched along the
        image axis in chunks instead of one UNet forward per timestep."""
        if not t_steps:
            t_steps = self.timesteps
        step = self.timesteps // t_steps
        n, device = x_0.shape[0], x_0.device
        ts = list(range(0, self.timesteps, step))
        per_call = max(1, 256 // n)
        losses = []
        with torch.no_grad():
            for i in range(0, len(ts), per_call):
                chunk = ts[i:i + per_call]
                t = torch.tensor(chunk, device=device, dtype=torch.long).repeat_interleave(n)
                xr = x_0.repeat(len(chunk), 1, 1, 1)
                model_out, noise = self(xr, t, cond=cond)
                target = noise if self.objective == "pred_noise" else xr
                _, per_img, _ = N.l1_loss(model_out.contiguous(), self._prep(target), t, self.p2_loss_weight)
                losses.append(per_img.reshape(len(chunk), n).mean(dim=1))
        return torch.cat(losses).mean()

    @torch.no_grad()
    def sample_timestep(self, x_t: Tensor, t: int, cond: Optional[Tensor] = None, noise: Optional[Tensor] = None) -> Tensor:
        """One ancestral step (:205-219 with :221-286 inlined): UNet, then ONE kernel for x0_hat, exact
        dynamic-threshold quantile, clip/scale, posterior mean and noise add."""
        mean_or_sample, _, _ = self._reverse(x_t, int(t), cond, noise, add_noise=True)
        return mean_or_sample

    def _reverse(self, x_t: Tensor, t: int, cond, noise: Optional[Tensor], add_noise: bool, want_x0: bool = False):
        if self.objective != "pred_noise":
            raise ValueError("only objective='pred_noise' can sample (as in the reference, :247-255)")
        x_t = self._prep(x_t)
        n = x_t.shape[0]
        tt = torch.full((n,), t, device=x_t.device, dtype=torch.long)
        eps = self.model(x_t, tt, cond)
        chw = x_t[0].numel()
        # torch.quantile: rank = q * (n-1) in the input dtype, lerp between floor(rank) and the next one
        rank = torch.tensor(self.dynamic_threshold_percentile, dtype=torch.float32) * (chw - 1)
        k_lo = int(torch.floor(rank).item())
        weight = float((rank - torch.floor(rank)).item())
        z = None
        if add_noise and t > 0:
            z = torch.randn_like(x_t) if noise is None else self._prep(noise)
        sigma = float((0.5 * self._host("posterior_log_variance_clipped")[t]).exp())
        out, x0h, s = N.sampler_step(
            x_t, eps.contiguous(), z, float(self._host("sqrt_recip_alphas_cumprod")[t]),
            float(self._host("sqrt_recipm1_alphas_cumprod")[t]), float(self._host("posterior_mean_coef1")[t]),
            float(self._host("posterior_mean_coef2")[t]), sigma, k_lo, weight, want_x0=want_x0)
        return out, x0h, eps

    def reverse_tables(self, device) -> Tensor:
        """(T, 5) device table of the per-step scalars of a reverse step: sqrt_recip_ac, sqrt_recipm1_ac, coef1, coef2, sigma
        (sigma = exp(0.5 logvar_t), 0 at t = 0 where no noise is added)."""
        sig = (0.5 * self._host("posterior_log_variance_clipped")).exp()
        sig[0] = 0.0
        tab = torch.stack([self._host("sqrt_recip_alphas_cumprod"), self._host("sqrt_recipm1_alphas_cumprod"),
                           self._host("posterior_mean_coef1"), self._host("posterior_mean_coef2"), sig], dim=1)
        return tab.float().contiguous().to(device)

    def p_mean_variance(self, x_t: Tensor, t: Tensor, clip_denoised: bool = True, cond: Optional[Tensor] = None):
        """(:221-235) -> (model_mean, posterior_log_variance, pred_x_0).  All images must share the timestep."""
        if not clip_denoised:
            raise NotImplementedError("clip_denoised=False is never used by the reference")
        tv = int(t.flatten()[0].item())
        mean, x0h, _ = self._reverse(x_t, tv, cond, None, add_noise=False, want_x0=True)
        logvar = self.posterior_log_variance_clipped[t.long()].reshape(-1, 1, 1, 1)
        return mean, logvar, x0h

    def model_predictions(self, x_t: Tensor, t: Tensor, cond: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
        """(:237-257) -> (pred_noise, pred_x_0)."""
        if self.objective != "pred_noise":
            raise ValueError("only objective='pred_noise' is supported here (as in the reference sampler)")
        pred_noise = self.model(x_t, t, cond)
        return pred_noise, self.predict_x_0_from_noise(x_t, t, pred_noise)

    def q_posterior(self, x_start: Tensor, x_t: Tensor, t: Tensor) -> Tuple[Tensor, Tensor]:  # :259-267
        mean = self._lincomb(self.posterior_mean_coef1, x_start, self.posterior_mean_coef2, x_t, t)
        return mean, self.posterior_log_variance_clipped[t.long()].reshape(-1, 1, 1, 1)

    def predict_x_0_from_noise(self, x_t: Tensor, t: Tensor, noise: Tensor) -> Tensor:  # :269-286
        neg = -self.sqrt_recipm1_alphas_cumprod
        return self._lincomb(self.sqrt_recip_alphas_cumprod, x_t, neg, noise, t)

    def predict_noise_from_x_0(self, x_t: Tensor, t: Tensor, x_0: Tensor) -> Tensor:  # :288-301
        a = self.sqrt_recip_alphas_cumprod[t.long()].reshape(-1, 1, 1, 1)
        b = self.sqrt_recipm1_alphas_cumprod[t.long()].reshape(-1, 1, 1, 1)
        return (a * x_t - x_0) / b
