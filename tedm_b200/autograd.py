"""Autograd glue for the DDPM loss (models/diffusion_model.py:138-143): the fused kernel produces the
loss and d loss / d prediction in one pass; backward only scales that stored gradient."""
from __future__ import annotations

import torch
from torch import Tensor

from . import native as N


class L1P2Loss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred: Tensor, target: Tensor, t: Tensor, p2_weight: Tensor) -> Tensor:
        loss, _, grad = N.l1_loss(pred.detach().contiguous(), target, t.contiguous(), p2_weight, want_grad=True)
        ctx.save_for_backward(grad)
        return loss.clone()

    @staticmethod
    def backward(ctx, dloss: Tensor):
        (grad,) = ctx.saved_tensors
        return grad * dloss, None, None, None
