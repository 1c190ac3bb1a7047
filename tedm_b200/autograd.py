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


class BCEWithLogitsRows(torch.autograd.Function):
    """reduce(binary_cross_entropy_with_logits(pred, y, reduction='none'), 'b c h w -> b c', 'mean')
    (trainers/train_baseline.py:44): one fused pass gives the per-(b, c) means and d(mean of them)/d pred; `y` may
    hold B / S images (the TEDM label repetition of :30-31 is indexed, not materialised)."""

    @staticmethod
    def forward(ctx, pred: Tensor, target: Tensor) -> Tensor:
        _, rows, grad = N.bce_logits(pred.detach().float().contiguous(), target.detach().float().contiguous(), want_grad=True)
        ctx.save_for_backward(grad)
        ctx.n_rows = rows.numel()
        return rows

    @staticmethod
    def backward(ctx, drows: Tensor):
        (grad,) = ctx.saved_tensors           # = d mean(rows) / d pred, i.e. every row weighted 1 / n_rows
        w = (drows.reshape(-1) * ctx.n_rows).to(grad.dtype)
        return grad * w.view(grad.shape[0], grad.shape[1], 1, 1), None


def bce_with_logits_rows(pred: Tensor, target: Tensor) -> Tensor:
    """Per-(b, c) mean BCE (B, C); `.mean()` of it is the reference's training loss."""
    if pred.requires_grad:
        return BCEWithLogitsRows.apply(pred, target)
    return N.bce_logits(pred.detach().float().contiguous(), target.detach().float().contiguous())[1]
