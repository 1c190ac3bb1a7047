"""Supervised segmentation loop shared by the baseline UNet and the LEDM / LEDMe / TEDM heads
(reference: trainers/train_baseline.py:17-211).

Same control flow, logging keys and checkpoint format as the reference; what changed is where the arithmetic runs:
  * `pred = model(x)` is the native path (tedm_b200.models), the loss is ONE fused BCE kernel whose per-(b, c) means
    also give the per-timestep losses, and TEDM's label repetition is indexed instead of materialised;
  * validation keeps logits, masks and labels on the device: sigmoid > .5, TP / FP / FN and dice / precision /
    recall per row come from one kernel (`tedm_seg_metrics`); the reference moves every prediction to the host first;
  * `autocast` / `GradScaler` are gone: the kernels compute in bf16 with fp32 accumulation regardless, and the
    reference's scaler is never unscaled before `optimizer.step()` (:47-48) -- `scaler` is accepted and ignored;
  * the reference tests `config.experiment == 'datasetDM'` (:24,29), a value `train.py` never sets, so its TEDM runs
    compare (B*S) logits with B labels and fail; here the repetition keys off `shared_weights_over_timesteps` alone.
"""
from __future__ import annotations

import os
from argparse import Namespace
from typing import Dict

import torch
from torch import Tensor

from .. import native as N
from ..autograd import bce_with_logits_rows
from ..models.unet_model import Unet
from ..optim import FusedAdam
from ..parallel import sum_over_ranks
from .utils import TensorboardLogger, dp_optimizer_step, init_distributed, seed_everything


def _n_steps(config, model) -> int:
    return len(model.steps) if getattr(config, "shared_weights_over_timesteps", False) and hasattr(model, "steps") else 1


def save(model, optimizer, config, path, step) -> None:  # trainers/train_base_diffusion.py:164-170
    torch.save({"model_state_dict": model.state_dict(), "optimizer_state_dict": optimizer.state_dict(),
                "config": config, "step": step}, path)


def train(config, model, optimizer, train_dl, val_dl, logger, scaler=None, step: int = 0):
    best_val_loss = float("inf")
    train_losses = []
    n_steps = _n_steps(config, model)
    per_timestep = [] if n_steps > 1 else None
    world = getattr(config, "world_size", 1)
    rank = getattr(config, "rank", 0)
    while True:
        for x, y in train_dl:
            step += 1
            x = x.to(config.device)
            y = y.to(config.device)                       # (B, 1, H, W): row b*S+s of pred is compared with label b
            optimizer.zero_grad()
            pred = model(x)
            expanded_loss = bce_with_logits_rows(pred, y)   # 'b c h w -> b c' means
            loss = expanded_loss.mean()
            loss.backward()
            dp_optimizer_step(optimizer, world)

            train_losses.append(loss.detach())
            if per_timestep is not None:
                per_timestep.append(expanded_loss.detach().reshape(-1, n_steps).mean(0))

            if step % config.log_freq == 0 or config.debug:
                avg_train_loss = torch.stack(train_losses).mean().item()      # the only host sync of the loop
                if rank == 0:
                    print(f"Step {step} - Train loss: {avg_train_loss:.4f}")
                logger.log({"train/loss": avg_train_loss}, step=step)
                if per_timestep is not None:
                    avg = torch.stack(per_timestep).mean(0).cpu()
                    for i, model_step in enumerate(model.steps):
                        logger.log({"train_loss/step_" + str(model_step): avg[i].item()}, step=step)

            if step % config.val_freq == 0 or config.debug:
                val_results = validate(config, model, val_dl)
                logger.log(val_results, step=step)
                if val_results["val/loss"] < best_val_loss and not config.debug:
                    best_val_loss = val_results["val/loss"]
                    if rank == 0:
                        print(f"Step {step} - New best validation loss: {best_val_loss:.4f}, saving model in {config.log_dir}")
                        save(model, optimizer, config, os.path.join(str(config.log_dir), "best_model.pt"), step)
                elif val_results["val/loss"] > best_val_loss * 1.5 and getattr(config, "early_stop", False):
                    print(f"Step {step} - Validation loss increased by more than 50%")
                    return model

            if step >= config.max_steps or config.debug:
                return model


@torch.no_grad()
def validate(config, model, val_dl) -> Dict[str, float]:
    """(train_baseline.py:99-143) val/loss = mean BCE over every pixel of every batch; dice / precision / recall =
    nanmean over (image[, step]) rows."""
    model.eval()
    rows, loss_sum, loss_n = [], None, 0
    for i, (x, y) in enumerate(val_dl):
        x = x.to(config.device)
        y = y.to(config.device)
        pred = model(x).float().contiguous()
        if pred.shape[1] != 1:
            raise NotImplementedError("multi-class (argmax) validation is the reference's BRATS branch; the path is binary")
        rows.append(N.seg_metrics(pred, y))                       # sigmoid(pred) > .5 vs y, per row, on the device
        _, row_mean, _ = N.bce_logits(pred, y)
        s = row_mean.sum()
        loss_sum = s if loss_sum is None else loss_sum + s
        loss_n += row_mean.numel()
        if i + 1 == config.max_val_steps or config.debug:
            break
    m = torch.cat(rows).reshape(-1, 8)[:, :3].double()
    ok = ~torch.isnan(m)
    # [sum of row-mean losses, rows, nansum and non-NaN count of dice / precision / recall]: summed over ranks, so every
    # rank reports the metrics of the WHOLE validation set and takes the same best-model / early-stop decision
    acc = torch.cat([torch.stack([loss_sum.double(), torch.tensor(float(loss_n), device=m.device, dtype=torch.float64)]),
                     torch.where(ok, m, torch.zeros_like(m)).sum(0), ok.double().sum(0)])
    acc = sum_over_ranks(acc).cpu()
    out = {"val/loss": (acc[0] / acc[1]).item(),                # rows are equally long: mean of row means = pixel mean
           "val/dice": (acc[2] / acc[5]).item(), "val/precision": (acc[3] / acc[6]).item(),
           "val/recall": (acc[4] / acc[7]).item()}
    if getattr(config, "rank", 0) == 0:
        print(f"Validation loss: {out['val/loss']:.4f}")
    model.train()
    return out


def _as_bool(x: Tensor) -> Tensor:
    return x if x.dtype == torch.bool else x != 0


def dice(x_hat: Tensor, x: Tensor) -> Tensor:
    """(train_baseline.py:146-149) per-(b, c) dice of a predicted mask against labels, on the device."""
    return N.seg_metrics(_as_bool(x_hat).contiguous(), x.float().contiguous())[..., 0]


def precision(x_hat: Tensor, x: Tensor) -> Tensor:  # :151-155
    return N.seg_metrics(_as_bool(x_hat).contiguous(), x.float().contiguous())[..., 1]


def recall(x_hat: Tensor, x: Tensor) -> Tensor:  # :157-161
    return N.seg_metrics(_as_bool(x_hat).contiguous(), x.float().contiguous())[..., 2]


def build_segmentation_dataloaders(config):
    """JSRT pairs from disk when `config.data_dir` exists, otherwise the deterministic synthetic stand-in."""
    rank, world = getattr(config, "rank", 0), getattr(config, "world_size", 1)
    data_dir = getattr(config, "data_dir", None)
    if getattr(config, "dataset", "JSRT") != "synthetic" and (data_dir is None or not os.path.isdir(str(data_dir))):
        raise FileNotFoundError(f"--data_dir {data_dir} does not exist; pass --dataset synthetic to train on generated pairs "
                                "(real data: <data_dir>/ holds images and masks, the split CSVs are read from --csv_dir or <repo>/data)")
    if getattr(config, "dataset", "JSRT") == "synthetic":
        from ..dataloaders.device_loader import build_synthetic_dataloaders
        return build_synthetic_dataloaders(config.img_size, config.batch_size, 0, labelled=True, device=config.device,
                                           rank=rank, world_size=world, n_labelled_images=config.n_labelled_images)
    if config.dataset != "JSRT":
        raise ValueError(f"Unknown dataset: {config.dataset}")
    from ..dataloaders.JSRT import build_dataloaders
    return build_dataloaders(config.data_dir, config.img_size, config.batch_size, config.num_workers,
                             config.n_labelled_images, device=config.device, rank=rank, world_size=world,
                             **({"csv_dir": config.csv_dir} if getattr(config, "csv_dir", None) else {}))


def write_config(config) -> None:
    os.makedirs(str(config.log_dir), exist_ok=True)
    with open(os.path.join(str(config.log_dir), "config.txt"), "w") as f:
        for k, v in vars(config).items():
            f.write(f"{k}: {v}\n")


def main(config: Namespace) -> None:
    """Baseline: the UNet itself as a segmenter, `timestep=None` (train_baseline.py:164-210)."""
    init_distributed(config)
    write_config(config)
    seed_everything(config.seed)
    model = Unet(config.dim, dim_mults=config.dim_mults, channels=config.channels, out_dim=config.out_channels)
    model.to(config.device)
    model.train()
    optimizer = FusedAdam(model.parameters(), lr=config.lr, weight_decay=config.weight_decay)
    dataloaders = build_segmentation_dataloaders(config)
    train_dl, val_dl = dataloaders["train"], dataloaders["val"]
    print(f"Loaded {len(train_dl.dataset)} training and {len(val_dl.dataset)} validation images")
    logger = TensorboardLogger(config.log_dir, enabled=not config.debug and getattr(config, "rank", 0) == 0)
    train(config, model, optimizer, train_dl, val_dl, logger, None, 0)
