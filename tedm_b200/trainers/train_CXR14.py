"""DDPM pre-training loop (reference: trainers/train_CXR14.py:16-159; the JSRT variant
trainers/train_base_diffusion.py has the same body).

`loss = model.train_step(x); loss.backward(); optimizer.step()` is kept verbatim for odd-sized batches; full
batches go through `GraphedTrainStep`, which replays the same step from two CUDA graphs (with ONE NCCL all-reduce of
the flat gradient arena between them under data parallelism).  Checkpoints use the reference's dict format."""
from __future__ import annotations

import os
from argparse import Namespace
from pathlib import Path

import torch

from ..models.diffusion_model import DiffusionModel
from ..optim import FusedAdam
from ..parallel import sum_over_ranks
from ..train import GraphedTrainStep
from .train_baseline import save, write_config
from .utils import (TensorboardLogger, compare_configs, dp_optimizer_step, init_distributed, sample_plot_image,
                    seed_everything)


def train(config, model, optimizer, train_loader, val_loader, logger, scaler=None, step: int = 0):
    best_val_loss = float("inf")
    train_losses = []
    world, rank = getattr(config, "world_size", 1), getattr(config, "rank", 0)
    graphed = None
    while True:
        for x in train_loader:
            step += 1
            x = x.to(config.device)
            use_graph = getattr(config, "cuda_graph", True) and not config.debug and x.shape[0] == config.batch_size
            if use_graph:
                if graphed is None:
                    graphed = GraphedTrainStep(model, optimizer, x)
                loss = graphed(x).clone()
            else:                                   # the reference's loop body (train_CXR14.py:28-40)
                optimizer.zero_grad()
                loss = model.train_step(x)
                loss.backward()
                dp_optimizer_step(optimizer, world)
            train_losses.append(loss.detach())

            if step % config.log_freq == 0 or config.debug:
                avg_train_loss = torch.stack(train_losses).mean().item()
                if rank == 0:
                    print(f"Step {step} - Train loss: {avg_train_loss:.4f}")
                logger.log({"train/loss": avg_train_loss}, step=step)

            if step % config.val_freq == 0 or config.debug:
                val_results = validate(config, model, val_loader)
                logger.log(val_results, step=step)
                if val_results["val/loss"] < best_val_loss and not config.debug:
                    best_val_loss = val_results["val/loss"]
                    if rank == 0:
                        print(f"Step {step} - New best validation loss: {best_val_loss:.4f}, saving model in {config.log_dir}")
                        save(model, optimizer, config, os.path.join(str(config.log_dir), "best_model.pt"), step)

            if step >= config.max_steps or config.debug:
                if graphed is not None:
                    graphed.close()             # captured NCCL work must be gone before the process group is torn down
                return model


@torch.no_grad()
def validate(config, model, val_loader):
    """(train_CXR14.py:63-93) mean `train_step` loss over the validation batches + a grid of sampled images."""
    model.eval()
    losses = []
    for i, x in enumerate(val_loader):
        x = x.to(config.device)
        losses.append(model.train_step(x))
        if i + 1 == config.max_val_steps or config.debug:
            break
    acc = torch.stack([torch.stack(losses).double().sum(), torch.tensor(float(len(losses)), device=losses[0].device, dtype=torch.float64)])
    acc = sum_over_ranks(acc).cpu()                  # every rank logs / compares the loss of the whole validation set
    avg_loss = (acc[0] / acc[1]).item()
    if getattr(config, "rank", 0) == 0:
        print(f"Validation loss: {avg_loss:.4f}")
    out = {"val/loss": avg_loss}
    n_imgs = config.n_sampled_imgs if not config.debug else min(1, config.n_sampled_imgs)
    if n_imgs > 0 and getattr(config, "sample_at_validation", True):
        out["val/sampled images"] = sample_plot_image(model, config.timesteps, config.img_size, n_imgs)
    model.train()
    return out


def load(new_config, path):
    """(train_CXR14.py:104-114) also reads checkpoints written by the reference itself (torch.optim.Adam state)."""
    checkpoint = torch.load(path, map_location=torch.device(new_config.device), weights_only=False)
    old_config = checkpoint["config"]
    compare_configs(old_config, new_config)
    model = DiffusionModel(old_config).to(new_config.device)
    model.load_state_dict(checkpoint["model_state_dict"])
    optimizer = FusedAdam(model.parameters(), lr=new_config.lr)
    optimizer.load_state_dict(checkpoint["optimizer_state_dict"])
    return model, optimizer, checkpoint["step"]


def build_image_dataloaders(config):
    rank, world = getattr(config, "rank", 0), getattr(config, "world_size", 1)
    data_dir = getattr(config, "data_dir", None)
    if getattr(config, "dataset", "CXR14") != "synthetic" and (data_dir is None or not os.path.isdir(str(data_dir))):
        raise FileNotFoundError(f"--data_dir {data_dir} does not exist; pass --dataset synthetic to train on generated images "
                                "(real data: <data_dir>/ holds the images, the split CSVs are read from --csv_dir or <repo>/data)")
    if getattr(config, "dataset", "CXR14") == "synthetic":
        from ..dataloaders.device_loader import build_synthetic_dataloaders
        return build_synthetic_dataloaders(config.img_size, config.batch_size, 0, labelled=False, device=config.device,
                                           rank=rank, world_size=world, n_train=max(256, 16 * config.batch_size))
    from ..dataloaders.CXR14 import build_dataloaders
    return build_dataloaders(config.data_dir, config.img_size, config.batch_size, config.num_workers,
                             device=config.device, rank=rank, world_size=world,
                             **({"csv_dir": config.csv_dir} if getattr(config, "csv_dir", None) else {}))


def main(config: Namespace) -> None:
    init_distributed(config)
    config.log_dir = Path(config.log_dir).parent / "CXR14" / Path(config.log_dir).name
    write_config(config)
    seed_everything(config.seed)                                  # identical initial replicas on every rank
    if config.resume_path is not None:
        print("Loading model from", config.resume_path)
        diffusion_model, optimizer, step = load(config, config.resume_path)
    else:
        diffusion_model = DiffusionModel(config).to(config.device)
        optimizer = FusedAdam(diffusion_model.parameters(), lr=config.lr)
        step = 0
    diffusion_model.train()
    if getattr(config, "world_size", 1) > 1:
        seed_everything(config.seed + config.rank)                # rank-distinct t / noise streams
    dataloaders = build_image_dataloaders(config)
    logger = TensorboardLogger(config.log_dir, enabled=not config.debug and getattr(config, "rank", 0) == 0)
    train(config, diffusion_model, optimizer, dataloaders["train"], dataloaders["val"], logger, None, step)
