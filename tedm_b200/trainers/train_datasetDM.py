"""LEDM / LEDMe / TEDM head training (reference: trainers/train_datasetDM.py:13-63): a frozen DDPM UNet, a per-pixel
MLP head trained with BCE on a handful of labelled images.  Only `model.classifier.parameters()` are optimised."""
from __future__ import annotations

from argparse import Namespace

from ..models.datasetDM_model import DatasetDM, tedm_classifier
from ..optim import FusedAdam
from .train_baseline import build_segmentation_dataloaders, train, write_config
from .utils import TensorboardLogger, init_distributed, seed_everything


def build_model(config: Namespace) -> DatasetDM:
    model = DatasetDM(config)
    if getattr(config, "shared_weights_over_timesteps", False):       # TEDM: one 960-input head shared by all timesteps
        model.classifier = tedm_classifier(len(model.steps), getattr(config, "out_channels", 1))
    return model


def main(config: Namespace) -> None:
    init_distributed(config)
    write_config(config)
    print("Experiment folder: %s" % (config.log_dir))
    seed_everything(config.seed)
    model = build_model(config).to(config.device)
    model.sync_bn = bool(getattr(config, "sync_bn", False))          # tedm_b200/head_train.py
    model.train()
    model.diffusion_model.eval()
    optimizer = FusedAdam(model.classifier.parameters(), lr=config.lr, weight_decay=config.weight_decay)
    dataloaders = build_segmentation_dataloaders(config)
    logger = TensorboardLogger(config.log_dir, enabled=not config.debug and getattr(config, "rank", 0) == 0)
    train(config, model, optimizer, dataloaders["train"], dataloaders["val"], logger, None, 0)
