"""Helpers the hot path shares with its callers (reference: trainers/utils.py:18-98)."""
from __future__ import annotations

import os
import random
from inspect import isfunction
from typing import Any, Optional, Tuple

import numpy as np
import torch
from torch import Tensor


def seed_everything(seed: int) -> None:  # trainers/utils.py:18-25
    random.seed(seed)
    os.environ["PYTHONHASHSEED"] = str(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed(seed)


def normalize_to_neg_one_to_one(img: Tensor) -> Tensor:  # trainers/utils.py:28-29
    return img * 2 - 1


def unnormalize_to_zero_to_one(img: Tensor) -> Tensor:  # trainers/utils.py:32-33
    return (img + 1) * 0.5


def exists(x: Any) -> bool:
    return x is not None


def default(val: Any, d: Any) -> Any:
    if exists(val):
        return val
    return d() if isfunction(d) else d


def get_index_from_list(vals: Tensor, t: Tensor, x_shape: Tuple[int, ...]) -> Tensor:  # trainers/utils.py:48-59
    return vals.gather(-1, t).reshape(t.shape[0], *((1,) * (len(x_shape) - 1)))


@torch.no_grad()
def sample_images(diffusion_model, T: int, img_size: int, batch: int, channels: int = 1,
                  cond: Optional[Tensor] = None, n_snapshots: int = 8, generator: Optional[torch.Generator] = None,
                  use_graph: bool = True):
    """Ancestral sampling driver (reference: sample_plot_image, trainers/utils.py:62-98, without the
    torchvision grid assembly): x_T ~ N(0, I); for t = T-1 .. 0: x <- sample_timestep(x, t).
    Returns (final images in [0,1] space, list of snapshots taken every T/n_snapshots steps)."""
    device = next(diffusion_model.parameters()).device
    if use_graph and device.type == "cuda" and cond is None and T == diffusion_model.timesteps:
        return GraphedSampler(diffusion_model, batch, channels, img_size).sample(T, n_snapshots, generator)
    img = torch.randn((batch, channels, img_size, img_size), device=device, generator=generator)
    stepsize = max(1, int(T / n_snapshots))
    snaps = []
    for t in range(T - 1, -1, -1):
        img = diffusion_model.sample_timestep(img, t=t, cond=cond)
        if t % stepsize == 0:
            snaps.append(unnormalize_to_zero_to_one(img))
    return unnormalize_to_zero_to_one(img), snaps


@torch.no_grad()
def sample_plot_image(diffusion_model, T: int, img_size: int, batch: int, channels: Optional[int] = 1,
                      cond: Optional[Tensor] = None, **kwargs) -> Tensor:
    """(trainers/utils.py:62-98) grids of 8 snapshots per sampled image, (batch, C, H, W) on the host.  The reference's
    own caller passes a `normalized=` keyword the function does not take (train_CXR14.py:84); extra keywords are
    accepted and ignored here.  The 8 D2H copies happen once at the end instead of inside the loop."""
    _, snaps = sample_images(diffusion_model, T, img_size, batch, channels, cond, n_snapshots=8)
    imgs = torch.stack([s.cpu() for s in snaps]).transpose(0, 1)             # n b c h w -> b n c h w
    grids = torch.stack([snapshot_grid(img_row) for img_row in imgs])        # (b, 3, H, W)
    return grids


def snapshot_grid(images: Tensor, nrow: int = 4, padding: int = 2) -> Tensor:
    """torchvision.utils.make_grid(images, nrow=nrow) for a (n, c, h, w) stack: single-channel images are repeated to three
    channels, tiles sit on a zero canvas with `padding` pixels between and around them."""
    n, c, h, w = images.shape
    if c == 1:
        images = images.expand(n, 3, h, w)
    ncol, nrows = min(nrow, n), (n + nrow - 1) // nrow
    grid = torch.zeros(images.shape[1], nrows * (h + padding) + padding, ncol * (w + padding) + padding, dtype=images.dtype)
    for k in range(n):
        r, cc = divmod(k, ncol)
        grid[:, r * (h + padding) + padding:r * (h + padding) + padding + h,
             cc * (w + padding) + padding:cc * (w + padding) + padding + w] = images[k]
    return grid


class TensorboardLogger:
    """(trainers/utils.py:101-147) scalar / image logging; a no-op when disabled or when tensorboard is absent."""

    def __init__(self, log_dir=None, config=None, enabled: bool = True, **kwargs):
        self.enabled = enabled
        self.writer = None
        if enabled:
            try:
                from torch.utils.tensorboard import SummaryWriter
                self.writer = SummaryWriter(log_dir=str(log_dir), **kwargs)
            except Exception as e:  # tensorboard is an optional dependency
                print(f"tensorboard unavailable ({e}); logging to stdout only")
                self.enabled = False

    def log(self, data, step: int) -> None:
        if not self.enabled:
            return
        for k, v in data.items():
            if isinstance(v, (int, float)):
                self.writer.add_scalar(k, v, step)
            elif isinstance(v, (np.ndarray, torch.Tensor)) and len(v.shape) == 3:
                self.writer.add_image(k, v, step)
            elif isinstance(v, (np.ndarray, torch.Tensor)) and len(v.shape) == 4:
                self.writer.add_images(k, v, step)
            else:
                raise ValueError(f"Unsupported data type: {type(v)}")


def compare_configs(config_old, config_new) -> None:  # trainers/utils.py:150-169
    c_old, c_new = vars(config_old), vars(config_new)
    for k, v in c_old.items():
        if k in c_new and c_new[k] != v:
            print(f"{k} differs - old: {v} new: {c_new[k]}")
    for k, v in c_new.items():
        if k not in c_old:
            print(f"{k} is new - {v}")
    for k, v in c_old.items():
        if k not in c_new:
            print(f"{k} is removed - {v}")


# -- data-parallel launcher glue (one process per GPU under torchrun; a single process otherwise) -------------------
def init_distributed(config) -> Tuple[int, int]:
    """Join the torchrun rendezvous when there is one; pin this process to its GPU.  -> (rank, world_size)."""
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if str(getattr(config, "device", "cuda")).startswith("cuda") and torch.cuda.is_available():
        torch.cuda.set_device(local)
        config.device = f"cuda:{local}"
    if world > 1 and not dist.is_initialized():
        backend = "nccl" if str(config.device).startswith("cuda") else "gloo"
        dist.init_process_group(backend=backend)
    config.rank, config.world_size = rank, world
    return rank, world


def dp_optimizer_step(optimizer, world_size: int = 1) -> None:
    """optimizer.step() of the reference loops; with world_size > 1 the gradients are summed over ranks first
    (ONE all-reduce when the optimiser keeps a flat gradient arena) and the mean is taken inside Adam."""
    import torch.distributed as dist
    from ..optim import FusedAdam
    if world_size > 1:
        if isinstance(optimizer, FusedAdam):
            flat = optimizer.flat_grad()
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)
            optimizer.step(grad_scale=1.0 / world_size, flat_grad=flat)
            return
        from ..parallel import allreduce_gradients
        allreduce_gradients([p for g in optimizer.param_groups for p in g["params"]])
    optimizer.step()


class GraphedSampler:
    """The reverse-diffusion loop of `sample_plot_image` with ONE reverse step (UNet forward + posterior update, ~600
    launches) captured in a CUDA graph.  Small sampling batches -- the 8 images the reference samples at every
    validation -- are launch-bound from Python (5-6 ms of issue time per step against 1-2 ms of GPU time); a replayed
    step costs three tiny launches (timestep fill, schedule-row copy, noise draw) and one graph launch.
    The graph is tied to the current parameter storage: build a new sampler after the weights changed."""

    def __init__(self, diffusion_model, batch: int, channels: int, img_size: int):
        dm = diffusion_model
        if dm.objective != "pred_noise":
            raise ValueError("only objective='pred_noise' can sample (as in the reference)")
        self.dm = dm
        dev = next(dm.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("GraphedSampler runs on CUDA (sm_100a) only; there is no CPU fallback")
        self.x = torch.zeros(batch, channels, img_size, img_size, device=dev)
        self.z = torch.zeros_like(self.x)
        self.tt = torch.zeros(batch, device=dev, dtype=torch.long)
        self.coefs = torch.zeros(5, device=dev)
        self.table = dm.reverse_tables(dev)
        chw = self.x[0].numel()
        rank = torch.tensor(dm.dynamic_threshold_percentile, dtype=torch.float32) * (chw - 1)
        self.k_lo = int(torch.floor(rank).item())
        self.q_weight = float((rank - torch.floor(rank)).item())
        from .. import native as N
        self._N = N
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():          # warm-up off the capture (allocator growth, weight re-layout)
            self.coefs.copy_(self.table[dm.timesteps - 1])
            for _ in range(2):
                self._step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.x.zero_()
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self._step()

    def _step(self) -> None:
        eps = self.dm.model(self.x, self.tt)
        out = self._N.sampler_step_dev(self.x, eps.contiguous(), self.z, self.coefs, self.k_lo, self.q_weight)
        self.x.copy_(out)

    @torch.no_grad()
    def step(self, t: int, noise: Optional[Tensor] = None) -> Tensor:
        """x_t (held in `self.x`) -> x_{t-1} in place; returns `self.x`."""
        self.tt.fill_(t)
        self.coefs.copy_(self.table[t])
        if noise is None:
            self.z.normal_()
        else:
            self.z.copy_(noise)
        self.graph.replay()
        return self.x

    @torch.no_grad()
    def sample(self, T: int, n_snapshots: int = 8, generator: Optional[torch.Generator] = None):
        self.x.normal_(generator=generator)
        stepsize = max(1, int(T / n_snapshots))
        snaps = []
        for t in range(T - 1, -1, -1):
            self.step(t)
            if t % stepsize == 0:
                snaps.append(unnormalize_to_zero_to_one(self.x.clone()))
        return unnormalize_to_zero_to_one(self.x.clone()), snaps
