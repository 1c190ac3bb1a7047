"""DDPM pre-training step as the reference runs it (trainers/train_CXR14.py:28-40:
`loss = model.train_step(x); loss.backward(); optimizer.step()`), replayed from CUDA graphs.

One step is ~450 kernel launches (q_sample, UNet forward, loss, UNet backward, Adam); at the
reference's batch size of 16 the GPU finishes them faster than Python can issue them, so the
launch-bound sequence is captured once and replayed:

    graph A: zero_grad -> weight re-layout -> q_sample -> UNet fwd -> L1/p2 loss -> UNet bwd   (gradient arena)
             [world > 1] the gradient all-reduce is INSIDE this graph, on a communication stream, in two regions: the
             decoder / mid / output half of the arena is summed over the ranks while the encoder's backward still runs,
             the encoder half (+ the time projections) at the end (sum; the 1/world goes into Adam's grad_scale)
    graph B: fused Adam over the flat parameter / moment arenas -> version bump
If NCCL cannot be captured on this build, the step falls back to ONE all-reduce of the whole arena between the two graphs.

Shapes are static (batch, image size); timesteps and noise are drawn inside the graph by torch's
graph-safe Philox generator, exactly where the reference draws them (diffusion_model.py:126-129,193).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist
from torch import Tensor

from .models.diffusion_model import DiffusionModel
from .optim import FusedAdam


class GradReducer:
    """Sums regions of the gradient arena over the ranks on its own stream while the backward pass is still producing the
    rest (hooked into tedm_b200.engine.UnetEngine.backward).  Bucket sizes follow the backward's structure, not a link
    count: NVSwitch gives every pair of GPUs full bandwidth, so two large collectives cost two launch latencies."""

    def __init__(self, device):
        self.comm = torch.cuda.Stream(device=device)
        self.collectives = 0

    def early(self, flat_slice: Tensor, side: Optional[torch.cuda.Stream]) -> None:
        self.comm.wait_stream(torch.cuda.current_stream())
        if side is not None:
            self.comm.wait_stream(side)                       # weight gradients are written on the engine's side stream
        with torch.cuda.stream(self.comm):
            dist.all_reduce(flat_slice, op=dist.ReduceOp.SUM)
        self.collectives += 1

    def late(self, tensors) -> None:
        cur = torch.cuda.current_stream()
        self.comm.wait_stream(cur)
        with torch.cuda.stream(self.comm):
            for t in tensors:
                dist.all_reduce(t, op=dist.ReduceOp.SUM)
                self.collectives += 1
        cur.wait_stream(self.comm)                            # join: Adam (and the time-projection scatter) see reduced gradients


class GraphedTrainStep:
    def __init__(self, model: DiffusionModel, optimizer: FusedAdam, example: Tensor, warmup: int = 3,
                 use_graph: bool = True):
        if not example.is_cuda:
            raise RuntimeError("GraphedTrainStep runs on CUDA (sm_100a) only; there is no CPU fallback")
        self.model, self.opt = model, optimizer
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.x = example.detach().float().clone()
        self.loss = torch.zeros((), device=example.device)
        self.use_graph = use_graph
        self._ga: Optional[torch.cuda.CUDAGraph] = None
        self._gb: Optional[torch.cuda.CUDAGraph] = None
        self._params = [p for p in model.parameters() if p.requires_grad]
        self._flat: Optional[Tensor] = None      # the gradient arena graph A writes and graph B reads (fixed at capture)
        self._hyper = None                       # optimiser hyper-parameters frozen into graph B
        # overlapped, in-graph gradient reduction (world > 1); TEDM_DP_OVERLAP=0 keeps the single all-reduce between the graphs
        import os
        self.reducer: Optional[GradReducer] = None
        if self.world > 1 and os.environ.get("TEDM_DP_OVERLAP", "1") != "0":
            self.reducer = GradReducer(example.device)
        if use_graph:
            try:
                self._capture(warmup)
            except Exception as e:                                   # NCCL not capturable here: fall back, loudly
                if self.reducer is None:
                    raise
                print(f"GraphedTrainStep: in-graph gradient reduction unavailable ({type(e).__name__}: {str(e)[:120]}); "
                      "using one all-reduce between the graphs")
                torch.cuda.synchronize()
                self.reducer = None
                self._ga = self._gb = None
                self._capture(warmup)

    # -- the two halves of a step ---------------------------------------------------------------
    def _fwd_bwd(self) -> None:
        self.opt.zero_grad(set_to_none=True)
        loss = self.model.train_step(self.x)
        eng = self.model.model.engine
        eng.grad_reducer = self.reducer
        try:
            loss.backward()
        finally:
            eng.grad_reducer = None
        self.loss.copy_(loss.detach())

    def _update(self, flat: Optional[Tensor] = None) -> None:
        self.opt.step(grad_scale=1.0 / self.world, flat_grad=flat)

    def _allreduce(self, flat: Optional[Tensor] = None) -> None:
        if self.world > 1 and self.reducer is None:              # with a reducer the sums happened inside _fwd_bwd
            dist.all_reduce(self.opt.flat_grad() if flat is None else flat, op=dist.ReduceOp.SUM)

    def _hyper_sig(self):
        g = self.opt.param_groups[0]
        return (float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]), float(g["weight_decay"]))

    def _capture_update(self) -> None:
        """Graph B: Adam over the arena graph A fills.  lr / betas / eps / weight decay are kernel arguments frozen at
        capture, so the graph is rebuilt whenever the param_group changes (an LR scheduler)."""
        self._gb = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._gb, pool=self._ga.pool()):
            self._update(self._flat)
        self.opt._step -= 1                              # the capture recorded a step, it did not execute one
        self._hyper = self._hyper_sig()

    def _capture(self, warmup: int) -> None:
        eng = self.model.model.engine
        opt = self.opt
        # the warm-up below runs real steps (lazy initialisation, allocator growth, NCCL channels); the optimiser state it
        # touches is restored afterwards so that constructing a GraphedTrainStep is not a training step: the reference does
        # exactly one update per loop iteration (trainers/train_CXR14.py:28-40)
        snap = (opt.flat_param.clone(), opt.exp_avg.clone(), opt.exp_avg_sq.clone(), opt.step_counter.clone(), opt._step)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):            # warm-up off the capture: lazy attribute setting, allocator growth
            for _ in range(max(1, warmup)):
                self._fwd_bwd()
                self._allreduce()
                self._update()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        with torch.no_grad():
            opt.flat_param.copy_(snap[0])
            opt.exp_avg.copy_(snap[1])
            opt.exp_avg_sq.copy_(snap[2])
            opt.step_counter.copy_(snap[3])
        opt._step = snap[4]
        opt.bump_versions()
        del snap
        eng.force_refresh = True
        eng.cache.force = True
        from . import native as N
        l0 = N.launches
        try:
            self._ga = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._ga):
                self._fwd_bwd()
            # the arena graph A writes lives at a fixed address in the graph's pool: the all-reduce and graph B must use
            # THAT tensor, not whatever p.grad points at later (an eager odd-sized step rebinds p.grad to a fresh arena)
            self._flat = opt.flat_grad()
            first = self._params[0].grad
            if first is None or self._flat.data_ptr() != first.data_ptr():
                raise RuntimeError("GraphedTrainStep: the backward did not produce one flat gradient arena")
            self._capture_update()
        finally:
            eng.force_refresh = False
            eng.cache.force = False
        self.native_calls_per_step = N.launches - l0     # native entry points inside one replayed step

    # -- public ---------------------------------------------------------------------------------
    def close(self) -> None:
        """Release the CUDA graphs.  With the in-graph gradient reduction they hold captured NCCL operations, and
        `dist.destroy_process_group()` waits for those: close the step (or drop it) BEFORE tearing the process group down."""
        torch.cuda.synchronize()
        self._ga = self._gb = None
        self._flat = None

    def __call__(self, x: Tensor) -> Tensor:
        """One optimiser step on batch x (same shape as the example); returns the loss (device scalar)."""
        if x.shape != self.x.shape:
            raise ValueError(f"GraphedTrainStep was captured for batch shape {tuple(self.x.shape)}, got {tuple(x.shape)}")
        self.x.copy_(x, non_blocking=True)
        if self._ga is None:
            self._fwd_bwd()
            self._allreduce()
            self._update()
        else:
            if self._hyper != self._hyper_sig():
                self._capture_update()
            self._ga.replay()
            self._allreduce(self._flat)
            self._gb.replay()
            self.opt._step += 1
            self.opt.bump_versions()
        return self.loss
