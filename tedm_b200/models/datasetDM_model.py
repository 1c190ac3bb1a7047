"""LEDM / LEDMe / TEDM segmentation model behind the reference's `DatasetDM` surface
(models/datasetDM_model.py:30-88; the TEDM shared head of trainers/train_datasetDM.py:30-42).

Differences in execution, not in results:
  * all B x S (image, timestep) pairs run as ONE UNet batch (index b*S + s, the reference's '(b step)'
    order) and the part of the UNet after the last hooked map is skipped;
  * the hooked decoder maps stay on the device in NHWC bf16 (the reference copies them to the host);
  * the 960(*S)-channel upsampled feature tensor is never built for `forward`: layer 1 of the head is
    applied per level at native resolution on the tcgen05 GEMM, and one fused kernel does
    gather-upsample-sum + ReLU/BN + 128->32 + ReLU/BN + 32->1 (eval-mode BatchNorm).
`extract_features` still returns the reference's (B, 960*S, H, W) fp32 tensor for API compatibility.
"""
from __future__ import annotations

import os
from argparse import Namespace
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F
from torch import Tensor, nn

from .. import native as N
from ..engine import WeightCache
from .diffusion_model import DiffusionModel


class StepsToBatch(nn.Module):
    """'b (step act) h w -> (b step) act h w' (the einops Rearrange of train_datasetDM.py:34, parameter-free)."""

    def __init__(self, step: int):
        super().__init__()
        self.step = step

    def forward(self, x: Tensor) -> Tensor:
        b, c, h, w = x.shape
        return x.reshape(b * self.step, c // self.step, h, w)


def tedm_classifier(n_steps: int, out_channels: int = 1) -> nn.Sequential:
    """The shared-weight TEDM head with the reference's state_dict indices (classifier.1/3/4/6/7)."""
    return nn.Sequential(StepsToBatch(n_steps), nn.Conv2d(960, 128, 1), nn.ReLU(), nn.BatchNorm2d(128),
                         nn.Conv2d(128, 32, 1), nn.ReLU(), nn.BatchNorm2d(32), nn.Conv2d(32, 1, out_channels))


class DatasetDM(nn.Module):
    def __init__(self, args: Namespace) -> None:
        super().__init__()
        path = getattr(args, "saved_diffusion_model", None)
        if not path or not os.path.isfile(path):
            self.diffusion_model = DiffusionModel(args)
            if getattr(args, "verbose", False):
                print(f"No model found at {path}. Please load model!")
        else:
            ckpt = torch.load(path, map_location=torch.device(getattr(args, "device", "cpu")), weights_only=False)
            self.diffusion_model = DiffusionModel(ckpt["config"])
            self.diffusion_model.load_state_dict(ckpt["model_state_dict"])
        self.diffusion_model.eval()
        self._features = {}
        self.steps: List[int] = list(args.t_steps_to_save)
        self.classifier = nn.Sequential(
            nn.Conv2d(960 * len(self.steps), 128, 1), nn.ReLU(), nn.BatchNorm2d(128),
            nn.Conv2d(128, 32, 1), nn.ReLU(), nn.BatchNorm2d(32), nn.Conv2d(32, 1, 1))
        self._cache = WeightCache()

    def set_precision(self, precision: str) -> "DatasetDM":
        """'bf16' (default) | 'fp32': the UNet features AND the head then run at the reference's fp32 accuracy
        (tedm_b200/engine_fp32.py); inference only."""
        self.diffusion_model.set_precision(precision)
        return self

    @property
    def precision(self) -> str:
        return self.diffusion_model.model.precision

    # -- feature extraction -----------------------------------------------------------------------
    @torch.no_grad()
    def feature_maps(self, x_0: Tensor, noise: Optional[Tensor] = None) -> Tuple[List[Tensor], int, int]:
        """Native-resolution decoder maps for every (image, step): list over levels of NHWC bf16
        (B*S, h_l, w_l, C_l) tensors, batch index = b*S + s.  (datasetDM_model.py:67-79)"""
        if noise is not None:
            assert x_0.shape == noise.shape
        if not x_0.is_cuda:
            raise RuntimeError("tedm_b200.DatasetDM runs on CUDA (sm_100a) only; there is no CPU fallback")
        dm = self.diffusion_model
        b, s = x_0.shape[0], len(self.steps)
        x_rep = x_0.detach().float().repeat_interleave(s, dim=0).contiguous()
        key = (x_0.device, b, tuple(self.steps))
        if getattr(self, "_t_key", None) != key:         # device copy of the timestep vector, built once per (device, B)
            self._t_key, self._t_dev = key, torch.tensor(self.steps, device=x_0.device, dtype=torch.long).repeat(b)
        t = self._t_dev
        if noise is None:
            nz = torch.randn_like(x_rep)               # a fresh draw per step, as randn_like in the loop does
        else:
            nz = noise.detach().float().repeat_interleave(s, dim=0).contiguous()
        # NB: x_0 is NOT rescaled to [-1, 1] here (datasetDM_model.py:76)
        x_t = N.q_sample(x_rep, nz, t, dm.sqrt_alphas_cumprod, dm.sqrt_one_minus_alphas_cumprod, normalize=False)
        _, feats = dm.model.forward_features(x_t, t, skip_tail=True, time_key=(b, tuple(self.steps)))
        for i, f in enumerate(feats):
            self._features[i] = f
        return feats, b, s

    @torch.no_grad()
    def extract_features(self, x_0: Tensor, noise: Optional[Tensor] = None) -> Tensor:
        """Reference-format output (B, 960*S, H, W) fp32, channel order [step0: l0..l3, step1: ...]."""
        feats, b, s = self.feature_maps(x_0, noise)
        size = x_0.shape[-1]
        nchw = lambda f: f.permute(0, 3, 1, 2).contiguous() if f.dtype == torch.float32 else N.nhwc_to_nchw_f32(f)
        ups = [F.interpolate(nchw(f), size=[size, size]) for f in feats]                        # (B*S, C_l, H, W)
        per_step = torch.cat(ups, dim=1)                                                        # (B*S, 960, H, W)
        return per_step.reshape(b, s * per_step.shape[1], size, size)

    def _segment_graphed(self, x: Tensor, noise: Optional[Tensor]):
        if not x.is_cuda:
            raise RuntimeError("tedm_b200.DatasetDM runs on CUDA (sm_100a) only; there is no CPU fallback")
        sig = (tuple(x.shape), noise is not None, self.precision,
               tuple((t.data_ptr(), t._version) for t in list(self.parameters()) + list(self.buffers())))
        g = getattr(self, "_seg_graph", None)
        if g is None or g["sig"] != sig:
            sx = x.detach().float().clone()
            sn = noise.detach().float().clone() if noise is not None else None
            side = torch.cuda.Stream(device=x.device)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):                       # warm-up off the capture: caches, allocator growth
                for _ in range(2):
                    self.segment(sx, sn)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            l0, f0 = N.launches, N.conv_flops
            with torch.cuda.graph(graph):
                out = self.segment(sx, sn)
            g = self._seg_graph = {"sig": sig, "graph": graph, "x": sx, "noise": sn, "out": out, "calls": N.launches - l0,
                                   "conv_flops": N.conv_flops - f0}
        g["x"].copy_(x, non_blocking=True)
        if noise is not None:
            g["noise"].copy_(noise, non_blocking=True)
        g["graph"].replay()
        N.launches += g["calls"]          # native launches inside the replayed graph (bench.py's gpu_launches claim)
        N.conv_flops += g["conv_flops"]   # ... and the conv FLOPs they execute
        return g["out"]

    # -- head -------------------------------------------------------------------------------------
    def _head_layers(self):
        convs = [m for m in self.classifier if isinstance(m, nn.Conv2d)]
        bns = [m for m in self.classifier if isinstance(m, nn.BatchNorm2d)]
        if len(convs) != 3 or len(bns) != 2 or any(c.kernel_size != (1, 1) for c in convs):
            raise RuntimeError("unrecognised classifier: expected Conv1x1-ReLU-BN-Conv1x1-ReLU-BN-Conv1x1")
        return convs, bns

    def forward(self, x: Tensor, noise: Optional[Tensor] = None) -> Tensor:
        """(datasetDM_model.py:85-88) logits: (B, 1, H, W) for LEDM/LEDMe, (B*S, 1, H, W) for the TEDM head.
        BatchNorm in training mode -> batch statistics + autograd into the head parameters (tedm_b200/head_train.py)."""
        convs, bns = self._head_layers()
        if bns[0].training or bns[1].training:
            if not (bns[0].training and bns[1].training):
                raise RuntimeError("both BatchNorm layers of the classifier must be in the same mode")
            from ..head_train import head_train_forward
            return head_train_forward(self, x, convs, bns, noise)
        return self._head_infer(x, convs, bns, noise)

    def _layer1_maps(self, feats, w1: Tensor, shared: bool, b: int, s: int, chans, offs, skip=()) -> List[Tensor]:
        """Layer 1 of the head applied per level at native resolution (it commutes with the nearest upsample):
        fp32 NHWC maps [(B*S), h_l, w_l, 128] on the tcgen05 1x1 conv."""
        ctot = sum(chans)
        g_maps = []
        for l, f in enumerate(feats):
            cl = chans[l]
            if l in skip:
                continue
            if f.dtype == torch.float32:
                # fp32 mode: the same per-level 1x1 conv with (hi, lo) operand pairs (hi*hi + lo*hi + hi*lo on tcgen05)
                from ..engine_fp32 import split_weight
                if shared:
                    wl = self._cache.get(f"w1.l{l}:w3", (w1,), lambda w, o=offs[l], c=cl: split_weight(w[:, o:o + c], N.MODE_1X1))
                    g_maps.append(N.f32_conv(N.f32_split(f), wl, N.MODE_1X1, w1.shape[0]))
                else:
                    g = torch.empty(b * s, f.shape[1], f.shape[2], w1.shape[0], device=f.device, dtype=torch.float32)
                    hi, lo = N.f32_split(f)
                    for st in range(s):
                        wl = self._cache.get(f"w1.s{st}.l{l}:w3", (w1,),
                                             lambda w, o=st * ctot + offs[l], c=cl: split_weight(w[:, o:o + c], N.MODE_1X1))
                        N.f32_conv((hi[st::s], lo[st::s]), wl, N.MODE_1X1, w1.shape[0], out=g[st::s])
                    g_maps.append(g)
                continue
            if shared:
                wl = self._cache.get(f"w1.l{l}", (w1,), lambda w, o=offs[l], c=cl: w[:, o:o + c, 0, 0].to(torch.bfloat16).contiguous())
                g_maps.append(N.conv_igemm(f, wl, N.MODE_1X1, w1.shape[0], out_dtype=torch.float32))
            else:
                g = torch.empty(b * s, f.shape[1], f.shape[2], w1.shape[0], device=f.device, dtype=torch.float32)
                for st in range(s):
                    wl = self._cache.get(f"w1.s{st}.l{l}", (w1,),
                                         lambda w, o=st * ctot + offs[l], c=cl: w[:, o:o + c, 0, 0].to(torch.bfloat16).contiguous())
                    N.conv_igemm(f[st::s], wl, N.MODE_1X1, w1.shape[0], out=g[st::s])
                g_maps.append(g)
        return g_maps

    @torch.no_grad()
    def _head_infer(self, x: Tensor, convs: Sequence[nn.Conv2d], bns: Sequence[nn.BatchNorm2d],
                    noise: Optional[Tensor] = None) -> Tensor:
        feats, b, s = self.feature_maps(x, noise)
        chans = [f.shape[-1] for f in feats]
        ctot = sum(chans)
        c_in = convs[0].in_channels
        shared = c_in == ctot
        if not shared and c_in != ctot * s:
            raise RuntimeError(f"classifier expects {c_in} input channels; features give {ctot} per step x {s} steps")
        if bns[0].training or bns[1].training:
            raise RuntimeError("BatchNorm in training mode goes through head_train_forward")
        size = x.shape[-1]
        shifts = [(size // f.shape[1]).bit_length() - 1 for f in feats]
        offs = [sum(chans[:l]) for l in range(len(chans))]
        # shared head: the full-resolution level's layer 1 (64 -> 128) runs inside the tail kernel, its fp32 map never exists
        exact = feats[0].dtype == torch.float32         # fp32 mode: every level's layer 1 as an fp32 map, pure-fp32 tail kernel
        fuse_full = (not exact and shared and shifts[-1] == 0 and chans[-1] == 64 and convs[0].out_channels == 128
                     and (b * s * size * size) % 16 == 0)
        last = len(feats) - 1
        g_maps = self._layer1_maps(feats, convs[0].weight, shared, b, s, chans, offs, skip=(last,) if fuse_full else ())
        f_full = w1_full = None
        if fuse_full:
            f_full = feats[last]
            w1_full = self._cache.get(f"w1.l{last}", (convs[0].weight,),
                                      lambda w, o=offs[last], c=chans[last]: w[:, o:o + c, 0, 0].to(torch.bfloat16).contiguous())
            shifts = shifts[:last]

        def fold(bn: nn.BatchNorm2d, tag: str):
            def mk(w, bias, rm, rv):
                a = w.float() / torch.sqrt(rv.float() + bn.eps)
                return torch.stack([a, bias.float() - rm.float() * a]).contiguous()
            return self._cache.get(tag, (bn.weight, bn.bias, bn.running_mean, bn.running_var), mk)

        ac1, ac2 = fold(bns[0], "bn1"), fold(bns[1], "bn2")
        f32 = lambda p: p.detach().float().contiguous()
        logits = N.head_infer(g_maps, shifts, 1 if shared else s, b * s if shared else b, size, size,
                              f32(convs[0].bias), ac1[0], ac1[1], f32(convs[1].weight).reshape(convs[1].out_channels, -1),
                              f32(convs[1].bias), ac2[0], ac2[1], f32(convs[2].weight).reshape(-1),
                              float(self._cache.get("b3", (convs[2].bias,), lambda bb: bb.float().cpu()).item()),
                              f_full=f_full, w1_full=w1_full, exact=exact)
        return logits

    @torch.no_grad()
    def segment(self, x: Tensor, noise: Optional[Tensor] = None, graph: bool = False):
        """TEDM/LEDM inference with the reference's ensemble semantics
        (auxiliary/postprocessing/testing_shared_weights.py:113,120,133-138; app.py:79):
        sigmoid -> mean over steps -> > 0.5.  Returns (mask bool (B,1,H,W), prob, logits).

        graph=True replays the whole call (~600 launches) from a CUDA graph captured for this input shape and the
        current weights: small serving batches are launch-bound from Python (B = 1: 3.4 -> 1.5 ms).  The returned
        tensors are then the graph's static outputs, overwritten by the next call."""
        if graph:
            return self._segment_graphed(x, noise)
        convs, bns = self._head_layers()
        logits = self._head_infer(x, convs, bns, noise)
        n_steps = logits.shape[0] // x.shape[0]
        mask, prob = N.ensemble_mask(logits, n_steps)
        return mask, prob, logits
