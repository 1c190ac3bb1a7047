"""Training-mode forward of the LEDM/TEDM head (BatchNorm batch statistics, gradients into the
head parameters only; the UNet is frozen: datasetDM_model.py:67 is @torch.no_grad).

STATUS: interim.  Feature extraction (S UNet forwards per image) runs on the native sm_100a path;
the tiny head itself (0.13-0.99 M parameters) is evaluated here with torch autograd in the commuted
form -- layer 1 applied per level at native resolution, then nearest-upsampled and summed -- until
the three-phase native head-training kernels replace it.
"""
from __future__ import annotations

from typing import Sequence

import torch
import torch.nn.functional as F
from torch import Tensor, nn

from . import native as N


def head_train_forward(model, x: Tensor, convs: Sequence[nn.Conv2d], bns: Sequence[nn.BatchNorm2d]) -> Tensor:
    feats, b, s = model.feature_maps(x)
    chans = [f.shape[-1] for f in feats]
    ctot = sum(chans)
    shared = convs[0].in_channels == ctot
    size = x.shape[-1]
    w1 = convs[0].weight
    z1 = None
    for l, f in enumerate(feats):
        fl = N.nhwc_to_nchw_f32(f)                                   # (B*S, C_l, h, w) fp32, no grad
        off = sum(chans[:l])
        if shared:
            g = F.conv2d(fl, w1[:, off:off + chans[l]])
        else:
            fl = fl.reshape(b, s * chans[l], *fl.shape[2:])
            wl = torch.cat([w1[:, st * ctot + off: st * ctot + off + chans[l]] for st in range(s)], dim=1)
            g = F.conv2d(fl, wl)
        g = F.interpolate(g, size=[size, size])
        z1 = g if z1 is None else z1 + g
    h = bns[0](F.relu(z1 + convs[0].bias[None, :, None, None]))
    h = bns[1](F.relu(convs[1](h)))
    return convs[2](h)
