"""Training-mode forward/backward of the LEDM / LEDMe / TEDM head on native kernels
(classifier of models/datasetDM_model.py:57-64 and trainers/train_datasetDM.py:30-42; training loop
trainers/train_datasetDM.py:88-99: `pred = model(x)`, BCE-with-logits, `loss.backward()`, Adam on
`model.classifier.parameters()`).

The UNet is frozen (datasetDM_model.py:67 is @torch.no_grad), so gradients stop at the head's
parameters.  Schedule of one forward (BatchNorm uses BATCH statistics and updates its running buffers):

    features (S UNet forwards, native)                                         tedm_b200.engine
    g_l   = W1_l f_l      per level at native resolution, fp32                 tcgen05 conv (1x1)
    a1    = relu(b1 + sum_l up(g_l)) -> bf16, channel sums                     tedm_head_train_z1
    BN1 statistics; W2' = W2 diag(A1), b2' = b2 + W2 C1                        tedm_bn_finalize / tedm_head_fold_w2
    z2    = W2' a1 + b2'                                                       tcgen05 conv (1x1, fp32 out)
    BN2 statistics                                                             tedm_head_z2_stats / tedm_bn_finalize
    logit = w3 . BN2(relu(z2)) + b3                                            tedm_head_train_tail(0)

and of the backward: two reduction/apply passes over z2 (BN2, ReLU, layer 3), dW2' and d h1 on the
tcgen05 wgrad / conv kernels, two passes over (d h1, a1) (BN1, ReLU) that also pool d z1 to every
level's resolution, and one tcgen05 weight-gradient GEMM per level (per step for the unshared head).
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch
from torch import Tensor, nn

from . import native as N


def _sync_world(model) -> int:
    """> 1 when the head's BatchNorm statistics are shared over the data-parallel ranks (`model.sync_bn = True`,
    `--sync_bn`): the per-channel sum / sum of squares of both norms (2 x (128 + 32) floats) and, in the backward, the
    two reductions each norm's input gradient needs are summed over ranks, so that N replicas with B / N images each
    compute what the single-device reference computes on the whole batch of B (models/datasetDM_model.py:60,63 are
    plain BatchNorm2d on one device).  Every rank must hold the same number of images per step (DistributedSampler
    pads the shards to equal length)."""
    import torch.distributed as dist
    if getattr(model, "sync_bn", False) and dist.is_available() and dist.is_initialized():
        return dist.get_world_size()
    return 1


def _allreduce(t: Tensor) -> Tensor:
    import torch.distributed as dist
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


class HeadTrainFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, x: Tensor, noise: Optional[Tensor], w1, b1, g1, bt1, w2, b2, g2, bt2, w3, b3) -> Tensor:
        convs, bns = model._head_layers()
        if any(bn.momentum is None or not bn.track_running_stats for bn in bns):
            raise RuntimeError("head training: BatchNorm2d with momentum and running statistics (the reference's defaults)")
        feats, b, s = model.feature_maps(x, noise)
        chans = [f.shape[-1] for f in feats]
        ctot = sum(chans)
        c_in = convs[0].in_channels
        shared = c_in == ctot
        if not shared and c_in != ctot * s:
            raise RuntimeError(f"classifier expects {c_in} input channels; features give {ctot} per step x {s} steps")
        size = x.shape[-1]
        shifts = [(size // f.shape[1]).bit_length() - 1 for f in feats]
        offs = [sum(chans[:l]) for l in range(len(chans))]
        g_maps = model._layer1_maps(feats, convs[0].weight, shared, b, s, chans, offs)
        n_img, n_sum = (b * s, 1) if shared else (b, s)
        world = _sync_world(model)
        count = n_img * size * size * world             # pixels behind every BatchNorm statistic (all ranks when synced)
        f32 = lambda p: p.detach().float().contiguous()
        a1, sums1 = N.head_train_z1(g_maps, shifts, n_sum, n_img, size, size, f32(b1))
        if world > 1:
            _allreduce(sums1)
        stats1 = N.bn_finalize(sums1, count, f32(g1), f32(bt1), bns[0].eps, bns[0].momentum, bns[0].running_mean,
                               bns[0].running_var)
        w2f, b2f, w2t = N.head_fold_w2(f32(w2).reshape(w2.shape[0], -1), f32(b2), stats1)
        z2 = N.conv_igemm(a1, w2f, N.MODE_1X1, 64, bias=b2f, out_dtype=torch.float32)
        sums2 = N.head_z2_stats(z2)
        if world > 1:
            _allreduce(sums2)
        stats2 = N.bn_finalize(sums2, count, f32(g2), f32(bt2), bns[1].eps, bns[1].momentum,
                               bns[1].running_mean, bns[1].running_var)
        logits = N.head_train_tail(0, z2, stats2, f32(w3).reshape(-1), b3=f32(b3))
        for bn in bns:
            bn.num_batches_tracked += 1
        ctx.saved = (feats, a1, z2, stats1, stats2, w2t, f32(w3).reshape(-1))
        ctx.meta = (shared, b, s, chans, offs, shifts, count, tuple(w1.shape), tuple(w2.shape), tuple(w3.shape), world)
        return logits

    @staticmethod
    def backward(ctx, dlogits: Tensor):
        feats, a1, z2, stats1, stats2, w2t, w3 = ctx.saved
        shared, b, s, chans, offs, shifts, count, w1_shape, w2_shape, w3_shape, world = ctx.meta
        dev = z2.device
        dl = dlogits.detach().float().contiguous()
        # S / T: this rank's reductions (the parameter gradients come from them and are summed over ranks by the gradient
        # all-reduce like every other gradient); Sg / Tg: the same summed over ranks when the statistics are synced -- the
        # means a BatchNorm input gradient subtracts are means over the whole batch
        S = torch.zeros(5, 32, device=dev, dtype=torch.float32)
        T = torch.zeros(2, 128, device=dev, dtype=torch.float32)
        N.head_train_tail(1, z2, stats2, w3, dlogit=dl, S=S, count=count)
        Sg = _allreduce(S.clone()) if world > 1 else S
        dz2 = N.head_train_tail(2, z2, stats2, w3, dlogit=dl, S=Sg, count=count)
        if world > 1:
            S[4].copy_(Sg[4])         # row 4 (sum of dz2) was zero when S was reduced: it holds this rank's pass only
        dw2f = N.conv_wgrad(a1, dz2, N.MODE_1X1)                          # (64, 1, 128) w.r.t. the folded weights
        dh1 = N.conv_igemm(dz2, w2t, N.MODE_1X1, 128, out_dtype=torch.float32)   # fp32: BatchNorm-1's backward cancels means
        N.head_bn1_bwd(0, dh1, a1, stats1, T)
        Tg = _allreduce(T.clone()) if world > 1 else T
        db1 = torch.zeros(128, device=dev, dtype=torch.float32)
        pooled = N.head_bn1_bwd(1, dh1, a1, stats1, Tg, db1, shifts, count)
        ctot = sum(chans)
        dw1 = torch.empty(w1_shape[0], w1_shape[1], device=dev, dtype=torch.float32)
        for l, f in enumerate(feats):
            if shared:
                dw1[:, offs[l]:offs[l] + chans[l]] = N.conv_wgrad(f, pooled[l], N.MODE_1X1)[:, 0, :]
            else:
                for st in range(s):
                    o = st * ctot + offs[l]
                    dw1[:, o:o + chans[l]] = N.conv_wgrad(f[st::s], pooled[l], N.MODE_1X1)[:, 0, :]
        dw2 = torch.zeros(32, 128, device=dev, dtype=torch.float32)
        small = torch.zeros(32 + 128 + 128 + 32 + 32 + 32 + 1, device=dev, dtype=torch.float32)
        db2, dg1, dbt1, dg2, dbt2, dw3, db3 = torch.split(small, [32, 128, 128, 32, 32, 32, 1])
        N.head_param_grads(dw2f.reshape(64, 128), stats1, S, T, dw2, db2, dg1, dbt1, dg2, dbt2, dw3, db3)
        return (None, None, None, dw1.reshape(w1_shape), db1, dg1, dbt1, dw2.reshape(w2_shape), db2, dg2, dbt2,
                dw3.reshape(w3_shape), db3)


def head_train_forward(model, x: Tensor, convs: Sequence[nn.Conv2d], bns: Sequence[nn.BatchNorm2d],
                       noise: Optional[Tensor] = None) -> Tensor:
    return HeadTrainFunction.apply(model, x, noise, convs[0].weight, convs[0].bias, bns[0].weight, bns[0].bias,
                                   convs[1].weight, convs[1].bias, bns[1].weight, bns[1].bias, convs[2].weight,
                                   convs[2].bias)
