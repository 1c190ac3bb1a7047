"""Per-timestep and ensembled test metrics of a TEDM model (reference: auxiliary/postprocessing/
testing_shared_weights.py:104-144), computed on the device.

The reference moves every sigmoid map to the host, concatenates the whole test set, rearranges it to
'step b 1 h w' and calls dice / precision / recall once per timestep and once for the step-mean.  Here each batch
goes logits -> (a) per-(image, step) metrics straight from the logits, (b) sigmoid-mean-threshold ensemble mask ->
metrics, all in three kernel launches; only the (rows x 8) metric table ever reaches the host."""
from __future__ import annotations

from typing import Dict, Iterable, Tuple

import torch
from torch import Tensor

from . import native as N


def _stats(col: Tensor) -> Tuple[float, float]:
    return col.mean().item(), (col.std().item() if col.numel() > 1 else 0.0)


@torch.no_grad()
def evaluate_shared_weights(model, loader: Iterable, device="cuda") -> Dict:
    """-> {"per_timestep": {t: {"dice": (mean, std), "precision": .., "recall": ..}}, "ensemble": {...}, "n_images": n}."""
    was_training = model.training
    model.eval()
    steps = list(model.steps)
    s = len(steps)
    per_step, ens = [], []
    for x, y in loader:
        x, y = x.to(device), y.to(device).float().contiguous()
        logits = model(x).float().contiguous()                    # (B*S, 1, H, W), row b*S + step
        if logits.shape[0] != x.shape[0] * s:
            raise RuntimeError("evaluate_shared_weights needs the shared-weight (TEDM) head: one logit map per timestep")
        per_step.append(N.seg_metrics(logits, y).reshape(x.shape[0], s, 8))    # sigmoid(logit) > .5 per (image, step)
        mask, _ = N.ensemble_mask(logits, s)                      # mean_step sigmoid > .5
        ens.append(N.seg_metrics(mask, y).reshape(x.shape[0], 8))
    per_step, ens = torch.cat(per_step), torch.cat(ens)
    names = ("dice", "precision", "recall")
    out = {"per_timestep": {t: {k: _stats(per_step[:, i, j]) for j, k in enumerate(names)} for i, t in enumerate(steps)},
           "ensemble": {k: _stats(ens[:, j]) for j, k in enumerate(names)}, "n_images": int(ens.shape[0])}
    model.train(was_training)
    return out


def print_report(name: str, res: Dict) -> None:
    for t, m in res["per_timestep"].items():
        print(f"{name} {t} metrics: \n\tdice:      {m['dice'][0]:.3}+/-{m['dice'][1]:.3}")
        print(f"\tprecision: {m['precision'][0]:.3}+/-{m['precision'][1]:.3}")
        print(f"\trecall:    {m['recall'][0]:.3}+/-{m['recall'][1]:.3}")
    m = res["ensemble"]
    print(f"{name} metrics: \n\tdice:      {m['dice'][0]:.3}+/-{m['dice'][1]:.3}")
    print(f"\tprecision: {m['precision'][0]:.3}+/-{m['precision'][1]:.3}")
    print(f"\trecall:    {m['recall'][0]:.3}+/-{m['recall'][1]:.3}")
