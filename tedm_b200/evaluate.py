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


def main(argv=None) -> None:
    """`python -m tedm_b200.evaluate --experiment <log dir>`: the reference's test script
    (auxiliary/postprocessing/testing_shared_weights.py:29-144) -- loads the checkpoint found in the experiment
    directory, rebuilds the TEDM model from its stored config and reports per-timestep and ensembled metrics on the JSRT
    validation / test splits (and on NIH / Montgomery when `--nih_dir` / `--montgomery_dir` are given), saving one
    `<set>_metrics.pt` per test set."""
    import argparse
    import os
    from pathlib import Path

    from .dataloaders.device_loader import DeviceLoader, build_synthetic_dataloaders
    from .models.datasetDM_model import DatasetDM, tedm_classifier

    ap = argparse.ArgumentParser()
    ap.add_argument("--experiment", "-e", type=str, required=True, help="Experiment path")
    ap.add_argument("--rerun", "-r", default=False, action="store_true", help="Run the test again")
    ap.add_argument("--nih_dir", type=str, default=None)
    ap.add_argument("--nih_csv", type=str, default="correspondence_with_chestXray8.csv")
    ap.add_argument("--montgomery_dir", type=str, default=None)
    ap.add_argument("--montgomery_csv", type=str, default="patient_data.csv")
    args = ap.parse_args(argv)
    if not os.path.isdir(args.experiment):
        raise ValueError("Experiment path is not a directory")
    files = os.listdir(args.experiment)
    torch_file = next((f for f in files if "model" in f), None)
    if torch_file is None:
        raise ValueError("No checkpoint file found in experiment directory")
    print(f"Loading experiment from {torch_file}")
    data = torch.load(Path(args.experiment) / torch_file, map_location="cuda", weights_only=False)
    config = data["config"]
    if not getattr(config, "shared_weights_over_timesteps", False):
        raise ValueError("the per-timestep report needs a shared-weight (TEDM) experiment")
    model = DatasetDM(config)
    model.classifier = tedm_classifier(len(model.steps), getattr(config, "out_channels", 1))
    model.load_state_dict(data["model_state_dict"])
    model = model.eval().to("cuda")

    sets = {}
    data_dir = getattr(config, "data_dir", None)
    if data_dir is not None and os.path.isdir(str(data_dir)):
        from .dataloaders.JSRT import build_dataloaders
        dls = build_dataloaders(data_dir, config.img_size, config.batch_size, config.num_workers)
        sets.update(JSRT_val=dls["val"], JSRT_test=dls["test"])
    else:
        print(f"data_dir {data_dir} not found: evaluating on the synthetic validation / test pairs")
        dls = build_synthetic_dataloaders(config.img_size, config.batch_size, labelled=True)
        sets.update(synthetic_val=dls["val"], synthetic_test=dls["test"])
    if args.nih_dir:
        from .dataloaders.NIH import NIHDataset
        sets["NIH"] = DeviceLoader(NIHDataset(args.nih_dir, args.nih_dir, args.nih_csv, config.img_size), config.batch_size,
                                   False, config.num_workers, labelled=True)
    if args.montgomery_dir:
        from .dataloaders.Montgomery import MonDataset
        sets["Montgomery"] = DeviceLoader(MonDataset(args.montgomery_dir, args.montgomery_dir, args.montgomery_csv,
                                                     config.img_size), config.batch_size, False, config.num_workers, labelled=True)
    for name, loader in sets.items():
        out_file = Path(args.experiment) / f"{name}_metrics.pt"
        if out_file.exists() and not args.rerun:
            print(f"{name} already tested")
            print_report(name, torch.load(out_file, weights_only=False))
            continue
        print(f"Testing {name} set")
        res = evaluate_shared_weights(model, loader)
        print_report(name, res)
        torch.save(res, out_file)


if __name__ == "__main__":
    main()
