"""Two more fixtures from the LIVE reference (/root/reference, CPU), run in the build container only:

    python tests/golden/make_golden_floor.py

1. `noise_floor.json` -- the reference against ITSELF at reduced precision: the same modules, weights, inputs and
   noise as `tedm_full.npz` / `tedm_full_trained.npz` / `tedm_b16_trained.npz`, run under
   `torch.autocast('cpu', dtype=torch.bfloat16)` (what `--mixed_precision` turns on, trainers/train_CXR14.py:29),
   compared with the reference's fp32 outputs: relative logit error and mask agreement.  The GPU tests assert
   that the CUDA bf16 path is at least as close to reference-fp32 as reference-bf16 is.
2. `tedm_b16_trained.npz` -- the bench configuration (BASELINE configs[3] at config.py defaults: B = 16 images,
   S = 8 timesteps, 128 x 128 -> 262 144 mask pixels) with a head the reference trained: packed reference masks and
   fp16 ensemble probabilities for all 16 images, fp32 logits for the first 4.
Inputs are regenerated from tests/golden/synth.py by seed, so they are not stored.
"""
import json
import os
import sys
from argparse import Namespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, "/root/reference")

from tests.golden.make_golden import FixedNoise, load_synth, np_, tedm_head  # noqa: E402
from tests.golden.synth import synth_images, synth_noise  # noqa: E402

from models.datasetDM_model import DatasetDM  # noqa: E402  (reference)

torch.set_grad_enabled(False)
torch.set_num_threads(os.cpu_count())

STEPS = [1, 10, 25, 50, 200, 400, 600, 800]
SIZE = 128
B16 = {"batch": 16, "image_seed": 40, "noise_seed0": 60, "train_images": 16, "train_iters": 200}
SKIP = ("diffusion_model.sqrt_", "diffusion_model.posterior_", "diffusion_model.p2_")


def build(head_sd=None):
    cfg = Namespace(normalize=True, saved_diffusion_model="/nonexistent", verbose=False, device="cpu", t_steps_to_save=STEPS)
    ted = DatasetDM(cfg)
    ted.classifier = tedm_head(len(STEPS))
    ted.eval()
    load_synth(ted, 0, skip=SKIP)
    if head_sd is not None:
        ted.classifier.load_state_dict(head_sd)
    return ted


def ensemble(logits, b):
    pr = torch.sigmoid(logits.float()).reshape(b, len(STEPS), 1, SIZE, SIZE).mean(1)
    return pr, pr > 0.5


def floor_entry(ted, x0, noises, ref_logits, ref_mask):
    """reference under autocast(bf16) against reference fp32 on identical inputs"""
    with torch.autocast("cpu", dtype=torch.bfloat16), FixedNoise(noises):
        lb = ted(x0).float()
    _, mb = ensemble(lb, x0.shape[0])
    ref_logits = torch.as_tensor(ref_logits)
    rel = ((lb.double() - ref_logits.double()).norm() / ref_logits.double().norm()).item()
    agree = float((mb.numpy() == np.asarray(ref_mask)).mean())
    return {"logits_rel": rel, "mask_agree": agree, "pixels": int(np.asarray(ref_mask).size)}


def main():
    floor = {"how": "reference modules under torch.autocast('cpu', bfloat16) vs the same modules in fp32, same weights/inputs/noise",
             "torch": torch.__version__}
    # ---- the two existing full-size fixtures -------------------------------------------------------------
    x0 = synth_images(1, SIZE, 3)
    noises = [synth_noise((1, 1, SIZE, SIZE), 20 + i, "tedm") for i in range(len(STEPS))]
    g = np.load(os.path.join(HERE, "tedm_full.npz"))
    floor["tedm_full"] = floor_entry(build(), x0, noises, g["tedm_logits"], g["tedm_mask"])
    print("tedm_full", floor["tedm_full"], flush=True)
    g = np.load(os.path.join(HERE, "tedm_full_trained.npz"))
    head = {k[len("classifier."):]: torch.from_numpy(g[k]) for k in g.files if k.startswith("classifier.")}
    floor["tedm_full_trained"] = floor_entry(build(head), x0, noises, g["tedm_logits"], g["tedm_mask"])
    print("tedm_full_trained", floor["tedm_full_trained"], flush=True)

    # ---- bench configuration: 16 images x 8 steps, reference-trained head -----------------------------
    b = B16["batch"]
    x0 = synth_images(b, SIZE, B16["image_seed"])
    noises = [synth_noise((b, 1, SIZE, SIZE), B16["noise_seed0"] + i, "tedm") for i in range(len(STEPS))]
    ted = build()
    nt = B16["train_images"]
    with FixedNoise([n[:nt] for n in noises]):
        feats = ted.extract_features(x0[:nt])
    y = (x0[:nt] > 0.45).float().repeat_interleave(len(STEPS), dim=0)
    torch.manual_seed(0)
    with torch.enable_grad():
        opt = torch.optim.Adam(ted.classifier.parameters(), lr=2e-3)
        ted.classifier.train()
        for it in range(B16["train_iters"]):
            opt.zero_grad()
            out = ted.classifier(feats)
            loss = torch.nn.functional.binary_cross_entropy_with_logits(out, y, reduction="none").mean(dim=(2, 3)).mean()
            loss.backward()
            opt.step()
    del feats
    ted.classifier.eval()
    with FixedNoise(noises):
        logits = ted(x0)                                   # the reference's own call at B = 16
    pr, mask = ensemble(logits, b)
    target = (x0 > 0.45)
    print("b16 trained head: loss", float(loss), "frac |p-.5|<.02:", float(((pr - .5).abs() < .02).float().mean()),
          "<.005:", float(((pr - .5).abs() < .005).float().mean()),
          "acc vs target", float((mask == target).float().mean()), flush=True)
    out = {"steps": np.array(STEPS), "batch": np.array(b), "image_seed": np.array(B16["image_seed"]),
           "noise_seed0": np.array(B16["noise_seed0"]), "final_train_loss": np_(loss.detach()),
           "tedm_mask_packed": np.packbits(np_(mask).reshape(-1)), "tedm_prob_f16": np_(pr).astype(np.float16),
           "tedm_logits_first4": np_(logits[:4 * len(STEPS)])}
    for k, v in ted.classifier.state_dict().items():
        out[f"classifier.{k}"] = np_(v)
    np.savez_compressed(os.path.join(HERE, "tedm_b16_trained.npz"), **out)
    floor["tedm_b16_trained"] = floor_entry(ted, x0, noises, logits, np_(mask))
    print("tedm_b16_trained", floor["tedm_b16_trained"], flush=True)
    json.dump(floor, open(os.path.join(HERE, "noise_floor.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
