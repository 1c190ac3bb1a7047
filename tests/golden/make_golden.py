"""Generate tests/golden/*.npz|json by running the LIVE reference (/root/reference) on CPU.

Run in the build container only:   python tests/golden/make_golden.py
The reference cannot travel to the GPU box, so its outputs on deterministic synthetic
weights/inputs (tests/golden/synth.py) are committed as fixtures.  The reference's own tests
hold no golden vectors (it has no tests), so these fixtures ARE the parity pin.
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

from tests.golden.synth import (synth_images, synth_noise, synth_state_dict,  # noqa: E402
                                synth_timesteps)

from models.datasetDM_model import DatasetDM  # noqa: E402  (reference)
from models.diffusion_model import DiffusionModel  # noqa: E402  (reference)
from models.unet_model import Unet  # noqa: E402  (reference)

torch.set_grad_enabled(False)
torch.set_num_threads(os.cpu_count())


def shapes_of(m):
    return {k: tuple(v.shape) for k, v in m.state_dict().items()}


def load_synth(m, seed=0, skip=()):
    sd = synth_state_dict({k: s for k, s in shapes_of(m).items() if not k.startswith(skip)}, seed)
    missing = m.load_state_dict(sd, strict=False)
    assert not missing.unexpected_keys
    return sd


class FixedNoise:
    """Make the reference's torch.randn_like calls return our synthetic tensors, in order."""
    def __init__(self, tensors):
        self.q = list(tensors)
    def __enter__(self):
        self.orig = torch.randn_like
        torch.randn_like = lambda x, **kw: self.q.pop(0).to(x.dtype)
        return self
    def __exit__(self, *a):
        torch.randn_like = self.orig


def np_(x):
    return x.detach().cpu().numpy()


def main():
    out = {}
    # ---- schedule tables (bit-exact targets) --------------------------------------------
    sched = {}
    for kind in ("cosine", "linear"):
        m = DiffusionModel(Namespace(normalize=True, beta_schedule=kind, p2_loss_weight_gamma=1.0))
        for k, v in m.state_dict().items():
            if not k.startswith("model."):
                sched[f"{kind}.{k}"] = np_(v)
    np.savez_compressed(os.path.join(HERE, "schedule.npz"), **sched)

    # ---- state_dict inventories ----------------------------------------------------------
    inv = {"unet_default": {k: list(s) for k, s in shapes_of(Unet()).items()},
           "unet_mults12_outdim3": {k: list(s) for k, s in shapes_of(Unet(64, dim_mults=[1, 2], channels=2, out_dim=3)).items()}}
    cfg = Namespace(normalize=True, saved_diffusion_model="/nonexistent", verbose=False, device="cpu",
                    t_steps_to_save=[1, 10, 25, 50, 200, 400, 600, 800])
    dm = DatasetDM(cfg)
    inv["ledme"] = {k: list(s) for k, s in shapes_of(dm).items()}
    dm.classifier = tedm_head(len(dm.steps))
    inv["tedm"] = {k: list(s) for k, s in shapes_of(dm).items()}
    json.dump(inv, open(os.path.join(HERE, "state_dict_keys.json"), "w"), indent=0)

    # ---- small UNet / DDPM / sampler case: 32x32, B=2, default widths ---------------------
    B, S = 2, 32
    model = DiffusionModel(Namespace(normalize=True)).eval()
    load_synth(model, 0, skip=("sqrt_", "posterior_", "p2_"))
    x0 = synth_images(B, S, 0)
    t = synth_timesteps(B, 1000, 0)
    nz = synth_noise((B, 1, S, S), 0)
    g = {"x0": np_(x0), "t": np_(t), "noise": np_(nz)}
    x_t, _ = model.forward_diffusion_model(x0 * 2 - 1, t, nz)
    g["x_t"] = np_(x_t)
    feats = {}
    hooks = [a.register_forward_hook(lambda m_, i_, o_, i=i: feats.__setitem__(i, o_))
             for i, (_, _, a, _) in enumerate(model.model.ups)]
    g["unet_out"] = np_(model.model(x_t, t))
    for i in range(4):
        g[f"feat{i}"] = np_(feats[i])
    g["unet_out_t_none"] = np_(model.model(x_t, None))
    for h in hooks:
        h.remove()
    with FixedNoise([nz]):
        g["ddpm_loss"] = np_(model.train_step(x0, t=t))
    m2 = DiffusionModel(Namespace(normalize=True, p2_loss_weight_gamma=1.0)).eval()
    m2.load_state_dict({k: v for k, v in model.state_dict().items() if k.startswith("model.")}, strict=False)
    with FixedNoise([nz]):
        g["ddpm_loss_p2gamma1"] = np_(m2.train_step(x0, t=t))
    z = synth_noise((B, 1, S, S), 1, "z")
    for ts in (500, 999, 0):
        with FixedNoise([z]):
            g[f"sample_t{ts}"] = np_(model.sample_timestep(x_t, ts, cond=None))
    np.savez_compressed(os.path.join(HERE, "ddpm_small.npz"), **g)

    # ---- small TEDM / LEDM case: 32x32, B=2, 3 steps --------------------------------------
    steps = [10, 400, 800]
    cfg = Namespace(normalize=True, saved_diffusion_model="/nonexistent", verbose=False, device="cpu",
                    t_steps_to_save=steps)
    noises = [synth_noise((B, 1, S, S), 10 + i, "tedm") for i in range(len(steps))]
    g = {"x0": np_(x0), "steps": np.array(steps)}
    for i, n_ in enumerate(noises):
        g[f"noise{i}"] = np_(n_)
    led = DatasetDM(cfg).eval()
    load_synth(led, 0, skip=("diffusion_model.sqrt_", "diffusion_model.posterior_", "diffusion_model.p2_"))
    with FixedNoise(noises):
        g["ledm_logits"] = np_(led(x0))
    led.train()
    led.diffusion_model.eval()
    with FixedNoise(noises):
        g["ledm_logits_bn_train"] = np_(led(x0))
    ted = DatasetDM(cfg)
    ted.classifier = tedm_head(len(steps))
    ted.eval()
    load_synth(ted, 0, skip=("diffusion_model.sqrt_", "diffusion_model.posterior_", "diffusion_model.p2_"))
    with FixedNoise(noises):
        logits = ted(x0)
    g["tedm_logits"] = np_(logits)
    pr = torch.sigmoid(logits).reshape(B, len(steps), 1, S, S).mean(1)   # '(b step) 1 h w -> step b 1 h w', mean(0)
    g["tedm_prob"] = np_(pr)
    g["tedm_mask"] = np_(pr > 0.5)
    ted.train()
    ted.diffusion_model.eval()
    with FixedNoise(noises):
        g["tedm_logits_bn_train"] = np_(ted(x0))
    np.savez_compressed(os.path.join(HERE, "tedm_small.npz"), **g)

    # ---- full-size case (config.py defaults): 128x128, B=1, 8 TEDM steps -------------------
    steps = [1, 10, 25, 50, 200, 400, 600, 800]
    cfg.t_steps_to_save = steps
    S = 128
    x0 = synth_images(1, S, 3)
    noises = [synth_noise((1, 1, S, S), 20 + i, "tedm") for i in range(len(steps))]
    ted = DatasetDM(cfg)
    ted.classifier = tedm_head(len(steps))
    ted.eval()
    load_synth(ted, 0, skip=("diffusion_model.sqrt_", "diffusion_model.posterior_", "diffusion_model.p2_"))
    g = {"x0": np_(x0), "steps": np.array(steps)}
    with FixedNoise(noises):
        logits = ted(x0)
    g["tedm_logits"] = np_(logits)
    pr = torch.sigmoid(logits).reshape(1, len(steps), 1, S, S).mean(1)
    g["tedm_prob"] = np_(pr)
    g["tedm_mask"] = np_(pr > 0.5)
    t = torch.tensor([400])
    x_t, _ = ted.diffusion_model.forward_diffusion_model(x0, t, noises[5])
    feats = {}
    hooks = [a.register_forward_hook(lambda m_, i_, o_, i=i: feats.__setitem__(i, o_))
             for i, (_, _, a, _) in enumerate(ted.diffusion_model.model.ups)]
    g["unet_out_t400"] = np_(ted.diffusion_model.model(x_t, t))
    for i in range(4):
        f = feats[i]
        g[f"feat{i}_t400_first8ch"] = np_(f[:, :8])
        g[f"feat{i}_t400_norm"] = np_(f.norm())
        g[f"feat{i}_t400_chmean"] = np_(f.mean(dim=(0, 2, 3)))
    np.savez_compressed(os.path.join(HERE, "tedm_full.npz"), **g)

    # ---- heads that actually segment ------------------------------------------------------
    # A random-init head puts every logit at the decision threshold (|prob - 0.5| < 0.01 on ~64 % of
    # the pixels), so mask agreement there measures rounding noise.  These two fixtures train the
    # reference's TEDM head (reference modules, reference BCE loss, Adam) for a few dozen steps on the
    # frozen synthetic UNet's features against the mask (x0 > 0.45), then record the reference's
    # eval-mode outputs together with the trained head parameters.
    # (200 / 150 Adam steps: far short of the reference's full training runs, but enough that < 0.5 % of the
    # pixels sit within 0.02 of the 0.5 threshold, as for a converged segmenter.)
    for tag, (bsz, size, stps, seed0, iters) in {"small": (2, 32, [10, 400, 800], 10, 200),
                                                 "full": (1, 128, [1, 10, 25, 50, 200, 400, 600, 800], 20, 150)}.items():
        cfg.t_steps_to_save = stps
        x0 = synth_images(bsz, size, 0 if tag == "small" else 3)
        noises = [synth_noise((bsz, 1, size, size), seed0 + i, "tedm") for i in range(len(stps))]
        ted = DatasetDM(cfg)
        ted.classifier = tedm_head(len(stps))
        ted.eval()
        load_synth(ted, 0, skip=("diffusion_model.sqrt_", "diffusion_model.posterior_", "diffusion_model.p2_"))
        with FixedNoise(noises):
            feats = ted.extract_features(x0)
        y = (x0 > 0.45).float()
        yr = y.repeat_interleave(len(stps), dim=0)                      # 'b c h w -> (b step) c h w'
        torch.manual_seed(0)
        with torch.enable_grad():
            opt = torch.optim.Adam(ted.classifier.parameters(), lr=2e-3)
            ted.classifier.train()
            for it in range(iters):
                opt.zero_grad()
                out = ted.classifier(feats)
                loss = torch.nn.functional.binary_cross_entropy_with_logits(out, yr, reduction="none").mean(dim=(2, 3)).mean()
                loss.backward()
                opt.step()
        ted.classifier.eval()
        logits = ted.classifier(feats)
        pr = torch.sigmoid(logits).reshape(bsz, len(stps), 1, size, size).mean(1)
        g = {"x0": np_(x0), "steps": np.array(stps), "target": np_(y), "final_train_loss": np_(loss.detach()),
             "tedm_logits": np_(logits), "tedm_prob": np_(pr), "tedm_mask": np_(pr > 0.5)}
        for i, n_ in enumerate(noises):
            g[f"noise{i}"] = np_(n_)
        for k, v in ted.classifier.state_dict().items():
            g[f"classifier.{k}"] = np_(v)
        print(tag, "trained head: loss", float(loss), "logit std", float(logits.std()),
              "frac |p-.5|<.02:", float(((pr - .5).abs() < .02).float().mean()),
              "acc vs target", float(((pr > .5).float() == y).float().mean()))
        np.savez_compressed(os.path.join(HERE, f"tedm_{tag}_trained.npz"), **g)
    print("golden fixtures written:", sorted(f for f in os.listdir(HERE) if f.endswith((".npz", ".json"))))


def tedm_head(n_steps):
    """The shared-weight head, built exactly as the reference's trainer does (train_datasetDM.py:30-42)."""
    import torch.nn as nn
    from einops.layers.torch import Rearrange
    return nn.Sequential(Rearrange('b (step act) h w -> (b step) act h w', step=n_steps),
                         nn.Conv2d(960, 128, 1), nn.ReLU(), nn.BatchNorm2d(128),
                         nn.Conv2d(128, 32, 1), nn.ReLU(), nn.BatchNorm2d(32), nn.Conv2d(32, 1, 1))


if __name__ == "__main__":
    main()
