"""Generate tests/golden/ddpm_small_grads.npz by running the LIVE reference's training step
(DiffusionModel.train_step + loss.backward(), trainers/train_CXR14.py:30-40) on CPU.

Run in the build container only:   python tests/golden/make_golden_grads.py
Inputs and weights are those of ddpm_small.npz (tests/golden/synth.py).  The full gradient is 36 M floats, so the
fixture keeps, per parameter: the L2 norm, the sum, and a deterministic strided sample of at most 4096 elements.
It also keeps the parameters after one torch.optim.Adam step (same sampling) for the optimiser check.
"""
import os
import sys
from argparse import Namespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, "/root/reference")

from tests.golden.synth import synth_images, synth_noise, synth_state_dict, synth_timesteps  # noqa: E402
from tests.golden.make_golden import FixedNoise, load_synth  # noqa: E402

from models.diffusion_model import DiffusionModel  # noqa: E402  (reference)

from tests.golden.make_golden_grads_idx import sample_idx  # noqa: E402


def main():
    torch.set_grad_enabled(True)
    torch.set_num_threads(os.cpu_count())
    B, S = 2, 32
    model = DiffusionModel(Namespace(normalize=True)).train()
    load_synth(model, 0, skip=("sqrt_", "posterior_", "p2_"))
    x0 = synth_images(B, S, 0)
    t = synth_timesteps(B, 1000, 0)
    nz = synth_noise((B, 1, S, S), 0)
    opt = torch.optim.Adam(model.parameters(), lr=1e-4)
    with FixedNoise([nz]):
        loss = model.train_step(x0, t=t)
    loss.backward()
    out = {"loss": loss.detach().numpy()}
    for name, p in model.named_parameters():
        g = p.grad.detach().reshape(-1).double()
        out[f"norm/{name}"] = np.float64(g.norm().item())
        out[f"sum/{name}"] = np.float64(g.sum().item())
        out[f"sample/{name}"] = g[torch.from_numpy(sample_idx(g.numel()))].float().numpy()
    opt.step()
    for name, p in model.named_parameters():
        v = p.detach().reshape(-1)
        out[f"adam/{name}"] = v[torch.from_numpy(sample_idx(v.numel()))].numpy()
    np.savez_compressed(os.path.join(HERE, "ddpm_small_grads.npz"), **out)
    print("loss", float(loss), "params", sum(1 for _ in model.parameters()))


if __name__ == "__main__":
    main()
