"""Which half of the step the halo-tile convs perturb: gradients with halo tiles on/off in the forward and in the backward separately (found the flipped L1 sign)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from argparse import Namespace
from tedm_b200 import native as N
from tedm_b200.models import DiffusionModel
from golden.synth import synth_state_dict
from oracle import tedm_oracle as O

g = np.load("tests/golden/ddpm_small.npz")
T = lambda a: torch.from_numpy(np.asarray(a))
def model():
    m = DiffusionModel(Namespace(normalize=True)).train()
    m.load_state_dict(synth_state_dict(O.unet_param_shapes(prefix="model."), 0), strict=False)
    return m.cuda()
x0, t, nz = T(g["x0"]).cuda(), T(g["t"]).cuda(), T(g["noise"]).cuda()
def grads(fwd_halo, bwd_halo):
    m = model()
    N.load().tedm_conv_set_halo(fwd_halo)
    loss = m.train_step(x0, t=t, noise=nz)
    torch.cuda.synchronize()
    N.load().tedm_conv_set_halo(bwd_halo)
    loss.backward()
    torch.cuda.synchronize()
    return loss.item(), {n: p.grad.clone() for n, p in m.named_parameters()}
l0, g0 = grads(0, 0)
for fh, bh in ((1, 0), (0, 1), (1, 1)):
    l1, g1 = grads(fh, bh)
    errs = sorted(((float((g1[n] - g0[n]).norm() / g0[n].norm().clamp_min(1e-20)), n) for n in g0), reverse=True)
    tot = (sum(float((g1[n] - g0[n]).pow(2).sum()) for n in g0) / sum(float(g0[n].pow(2).sum()) for n in g0)) ** 0.5
    print(f"fwd_halo={fh} bwd_halo={bh}: loss {l0:.6f} -> {l1:.6f}; whole-gradient diff {tot:.4f}; worst", [(round(e, 4), n) for e, n in errs[:5]])
    print("   smallest", [(round(e, 5), n) for e, n in errs[-3:]])
