import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from argparse import Namespace
import torch
from tedm_b200.models import DiffusionModel
from tedm_b200.trainers.utils import GraphedSampler
m = DiffusionModel(Namespace(normalize=True)).cuda().eval()
for B in (1, 8, 64):
    x = torch.randn(B, 1, 128, 128, device="cuda")
    for t in range(999, 994, -1): x = m.sample_timestep(x, t)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for t in range(994, 944, -1): x = m.sample_timestep(x, t)
    torch.cuda.synchronize(); eager = (time.perf_counter() - t0) / 50 * 1e3
    gs = GraphedSampler(m, B, 1, 128)
    gs.x.normal_()
    for t in range(999, 994, -1): gs.step(t)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for t in range(994, 944, -1): gs.step(t)
    torch.cuda.synchronize(); graph = (time.perf_counter() - t0) / 50 * 1e3
    print(f"B={B}: eager {eager:.2f} ms per reverse step, graph {graph:.2f} ms  ({eager / graph:.2f}x)")
