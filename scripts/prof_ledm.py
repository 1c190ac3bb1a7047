import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from argparse import Namespace
import torch
from tedm_b200 import native as N
from tedm_b200.autograd import bce_with_logits_rows
from tedm_b200.models import DatasetDM
dev = torch.device("cuda")
m = DatasetDM(Namespace(normalize=True, saved_diffusion_model="", t_steps_to_save=[50, 150, 250])).to(dev)
m.diffusion_model.eval()
x = torch.rand(16, 1, 128, 128, device=dev); y = (torch.rand(16, 1, 128, 128, device=dev) > .5).float()
def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = N.launches
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (N.launches - l0) // n
m.eval()
print("features only", timeit(lambda: m.feature_maps(x)))
print("eval forward", timeit(lambda: m(x)))
m.train(); m.diffusion_model.eval()
print("train forward", timeit(lambda: m(x)))
def fb():
    for p in m.classifier.parameters(): p.grad = None
    bce_with_logits_rows(m(x), y).mean().backward()
print("train fwd+bwd", timeit(fb))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    fb(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=60))
from tedm_b200.optim import FusedAdam
opt = FusedAdam(m.classifier.parameters(), lr=1e-4)
def full():
    opt.zero_grad()
    loss = bce_with_logits_rows(m(x), y).mean()
    loss.backward()
    opt.step()
print("full step", timeit(full))
def nostep():
    opt.zero_grad()
    loss = bce_with_logits_rows(m(x), y).mean()
    loss.backward()
print("without opt.step", timeit(nostep))
import time
torch.cuda.synchronize(); t0 = time.perf_counter(); opt.step(); torch.cuda.synchronize(); print("opt.step wall ms", (time.perf_counter() - t0) * 1e3)
torch.cuda.synchronize(); t0 = time.perf_counter(); m(x); torch.cuda.synchronize(); print("forward after step wall ms", (time.perf_counter() - t0) * 1e3)
torch.cuda.synchronize(); t0 = time.perf_counter(); m(x); torch.cuda.synchronize(); print("forward again wall ms", (time.perf_counter() - t0) * 1e3)
