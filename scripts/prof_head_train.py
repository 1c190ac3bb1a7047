import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from argparse import Namespace
import torch, bench
from tedm_b200.autograd import bce_with_logits_rows
from tedm_b200.models import DatasetDM, tedm_classifier
from tedm_b200.optim import FusedAdam
dev = torch.device("cuda")
x = torch.rand(16, 1, 128, 128, device=dev); y = (torch.rand(16, 1, 128, 128, device=dev) > .5).float()
m = DatasetDM(Namespace(normalize=True, saved_diffusion_model="", t_steps_to_save=bench.STEPS_TEDM))
m.classifier = tedm_classifier(8)
m = m.to(dev).train(); m.diffusion_model.eval()
opt = FusedAdam(m.classifier.parameters(), lr=1e-4)
ts = []
for i in range(30):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    opt.zero_grad()
    loss = bce_with_logits_rows(m(x), y).mean()
    torch.cuda.synchronize(); t1 = time.perf_counter()
    loss.backward()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    opt.step()
    torch.cuda.synchronize(); t3 = time.perf_counter()
    ts.append((round((t1 - t0) * 1e3, 1), round((t2 - t1) * 1e3, 1), round((t3 - t2) * 1e3, 1)))
print(ts)
print(torch.cuda.memory_stats()["num_alloc_retries"], torch.cuda.memory_stats()["num_device_alloc"], torch.cuda.memory_stats()["num_device_free"],
      torch.cuda.max_memory_allocated() / 2**30, torch.cuda.memory_reserved() / 2**30)
