import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from argparse import Namespace
import torch, bench
from tedm_b200 import native as N
from tedm_b200.models import DatasetDM, tedm_classifier
dev = torch.device("cuda")
x = torch.rand(16, 1, 128, 128, device=dev); nz = torch.randn(16, 1, 128, 128, device=dev)
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = N.launches
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / n, 3), (N.launches - l0) // n
for name, steps, shared in (("TEDM", bench.STEPS_TEDM, True), ("LEDMe", bench.STEPS_TEDM, False), ("LEDM", [50, 150, 250], False)):
    m = DatasetDM(Namespace(normalize=True, saved_diffusion_model="", t_steps_to_save=steps))
    if shared: m.classifier = tedm_classifier(len(steps))
    m = m.to(dev).eval()
    print(name, "features", timeit(lambda: m.feature_maps(x, nz)), "segment", timeit(lambda: m.segment(x, nz)))
