"""Does replaying the TEDM inference step from a CUDA graph beat issuing its ~600 launches from Python?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from argparse import Namespace
import torch
import bench
from tedm_b200.models import DatasetDM, tedm_classifier
S = len(bench.STEPS_TEDM); B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = torch.device("cuda")
m = DatasetDM(Namespace(normalize=True, saved_diffusion_model="", t_steps_to_save=bench.STEPS_TEDM))
m.classifier = tedm_classifier(S)
m.load_state_dict(bench.synth_state(S), strict=False)
m = m.eval().to(dev)
xs = [bench.synth_batch(B, i).to(dev) for i in range(4)]
nz = [torch.randn(B, 1, 128, 128, device=dev) for i in range(4)]
def timeit(fn, n=10):
    for i in range(3): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
t_eager = timeit(lambda i: m.segment(xs[i % 4], nz[i % 4]))
t_graph = timeit(lambda i: m.segment(xs[i % 4], nz[i % 4], graph=True))
ref = m.segment(xs[1], nz[1])
got = m.segment(xs[1], nz[1], graph=True)
assert all(torch.equal(a, b) for a, b in zip(ref, got)), "graph replay differs from the eager call"
print(f"B={B}: eager {t_eager:.3f} ms/step ({B / t_eager * 1e3:.0f} img/s), graph {t_graph:.3f} ms/step ({B / t_graph * 1e3:.0f} img/s)")
