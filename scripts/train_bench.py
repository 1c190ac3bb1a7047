"""Time the DDPM training step (q_sample + UNet fwd + L1 loss + UNet bwd + Adam) at 128x128."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from argparse import Namespace
import torch
from tedm_b200.models import DiffusionModel
from tedm_b200.optim import FusedAdam
from tedm_b200 import native as N

GFLOP_PER_IMG = 176.9
def main():
    batches = [int(a) for a in sys.argv[1:]] or [16, 64]
    torch.manual_seed(0)
    m = DiffusionModel(Namespace(normalize=True)).cuda().train()
    opt = FusedAdam(m.parameters(), lr=1e-4)
    for B in batches:
        x = torch.rand(B, 1, 128, 128, device="cuda")
        if os.environ.get("GRAPH", "1") == "1":
            from tedm_b200.train import GraphedTrainStep
            gstep = GraphedTrainStep(m, opt, x)
            step = lambda: gstep(x)
        else:
            def step():
                opt.zero_grad(set_to_none=True)
                loss = m.train_step(x)
                loss.backward()
                opt.step()
                return loss
        for _ in range(int(os.environ.get('WARM', '3'))):
            step()
        torch.cuda.synchronize()
        n = int(os.environ.get('STEPS', '5'))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = N.launches
        t0 = time.perf_counter()
        e0.record()
        for _ in range(n):
            loss = step()
        e1.record()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / n * 1e3
        ms = e0.elapsed_time(e1) / n
        print(f"B={B}: {ms:.2f} ms/step (wall {wall:.2f}), {B / ms * 1e3:.1f} img/s, {B * GFLOP_PER_IMG / ms:.1f} TFLOP/s fwd+bwd, "
              f"calls/step {(N.launches - l0) // n}, loss {loss.item():.4f}, mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)
main()
