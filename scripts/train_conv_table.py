"""Per-shape table of every tcgen05 conv launch (forward, data gradient, weight gradient) of one training step.
usage: python scripts/train_conv_table.py [B] > table.txt"""
import contextlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from argparse import Namespace
import torch
from tedm_b200.models import DiffusionModel
from tedm_b200 import native as N

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
torch.manual_seed(0)
m = DiffusionModel(Namespace(normalize=True)).cuda().train()
x = torch.rand(B, 1, 128, 128, device="cuda")
for _ in range(2):
    m.zero_grad(set_to_none=True)
    m.train_step(x).backward()
torch.cuda.synchronize()
records = []

@contextlib.contextmanager
def timer(flops, shape=None):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    yield
    b.record()
    records.append((flops, a, b, shape))

N.conv_timer = timer
m.zero_grad(set_to_none=True)
m.train_step(x).backward()
N.conv_timer = None
torch.cuda.synchronize()
agg = {}
for f, a, b, shp in records:
    e = agg.setdefault(shp, [0, 0.0, 0.0])
    e[0] += 1; e[1] += a.elapsed_time(b); e[2] += f
tot = sum(e[1] for e in agg.values())
totf = sum(e[2] for e in agg.values())
print(f"B={B}: {len(records)} conv launches, {tot:.2f} ms, {totf / tot / 1e9:.1f} TFLOP/s overall")
for kind in ("fwd/dgrad", "wgrad"):
    sel = {k: v for k, v in agg.items() if (k[0] == "wgrad") == (kind == "wgrad")}
    t = sum(v[1] for v in sel.values()); fl = sum(v[2] for v in sel.values())
    print(f"--- {kind}: {t:.2f} ms, {fl / t / 1e9:.1f} TFLOP/s")
    print("shape (mode B H W c0 c1 cout [gn res]) | launches  ms_total  TFLOP/s  share")
    for shp, (cnt, t_ms, f) in sorted(sel.items(), key=lambda kv: -kv[1][1]):
        print(f"{shp} | {cnt:3d} {t_ms:9.3f} {f / (t_ms * 1e-3) / 1e12:8.1f} {100 * t_ms / tot:6.1f}%")
