import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tedm_b200 import native as N
def timeit(fn, iters=20):
    for i in range(3): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
for (H, C) in [(128, 64), (64, 128), (32, 256), (16, 512)]:
    nset = max(2, int(400e6 // (B * H * H * C * 2)) + 1)
    xs = [torch.randn(B, H, H, C, device="cuda").to(torch.bfloat16) for _ in range(nset)]
    rs = [torch.randn(B, H, H, C, device="cuda").to(torch.bfloat16) for _ in range(nset)]
    g = torch.ones(C, device="cuda"); dg = torch.zeros(C, device="cuda")
    nb = B * H * H * C * 2
    t1 = timeit(lambda i: N.layernorm(xs[i % nset], g))
    t2 = timeit(lambda i: N.layernorm(xs[i % nset], g, residual=rs[i % nset]))
    t3 = timeit(lambda i: N.layernorm_bwd(xs[i % nset], g, rs[i % nset], dg, add=rs[(i + 1) % nset]))
    print(f"B={B} {H}x{H}x{C}: ln {t1*1e3:.1f} us ({2*nb/t1/1e6:.0f} GB/s)  ln+res {t2*1e3:.1f} us ({3*nb/t2/1e6:.0f} GB/s)  bwd+add {t3*1e3:.1f} us ({4*nb/t3/1e6:.0f} GB/s)")
