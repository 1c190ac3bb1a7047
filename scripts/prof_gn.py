"""Time gn_silu fwd / bwd per shape (CUDA events, rotating buffers larger than L2)."""
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

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
for (H, C) in [(128, 64), (64, 128), (32, 256), (16, 512)]:
    nset = max(2, int(400e6 // (B * H * H * C * 2)) + 1)
    xs = [torch.randn(B, H, H, C, device="cuda").to(torch.bfloat16) for _ in range(nset)]
    dys = [torch.randn(B, H, H, C, device="cuda").to(torch.bfloat16) for _ in range(nset)]
    parts = N.conv_gn_parts(H, H)
    part = torch.rand(B, parts, 8, 2, device="cuda")
    part[..., 1] += 200.0
    gamma, beta = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
    ss = torch.randn(B, 2 * C, device="cuda") * 0.1
    dg, db, dbias, dss = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda"), torch.zeros_like(ss)
    nbytes = B * H * H * C * 2
    t_f = timeit(lambda i: N.gn_silu(xs[i % nset], part, gamma, beta, 8, scale_shift=ss))
    t_fr = timeit(lambda i: N.gn_silu(xs[i % nset], part, gamma, beta, 8, scale_shift=ss, residual=dys[i % nset]))
    t_b = timeit(lambda i: N.gn_silu_bwd(xs[i % nset], dys[i % nset], part, gamma, beta, 8, dg, db, dbias, scale_shift=ss, dscale_shift=dss))
    print(f"B={B} {H}x{H}x{C}: fwd {t_f*1e3:.1f} us ({2*nbytes/t_f/1e6:.0f} GB/s), fwd+res {t_fr*1e3:.1f} us ({3*nbytes/t_fr/1e6:.0f} GB/s), "
          f"bwd(reduce+apply) {t_b*1e3:.1f} us ({6*nbytes/t_b/1e6:.0f} GB/s of 12 B/elem)")
