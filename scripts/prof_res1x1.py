"""1x1 res_conv variants at the bench sizes: plain, + residual, + residual normalised in the epilogue (GroupNorm + SiLU), next
to the GroupNorm + SiLU + add pass they replace.  CUDA events, L2 flushed by the working set (tensors >= 268 MB)."""
import sys
import torch
from tedm_b200 import native as N

def t_us(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
for (H, c0, c1, cout) in [(128, 64, 64, 64), (64, 128, 64, 128), (32, 256, 128, 256), (16, 512, 256, 512)]:
    g = torch.Generator(device="cuda").manual_seed(0)
    r = lambda *s: torch.randn(*s, device="cuda", generator=g).to(torch.bfloat16)
    x0, x1, h2 = r(B, H, H, c0), r(B, H, H, c1), r(B, H, H, cout)
    w = r(cout, 1, 1, c0 + c1)
    bias = torch.zeros(cout, device="cuda")
    gamma, beta = torch.ones(cout, device="cuda"), torch.zeros(cout, device="cuda")
    part = torch.zeros(B, N.conv_gn_parts(H, H), 8, 2, device="cuda")
    part[..., 1] = 1.0
    aff = N.gn_affine(part, gamma, beta, 8, H * H)
    res = N.conv_igemm(x0, w, 0, cout, bias=bias, src1=x1)
    mb = lambda n: n * B * H * H * 2 / 1e6
    rows = [("conv1x1", lambda: N.conv_igemm(x0, w, 0, cout, bias=bias, src1=x1), mb(c0 + c1 + cout)),
            ("gn_silu+add pass", lambda: N.gn_silu(h2, part, gamma, beta, 8, residual=res), mb(3 * cout)),
            ("conv1x1 + residual", lambda: N.conv_igemm(x0, w, 0, cout, bias=bias, src1=x1, residual=h2), mb(c0 + c1 + 2 * cout)),
            ("conv1x1 + SiLU(GN(residual))", lambda: N.conv_igemm(x0, w, 0, cout, bias=bias, src1=x1, residual=h2, residual_affine=aff), mb(c0 + c1 + 2 * cout))]
    for name, fn, mbytes in rows:
        us = t_us(fn)
        print(f"B={B} {H}x{H} ({c0}+{c1})->{cout}  {name:32s} {us:8.1f} us  {mbytes / us * 1e6 / 1e6:7.0f} GB/s")
