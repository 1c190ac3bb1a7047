#!/usr/bin/env python
"""Time (CUDA events) one conv_igemm shape; the ncu target for the dominant kernel.
usage: python scripts/prof_conv.py MODE B H W C0 C1 COUT [GN] [iters]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tedm_b200 import native as N  # noqa: E402


def main():
    mode, B, H, W, c0, c1, cout = (int(a) for a in sys.argv[1:8])
    gn = int(sys.argv[8]) if len(sys.argv) > 8 else 0
    iters = int(sys.argv[9]) if len(sys.argv) > 9 else 20
    taps = {0: 1, 1: 9, 2: 16, 3: 16}[mode]
    g = torch.Generator(device="cuda").manual_seed(0)
    # several input sets so consecutive launches do not hit in L2 (126 MB)
    nset = max(2, int(300e6 // (B * H * W * (c0 + c1) * 2)) + 1)
    xs = [torch.randn(B, H, W, c0, device="cuda", generator=g).to(torch.bfloat16) for _ in range(nset)]
    ys = [torch.randn(B, H, W, c1, device="cuda", generator=g).to(torch.bfloat16) for _ in range(nset)] if c1 else None
    w = (torch.randn(cout * taps * (c0 + c1), device="cuda", generator=g) * 0.05).to(torch.bfloat16)
    bias = torch.zeros(cout, device="cuda")
    oh, ow = (H // 2, W // 2) if mode == 2 else ((2 * H, 2 * W) if mode == 3 else (H, W))
    out = torch.empty(B, oh, ow, cout, device="cuda", dtype=torch.bfloat16)
    run = lambda i: N.conv_igemm(xs[i % nset], w, mode, cout, bias=bias, src1=ys[i % nset] if c1 else None,
                                 gn_groups=gn, out=out)
    for i in range(5):
        run(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        run(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    flops = 2 * B * (H * W if mode == 3 else oh * ow) * cout * taps * (c0 + c1)
    byts = B * H * W * (c0 + c1) * 2 + B * oh * ow * cout * 2
    print(f"conv mode={mode} B={B} {H}x{W} {c0}+{c1}->{cout} gn={gn}: {ms * 1e3:.1f} us  {flops / ms / 1e9:.1f} TFLOP/s  "
          f"{byts / ms / 1e6:.0f} GB/s (algorithmic in+out bytes)")


if __name__ == "__main__":
    main()
