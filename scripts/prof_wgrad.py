"""Time conv_wgrad per shape: halo kernel on/off, OIHW-accumulate vs scratch layout."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tedm_b200 import native as N

def timeit(fn, iters=10):
    for i in range(2): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

B = 64
shapes = [(128, 64, 0, 64), (128, 64, 64, 64), (64, 64, 0, 64), (64, 128, 0, 128), (64, 128, 64, 128), (32, 128, 0, 128),
          (32, 256, 0, 256), (32, 256, 128, 256), (16, 256, 0, 256), (16, 512, 0, 512), (16, 512, 256, 512)]
if len(sys.argv) > 1:
    shapes = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]]
for (H, c0, c1, cout) in shapes:
    nset = max(2, int(300e6 // (B * H * H * (c0 + c1 + cout) * 2)) + 1)
    xs = [torch.randn(B, H, H, c0, device="cuda").to(torch.bfloat16) for _ in range(nset)]
    x1 = [torch.randn(B, H, H, c1, device="cuda").to(torch.bfloat16) for _ in range(nset)] if c1 else None
    dys = [torch.randn(B, H, H, cout, device="cuda").to(torch.bfloat16) for _ in range(nset)]
    grad = torch.zeros(cout, c0 + c1, 3, 3, device="cuda")
    fl = 2 * B * H * H * cout * 9 * (c0 + c1)
    res = []
    for halo in (0, 2):
        N.load().tedm_conv_set_wgrad_halo(halo)
        for oihw in (0, 1):  # (with the workspace the reduce kernel does the layout)
            t = timeit(lambda i: N.conv_wgrad(xs[i % nset], dys[i % nset], 1, src1=x1[i % nset] if c1 else None,
                                              grad_oihw=grad if oihw else None))
            res.append(f"halo={halo} oihw={oihw}: {t*1e3:7.1f} us {fl/t/1e9:6.0f} TF")
    N.load().tedm_conv_set_wgrad_halo(1)
    print(f"{H}x{H} {c0}+{c1}->{cout} | " + " | ".join(res), flush=True)
