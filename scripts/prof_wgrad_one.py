"""One 3x3 weight-gradient launch set for ncu: python scripts/prof_wgrad_one.py H C0 C1 COUT [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tedm_b200 import native as N
H, c0, c1, cout = (int(a) for a in sys.argv[1:5])
B = int(sys.argv[5]) if len(sys.argv) > 5 else 64
x = torch.randn(B, H, H, c0, device="cuda").to(torch.bfloat16)
x1 = torch.randn(B, H, H, c1, device="cuda").to(torch.bfloat16) if c1 else None
dy = torch.randn(B, H, H, cout, device="cuda").to(torch.bfloat16)
grad = torch.zeros(cout, c0 + c1, 3, 3, device="cuda")
for _ in range(3):
    N.conv_wgrad(x, dy, 1, src1=x1, grad_oihw=grad)
torch.cuda.synchronize()
print("done")
