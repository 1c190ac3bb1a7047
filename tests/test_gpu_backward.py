"""Backward kernels against torch autograd on the same bf16-rounded operands."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _rand(shape, seed, scale=1.0):
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed)) * scale


def _bf(x):
    return x.to(torch.bfloat16).float()


def _nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).cuda()


def _ref_fwd(x, w, mode):
    if mode == 0:
        return F.conv2d(x, w)
    if mode == 1:
        return F.conv2d(x, w, padding=1)
    if mode == 2:
        return F.conv2d(x, w, stride=2, padding=1)
    return F.conv2d(F.interpolate(x, scale_factor=2, mode="nearest"), w, padding=1)


def _unfold_mode3(dwf):
    """(cout, 16, cin) folded-kernel gradient -> (cout, cin, 3, 3)."""
    cout, _, cin = dwf.shape
    d = dwf.reshape(cout, 2, 2, 2, 2, cin)          # [co][py][px][a][b][ci]
    out = torch.zeros(cout, cin, 3, 3, dtype=dwf.dtype, device=dwf.device)
    amap = {0: [0, 1, 1], 1: [0, 0, 1]}             # parity -> tap index (a) of ky = 0, 1, 2
    for py in range(2):
        for px in range(2):
            for ky in range(3):
                for kx in range(3):
                    out[:, :, ky, kx] += d[:, py, px, amap[py][ky], amap[px][kx], :]
    return out


WG_CASES = [
    # (B, H, W, c0, c1, cout, mode, force_bn)
    (3, 16, 16, 64, 0, 128, 0, 0),
    (1, 128, 128, 64, 0, 64, 1, 0),
    (2, 32, 32, 128, 64, 128, 1, 0),
    (5, 4, 4, 512, 0, 512, 1, 0),
    (2, 16, 16, 64, 0, 64, 1, 0),
    (2, 32, 32, 64, 0, 128, 2, 0),
    (2, 16, 16, 128, 0, 64, 3, 0),
    (2, 8, 8, 256, 128, 256, 1, 128),
    (9, 8, 8, 192, 0, 384, 0, 0),
]


@pytest.mark.parametrize("case", WG_CASES, ids=[str(c) for c in WG_CASES])
def test_conv_wgrad(case):
    from tedm_b200 import native as N
    B, H, W, c0, c1, cout, mode, force_bn = case
    kh = {0: 1, 1: 3, 2: 4, 3: 3}[mode]
    cin = c0 + c1
    x = _bf(_rand((B, cin, H, W), 1)).cuda()
    w = _rand((cout, cin, kh, kh), 2, (cin * kh * kh) ** -0.5).cuda().requires_grad_(True)
    y = _ref_fwd(x, w, mode)
    dy = _bf(_rand(tuple(y.shape), 3)).cuda()
    (dw_ref,) = torch.autograd.grad(y, w, dy)
    xs = _nhwc(x.cpu())
    x0, x1 = (xs[..., :c0].contiguous(), xs[..., c0:].contiguous()) if c1 else (xs, None)
    N.load().tedm_conv_set_tile_n(force_bn)
    try:
        dw = N.conv_wgrad(x0, _nhwc(dy.cpu()), mode, src1=x1)
    finally:
        N.load().tedm_conv_set_tile_n(0)
    torch.cuda.synchronize()
    if mode == 3:
        got = _unfold_mode3(dw)
    else:
        got = dw.reshape(cout, kh, kh, cin).permute(0, 3, 1, 2)
    assert _rel(got, dw_ref) < 2e-3, _rel(got, dw_ref)
