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
    # halo-tile 3x3 kernel: every tile geometry (128/64/32/16/8-pixel-wide tiles), ragged batch, two sources
    (3, 64, 64, 64, 0, 128, 1, 0),
    (5, 32, 32, 128, 64, 64, 1, 0),
    (3, 8, 8, 64, 0, 64, 1, 0),
    (2, 128, 128, 64, 64, 64, 1, 0),
    (7, 16, 16, 512, 0, 64, 1, 0),
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
    N.load().tedm_conv_set_wgrad_halo(2)        # the halo-tile kernel wherever its geometry allows, not only where it pays
    try:
        dw = N.conv_wgrad(x0, _nhwc(dy.cpu()), mode, src1=x1)
        if mode == 1:                           # and the generic kernel on the same inputs
            N.load().tedm_conv_set_wgrad_halo(0)
            dw_generic = N.conv_wgrad(x0, _nhwc(dy.cpu()), mode, src1=x1)
            assert _rel(dw_generic, dw) < 1e-4
    finally:
        N.load().tedm_conv_set_tile_n(0)
        N.load().tedm_conv_set_wgrad_halo(1)
    torch.cuda.synchronize()
    if mode == 3:
        got = _unfold_mode3(dw)
    else:
        got = dw.reshape(cout, kh, kh, cin).permute(0, 3, 1, 2)
    assert _rel(got, dw_ref) < 2e-3, _rel(got, dw_ref)


# ---- weight re-layouts: dgrad through the forward kernel ------------------------------------------
DG_CASES = [
    # (B, H, W, cin, cout, fwd mode)
    (2, 16, 16, 128, 64, 0),
    (2, 16, 16, 64, 128, 1),
    (1, 128, 128, 64, 64, 1),
    (2, 32, 32, 64, 128, 2),
    (2, 16, 16, 128, 64, 3),
]


@pytest.mark.parametrize("mode,halo", [(1, 2), (1, 0), (3, 0), (2, 0), (0, 0)])
def test_conv_wgrad_accumulates_into_oihw(mode, halo):
    """oihw_accumulate: the kernel adds straight into the fp32 OIHW parameter gradient (mode 3 un-folds to 3x3)."""
    from tedm_b200 import native as N
    B, H, W, cin, cout = 3, 16, 16, 128, 64
    kh = {0: 1, 1: 3, 2: 4, 3: 3}[mode]
    x = _bf(_rand((B, cin, H, W), 1)).cuda()
    w = _rand((cout, cin, kh, kh), 2, (cin * kh * kh) ** -0.5).cuda().requires_grad_(True)
    y = _ref_fwd(x, w, mode)
    dy = _bf(_rand(tuple(y.shape), 3)).cuda()
    (dw_ref,) = torch.autograd.grad(y, w, dy)
    grad = _rand((cout, cin, kh, kh), 4).cuda()
    base = grad.clone()
    N.load().tedm_conv_set_wgrad_halo(halo)
    try:
        N.conv_wgrad(_nhwc(x.cpu()), _nhwc(dy.cpu()), mode, grad_oihw=grad)
    finally:
        N.load().tedm_conv_set_wgrad_halo(1)
    torch.cuda.synchronize()
    assert _rel(grad - base, dw_ref) < 2e-3, _rel(grad - base, dw_ref)


@pytest.mark.parametrize("case", DG_CASES, ids=[str(c) for c in DG_CASES])
def test_conv_dgrad_via_forward_kernel(case):
    from tedm_b200 import native as N
    B, H, W, cin, cout, mode = case
    kh = {0: 1, 1: 3, 2: 4, 3: 3}[mode]
    x = _bf(_rand((B, cin, H, W), 1)).cuda().requires_grad_(True)
    w = _bf(_rand((cout, cin, kh, kh), 2, (cin * kh * kh) ** -0.5)).cuda()
    y = _ref_fwd(x, w, mode)
    dy = _bf(_rand(tuple(y.shape), 3)).cuda()
    (dx_ref,) = torch.autograd.grad(y, x, dy)
    wd = N.weight_to_dgrad(w, mode)
    run_mode = {0: N.MODE_1X1, 1: N.MODE_3X3, 2: N.MODE_UP3X3, 3: N.MODE_4X4S2}[mode]
    dx = N.conv_igemm(_nhwc(dy.cpu()), wd, run_mode, cin)
    torch.cuda.synchronize()
    got = dx.float().permute(0, 3, 1, 2)
    assert _rel(got, dx_ref) < 6e-3, _rel(got, dx_ref)


def test_conv_dgrad_two_sources_and_residual():
    """Skip-concat conv: the gradient splits by weight row ranges; `residual` accumulates a second gradient."""
    from tedm_b200 import native as N
    B, H, W, c0, c1, cout = 2, 16, 16, 128, 64, 128
    x = _bf(_rand((B, c0 + c1, H, W), 1)).cuda().requires_grad_(True)
    w = _bf(_rand((cout, c0 + c1, 3, 3), 2, ((c0 + c1) * 9) ** -0.5)).cuda()
    y = F.conv2d(x, w, padding=1)
    dy = _bf(_rand(tuple(y.shape), 3)).cuda()
    (dx_ref,) = torch.autograd.grad(y, x, dy)
    wd = N.weight_to_dgrad(w, 1)
    extra = _bf(_rand((B, c1, H, W), 4)).cuda()
    dyn = _nhwc(dy.cpu())
    dx0 = N.conv_igemm(dyn, wd[: c0 * 9 * cout], N.MODE_3X3, c0)
    dx1 = N.conv_igemm(dyn, wd[c0 * 9 * cout:], N.MODE_3X3, c1, residual=_nhwc(extra.cpu()))
    torch.cuda.synchronize()
    assert _rel(dx0.float().permute(0, 3, 1, 2), dx_ref[:, :c0]) < 6e-3
    assert _rel(dx1.float().permute(0, 3, 1, 2), dx_ref[:, c0:] + extra) < 6e-3
    # one launch, split output: 64-channel sub-tiles of the N tile are routed to the two gradient tensors
    for force_bn in (0, 64, 128):
        N.load().tedm_conv_set_tile_n(force_bn)
        try:
            s0, s1 = N.conv_igemm(dyn, wd, N.MODE_3X3, c0 + c1, split=c0, residual2=_nhwc(extra.cpu()))
            r0, r1 = N.conv_igemm(dyn, wd, N.MODE_3X3, c0 + c1, split=c0, residual=_nhwc(dx_ref[:, :c0].detach().cpu()))
        finally:
            N.load().tedm_conv_set_tile_n(0)
        torch.cuda.synchronize()
        assert torch.equal(s0, dx0) and torch.equal(s1, dx1), force_bn
        assert _rel(r0.float().permute(0, 3, 1, 2), 2 * dx_ref[:, :c0]) < 6e-3
        assert _rel(r1.float().permute(0, 3, 1, 2), dx_ref[:, c0:]) < 6e-3


@pytest.mark.parametrize("mode,cout,cin", [(0, 128, 64), (1, 64, 192), (2, 128, 64), (3, 64, 128)])
def test_wgrad_to_oihw(mode, cout, cin):
    from tedm_b200 import native as N
    kh = {0: 1, 1: 3, 2: 4, 3: 3}[mode]
    taps = {0: 1, 1: 9, 2: 16, 3: 16}[mode]
    dw = _rand((cout, taps, cin), 1).cuda()
    grad = _rand((cout, cin, kh, kh), 2).cuda()
    base = grad.clone()
    N.wgrad_to_oihw(dw, grad, mode)
    ref = _unfold_mode3(dw) if mode == 3 else dw.reshape(cout, kh, kh, cin).permute(0, 3, 1, 2)
    assert _rel(grad - base, ref) < 1e-6


# ---- GroupNorm + scale/shift + SiLU --------------------------------------------------------------
@pytest.mark.parametrize("B,H,C,with_ss,with_res", [(3, 16, 64, True, False), (2, 32, 128, False, True),
                                                    (2, 8, 512, True, True), (5, 4, 256, True, False),
                                                    (1, 128, 64, True, True)])
def test_gn_silu_bwd(B, H, C, with_ss, with_res):
    from tedm_b200 import native as N
    groups = 8
    cin = 64
    xin = _nhwc(_rand((B, cin, H, H), 1))
    w = _rand((C, cin, 3, 3), 2, (cin * 9) ** -0.5)
    bias = _rand((C,), 3, 0.1).cuda()
    h, part = N.conv_igemm(xin, N.weight_to_krsc(w.cuda()), N.MODE_3X3, C, bias=bias, gn_groups=groups)
    gamma, beta = (1 + _rand((C,), 4, 0.2)).cuda(), _rand((C,), 5, 0.2).cuda()
    ss = _rand((B, 2 * C + 10), 6, 0.3).cuda() if with_ss else None
    res = _nhwc(_rand((B, C, H, H), 7)) if with_res else None
    dy = _nhwc(_rand((B, C, H, H), 8))
    # reference on the same bf16 conv output
    hf = h.float().permute(0, 3, 1, 2).detach().requires_grad_(True)
    g_, b_ = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    ss_ = ss.clone().requires_grad_(True) if with_ss else None
    z = F.group_norm(hf, groups, g_, b_, eps=1e-5)
    if with_ss:
        sc, sh = ss_[:, 5:5 + C], ss_[:, 5 + C:5 + 2 * C]
        z = z * (sc[:, :, None, None] + 1) + sh[:, :, None, None]
    y = F.silu(z)
    if with_res:
        y = y + res.float().permute(0, 3, 1, 2)
    y.backward(dy.float().permute(0, 3, 1, 2))
    dgamma, dbeta, dbias = (torch.zeros(C, device="cuda") for _ in range(3))
    dss = torch.zeros_like(ss) if with_ss else None
    dx = N.gn_silu_bwd(h, dy, part, gamma, beta, groups, dgamma, dbeta, dbias, scale_shift=ss, ss_offset=5, dscale_shift=dss)
    torch.cuda.synchronize()
    assert _rel(dx.float().permute(0, 3, 1, 2), hf.grad) < 1e-2
    assert _rel(dgamma, g_.grad) < 2e-3 and _rel(dbeta, b_.grad) < 2e-3
    assert _rel(dbias, hf.grad.sum(dim=(0, 2, 3))) < 5e-3
    if with_ss:
        assert _rel(dss, ss_.grad) < 2e-3


@pytest.mark.parametrize("C,npix", [(64, 1000), (128, 4096), (256, 77), (512, 512)])
@pytest.mark.parametrize("with_add", [False, True])
def test_layernorm_bwd(C, npix, with_add):
    from tedm_b200 import native as N
    x = _rand((1, npix, 1, C), 1).to(torch.bfloat16).cuda()
    dy = _rand((1, npix, 1, C), 2).to(torch.bfloat16).cuda()
    add = _rand((1, npix, 1, C), 3).to(torch.bfloat16).cuda() if with_add else None
    g = (1 + _rand((C,), 4, 0.2)).cuda()
    xf = x.float().requires_grad_(True)
    g_ = g.clone().requires_grad_(True)
    mean = xf.mean(-1, keepdim=True)
    var = xf.var(-1, unbiased=False, keepdim=True)
    y = (xf - mean) * (var + 1e-5).rsqrt() * g_
    y.backward(dy.float())
    dg = torch.zeros(C, device="cuda")
    dx = N.layernorm_bwd(x, g, dy, dg, eps=1e-5, add=add)
    ref = xf.grad + (add.float() if with_add else 0)
    assert _rel(dx, ref) < 6e-3
    assert _rel(dg, g_.grad) < 2e-3


@pytest.mark.parametrize("C,npix", [(64, 5000), (384, 300), (512, 64)])
def test_bias_grad_and_add(C, npix):
    from tedm_b200 import native as N
    dy = _rand((1, npix, 1, C), 1).to(torch.bfloat16).cuda()
    db = torch.ones(C, device="cuda")
    N.bias_grad(dy, db)
    assert _rel(db - 1, dy.float().sum(dim=(0, 1, 2))) < 1e-5
    other = _rand((1, npix, 1, C), 2).to(torch.bfloat16).cuda()
    s = N.add_bf16(dy, other)
    assert _rel(s, (dy.float() + other.float()).to(torch.bfloat16)) < 1e-6


def test_final_conv_bwd():
    from tedm_b200 import native as N
    B, H, C, od = 3, 32, 64, 1
    h = _nhwc(_rand((B, C, H, H), 1))
    w = _rand((od, C), 2, 0.1).cuda()
    dout = _rand((B, od, H, H), 3).cuda()
    hf = h.float().requires_grad_(True)
    w_ = w.clone().requires_grad_(True)
    out = torch.einsum("bhwc,oc->bohw", hf, w_)
    out.backward(dout)
    dw, db = torch.zeros_like(w), torch.zeros(od, device="cuda")
    dh = N.final_conv1x1_bwd(h, w, dout, dw, db)
    assert _rel(dh, hf.grad) < 6e-3
    assert _rel(dw, w_.grad) < 1e-4
    assert _rel(db, dout.sum(dim=(0, 2, 3))) < 1e-4


@pytest.mark.parametrize("H,cout", [(64, 64), (128, 64), (48, 64), (32, 32)])
def test_stem_wgrad(H, cout):
    from tedm_b200 import native as N
    B = 3
    x = torch.rand((B, 1, H, H), generator=torch.Generator().manual_seed(1)).cuda()
    w = _rand((cout, 1, 7, 7), 2, 0.1).cuda().requires_grad_(True)
    y = F.conv2d(x, w, padding=3)
    dy = _bf(_rand(tuple(y.shape), 3)).cuda()
    (dw_ref,) = torch.autograd.grad(y, w, dy)
    dw, db = torch.zeros_like(w), torch.zeros(cout, device="cuda")
    N.stem_conv7x7_wgrad(x, _nhwc(dy.cpu()), dw, db)
    assert _rel(dw, dw_ref) < 1e-4
    assert _rel(db, dy.sum(dim=(0, 2, 3))) < 1e-4


def test_time_mlp_bwd():
    """time_embed_train + linear_bwd chain vs autograd of the reference's time path (unet_model.py:150-152,287-292)."""
    from tedm_b200 import native as N
    from tedm_b200.models.unet_model import SinusoidalPosEmb
    B, dim, tdim, total = 37, 64, 256, 1000
    w1, b1 = _rand((tdim, dim), 1, 0.125).cuda(), _rand((tdim,), 2, 0.1).cuda()
    w2, b2 = _rand((tdim, tdim), 3, 0.0625).cuda(), _rand((tdim,), 4, 0.1).cuda()
    wc, bc = _rand((total, tdim), 5, 0.0625).cuda(), _rand((total,), 6, 0.1).cuda()
    t = torch.randint(0, 1000, (B,), generator=torch.Generator().manual_seed(7)).cuda()
    freq = SinusoidalPosEmb(dim).frequencies("cpu").float().cuda()
    params = [p.clone().requires_grad_(True) for p in (w1, b1, w2, b2, wc, bc)]
    arg = t.float()[:, None] * freq[None, :]
    emb_ref = torch.cat([arg.sin(), arg.cos()], dim=-1)
    a1 = F.linear(emb_ref, params[0], params[1])
    temb_ref = F.linear(F.gelu(a1), params[2], params[3])
    proj = F.linear(F.silu(temb_ref), params[4], params[5])
    dproj = _rand((B, total), 8).cuda()
    proj.backward(dproj)
    emb, hid, temb = N.time_embed_train(t, freq, w1, b1, w2, b2)
    assert _rel(temb, temb_ref) < 2e-5
    grads = [torch.zeros_like(p) for p in (w1, b1, w2, b2, wc, bc)]
    d3 = N.linear_bwd(dproj, None, N.ACT_NONE, temb, N.ACT_SILU, wc, grads[4], grads[5])
    d2 = N.linear_bwd(d3, temb, N.ACT_SILU, hid, N.ACT_GELU, w2, grads[2], grads[3])
    N.linear_bwd(d2, hid, N.ACT_GELU, emb, N.ACT_NONE, w1, grads[0], grads[1], want_dx=False)
    for g, p in zip(grads, params):
        assert _rel(g, p.grad) < 1e-4, _rel(g, p.grad)


# ---- attention cores -------------------------------------------------------------------------------
@pytest.mark.parametrize("B,H", [(2, 16), (1, 64), (3, 8), (1, 40)])
def test_linear_attention_bwd(B, H):
    from tedm_b200 import native as N
    n, heads, dh = H * H, 4, 32
    qkv = _rand((B, H, H, 3 * heads * dh), 1).to(torch.bfloat16).cuda()
    dout = _rand((B, H, H, heads * dh), 2).to(torch.bfloat16).cuda()
    x = qkv.float().requires_grad_(True)
    q, k, v = (t.reshape(B, n, heads, dh).permute(0, 2, 3, 1) for t in x.reshape(B, n, -1).chunk(3, dim=-1))  # b h d n
    q = q.softmax(dim=-2) * dh ** -0.5
    k = k.softmax(dim=-1)
    v = v / n
    ctx = torch.einsum("bhdn,bhen->bhde", k, v)
    out = torch.einsum("bhde,bhdn->bhen", ctx, q)                     # b h e n
    out = out.permute(0, 3, 1, 2).reshape(B, H, H, heads * dh)
    out.backward(dout.float())
    o, ws = N.linear_attention(qkv, heads, dh, want_workspace=True)
    assert _rel(o, out) < 1.5e-2
    dqkv = N.linear_attention_bwd(qkv, dout, ws, heads, dh)
    torch.cuda.synchronize()
    for name, sl in (("dq", slice(0, 128)), ("dk", slice(128, 256)), ("dv", slice(256, 384))):
        assert _rel(dqkv[..., sl], x.grad[..., sl]) < 1.5e-2, (name, _rel(dqkv[..., sl], x.grad[..., sl]))


@pytest.mark.parametrize("B,H", [(2, 16), (3, 4), (1, 8), (2, 32), (1, 10), (1, 24)])
def test_attention_bwd(B, H):
    from tedm_b200 import native as N
    n, heads, dh, scale = H * H, 4, 32, 16.0
    qkv = _rand((B, H, H, 3 * heads * dh), 1).to(torch.bfloat16).cuda()
    dout = _rand((B, H, H, heads * dh), 2).to(torch.bfloat16).cuda()
    x = qkv.float().requires_grad_(True)
    q, k, v = (t.reshape(B, n, heads, dh).permute(0, 2, 3, 1) for t in x.reshape(B, n, -1).chunk(3, dim=-1))  # b h d n
    q, k = F.normalize(q, dim=-1), F.normalize(k, dim=-1)
    sim = torch.einsum("bhdi,bhdj->bhij", q, k) * scale
    attn = sim.softmax(dim=-1)
    out = torch.einsum("bhij,bhdj->bhid", attn, v)                    # b h n d
    out = out.permute(0, 2, 1, 3).reshape(B, H, H, heads * dh)
    out.backward(dout.float())
    variants = [("flash", N.attention_bwd(qkv, dout, heads, dh, scale, o=N.attention(qkv, heads, dh, scale)))]
    if n <= 256:
        variants.append(("single-CTA", N.attention_bwd(qkv, dout, heads, dh, scale)))
    torch.cuda.synchronize()
    for tag, dqkv in variants:
        for name, sl in (("dq", slice(0, 128)), ("dk", slice(128, 256)), ("dv", slice(256, 384))):
            assert _rel(dqkv[..., sl], x.grad[..., sl]) < 1e-2, (tag, name, _rel(dqkv[..., sl], x.grad[..., sl]))


def test_adam_step_matches_torch():
    from tedm_b200 import native as N
    n = 4096 + 64
    p0, g = _rand((n,), 1).cuda(), _rand((n,), 2).cuda()
    ref = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=1e-3, betas=(0.9, 0.999), eps=1e-8)
    p, m, v = p0.clone(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    for step in range(1, 4):
        ref.grad = g * step
        opt.step()
        N.adam_step(p, g * step, m, v, 1e-3, 0.9, 0.999, 1e-8, 0.0, step)
    assert _rel(p, ref.detach()) < 1e-6
