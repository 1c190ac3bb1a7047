"""tcgen05 implicit-GEMM convolution against F.conv2d on the same bf16-rounded operands."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _rand(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale)


def _nhwc(x):  # NCHW fp32 -> NHWC bf16 cuda
    return x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).cuda()


def _ref_conv(xs, w, b, mode):
    x = torch.cat([t.to(torch.bfloat16).float() for t in xs], dim=1).cuda()
    wq = w.to(torch.bfloat16).float().cuda()
    bb = b.cuda() if b is not None else None
    if mode == 0:
        return F.conv2d(x, wq, bb)
    if mode == 1:
        return F.conv2d(x, wq, bb, padding=1)
    if mode == 2:
        return F.conv2d(x, wq, bb, stride=2, padding=1)
    return F.conv2d(F.interpolate(x, scale_factor=2, mode="nearest"), wq, bb, padding=1)


CASES = [
    # (B, H, W, c0, c1, cout, mode, bias, residual, gn, force_bn)
    (2, 16, 16, 64, 0, 64, 0, True, False, 0, 0),
    (2, 16, 16, 64, 0, 64, 1, True, False, 8, 0),
    (1, 128, 128, 64, 0, 64, 1, True, False, 8, 0),
    (3, 32, 32, 128, 64, 128, 1, True, False, 8, 0),
    (2, 8, 8, 256, 128, 256, 1, True, False, 8, 0),
    (5, 4, 4, 512, 256, 512, 1, True, False, 8, 0),
    (5, 4, 4, 512, 0, 512, 1, True, False, 8, 128),
    (5, 4, 4, 512, 0, 512, 1, True, False, 8, 256),
    (2, 64, 64, 64, 0, 128, 2, True, False, 0, 0),
    (3, 8, 8, 128, 0, 256, 2, True, False, 0, 0),
    (2, 16, 16, 128, 0, 64, 3, True, False, 0, 0),
    (1, 64, 64, 128, 0, 64, 3, True, False, 0, 0),
    (2, 16, 16, 128, 0, 512, 0, True, True, 0, 0),
    (2, 32, 32, 64, 0, 384, 0, False, False, 0, 0),
    (2, 32, 32, 192, 64, 128, 0, True, False, 0, 0),
    (9, 8, 8, 64, 0, 64, 1, True, False, 8, 0),
    (2, 16, 16, 64, 0, 256, 1, True, False, 8, 256),
    (2, 16, 16, 64, 0, 128, 1, True, False, 8, 128),
    # weight-stationary row path (3x3, 128-pixel rows, cout 64) incl. two sources, 256-wide rows, many tiles per CTA
    (3, 128, 128, 64, 0, 64, 1, True, False, 8, 0),
    (1, 128, 128, 64, 64, 64, 1, True, False, 8, 0),
    (2, 4, 256, 64, 0, 64, 1, True, False, 8, 0),
    (1, 128, 128, 128, 0, 64, 1, False, False, 0, 0),
    # four-row tiles on 64-pixel rows (two images per tile; odd batch; 8 rows)
    (3, 64, 64, 64, 0, 64, 1, True, False, 8, 0),
    (2, 8, 64, 64, 0, 64, 1, False, False, 0, 0),
    # persistent loop with more tiles than SMs on the generic path
    (40, 16, 16, 128, 0, 128, 1, True, False, 8, 64),
    (6, 64, 64, 64, 0, 384, 0, False, False, 0, 0),
    # residual epilogue (prefetched rows) on every N tile width and mode; one GroupNorm group spanning both column halves
    (2, 16, 16, 64, 0, 64, 1, True, True, 0, 0),
    (2, 16, 16, 128, 0, 256, 1, True, True, 0, 256),
    (3, 16, 16, 128, 64, 128, 1, False, True, 0, 128),
    (1, 128, 128, 64, 0, 64, 1, False, True, 0, 0),
    (2, 16, 16, 128, 0, 64, 3, True, True, 0, 0),
    (2, 32, 32, 64, 0, 128, 2, True, True, 0, 0),
    (5, 4, 4, 512, 0, 512, 1, True, False, 8, 64),
]


@pytest.mark.parametrize("case", CASES, ids=[str(c) for c in CASES])
def test_conv_igemm(case):
    from tedm_b200 import native as N
    B, H, W, c0, c1, cout, mode, use_bias, use_res, gn, force_bn = case
    kh = {0: 1, 1: 3, 2: 4, 3: 3}[mode]
    x0 = _rand((B, c0, H, W), 1)
    x1 = _rand((B, c1, H, W), 2) if c1 else None
    cin = c0 + c1
    w = _rand((cout, cin, kh, kh), 3, scale=(cin * kh * kh) ** -0.5)
    b = _rand((cout,), 4, 0.1) if use_bias else None
    ref = _ref_conv([x0] + ([x1] if c1 else []), w, b, mode)
    res = None
    if use_res:
        res = _rand(tuple(ref.shape), 5)
        ref = ref + res.to(torch.bfloat16).float().cuda()
    wk = N.fold_upsample_weight(w.cuda()) if mode == 3 else N.weight_to_krsc(w.cuda())
    N.load().tedm_conv_set_tile_n(force_bn)
    try:
        got = N.conv_igemm(_nhwc(x0), wk, mode, cout, bias=b.cuda() if b is not None else None,
                           src1=_nhwc(x1) if c1 else None, residual=_nhwc(res) if use_res else None, gn_groups=gn)
    finally:
        N.load().tedm_conv_set_tile_n(0)
    torch.cuda.synchronize()
    if gn:
        got, part = got
        stats = part.double().sum(dim=1)                     # (B, groups, 2)
        r = ref.double().reshape(B, gn, -1)
        assert _rel(stats[..., 0], r.sum(-1)) < 2e-3 or (stats[..., 0] - r.sum(-1)).abs().max() < 0.5
        assert _rel(stats[..., 1], (r * r).sum(-1)) < 1e-3
    out = got.float().permute(0, 3, 1, 2)
    tol = 8e-3 if mode != 3 else 1.2e-2   # mode 3 sums taps in fp32 before the bf16 rounding of the weights
    assert _rel(out, ref) < tol, _rel(out, ref)


def test_conv_batch_strided_views():
    from tedm_b200 import native as N
    S, B, H, W, C, cout = 3, 2, 8, 8, 64, 128
    x = _rand((B * S, C, H, W), 7)
    xs = _nhwc(x)
    out = torch.zeros(B * S, H, W, cout, dtype=torch.bfloat16, device="cuda")
    ws = [_rand((cout, C, 1, 1), 10 + s, C ** -0.5) for s in range(S)]
    for s in range(S):
        N.conv_igemm(xs[s::S], N.weight_to_krsc(ws[s].cuda()), 0, cout, out=out[s::S])
    torch.cuda.synchronize()
    for s in range(S):
        ref = _ref_conv([x[s::S]], ws[s], None, 0)
        assert _rel(out[s::S].float().permute(0, 3, 1, 2), ref) < 8e-3


def test_conv_rejects_unsupported_shapes():
    from tedm_b200 import native as N
    x = torch.zeros(1, 8, 8, 48, dtype=torch.bfloat16, device="cuda")
    w = torch.zeros(64, 3, 3, 48, dtype=torch.bfloat16, device="cuda")
    with pytest.raises(RuntimeError, match="multiples of 64"):
        N.conv_igemm(x, w, 1, 64)


def test_conv_ws_path_matches_generic_path():
    from tedm_b200 import native as N
    x, w, b = _rand((2, 64, 128, 128), 1), _rand((64, 64, 3, 3), 2, 576 ** -0.5), _rand((64,), 3, 0.1)
    xh, wk = _nhwc(x), N.weight_to_krsc(w.cuda())
    y_ws, p_ws = N.conv_igemm(xh, wk, 1, 64, bias=b.cuda(), gn_groups=8)
    N.load().tedm_conv_set_ws(0)
    try:
        y_g, p_g = N.conv_igemm(xh, wk, 1, 64, bias=b.cuda(), gn_groups=8)
    finally:
        N.load().tedm_conv_set_ws(1)
    torch.cuda.synchronize()
    assert _rel(y_ws.float(), y_g.float()) < 2e-3          # same products, different accumulation order
    assert _rel(p_ws.sum(1), p_g.sum(1)) < 1e-4


@pytest.mark.parametrize("case", [
    # (B, H, W, c0, c1, cout, bias, gn, residual, force_bn)
    (3, 16, 16, 64, 0, 128, True, 8, False, 0), (2, 64, 64, 128, 64, 128, True, 8, False, 0), (5, 16, 16, 512, 256, 512, True, 8, False, 256),
    (2, 32, 32, 256, 0, 256, False, 0, True, 0), (1, 128, 128, 64, 0, 128, True, 0, True, 0), (40, 16, 32, 128, 0, 128, True, 8, False, 64),
    (2, 16, 8, 192, 0, 64, True, 8, False, 0), (3, 32, 64, 64, 0, 320, False, 0, False, 0)])
def test_conv_halo_tiles_match_per_tap_tiles(case):
    """3x3 convs on 8 x 16-pixel tiles whose 10 x 18 halo box is loaded once per 64-channel block (nine windows of it feed the
    tensor core, rows 1280 B apart) against one 128-pixel box per tap (tedm_conv_set_halo(0)) and F.conv2d: image borders,
    two sources, every N tile, GroupNorm partials, residual epilogue, more tiles than SMs, CTA pairs and single CTAs."""
    from tedm_b200 import native as N
    B, H, W, c0, c1, cout, use_bias, gn, use_res, force_bn = case
    ctot = c0 + c1
    x0, x1 = _rand((B, c0, H, W), 41), (_rand((B, c1, H, W), 42) if c1 else None)
    w = _rand((cout, ctot, 3, 3), 43, (9 * ctot) ** -0.5)
    b = _rand((cout,), 44, 0.1).cuda() if use_bias else None
    res = _nhwc(_rand((B, cout, H, W), 45)) if use_res else None
    wk = N.weight_to_krsc(w.cuda())
    outs = []
    try:
        N.load().tedm_conv_set_tile_n(force_bn)
        for halo, pairs in ((1, 1), (0, 1), (1, 0)):
            N.load().tedm_conv_set_halo(halo)
            N.set_cta_pairs(pairs)
            outs.append(N.conv_igemm(_nhwc(x0), wk, 1, cout, bias=b, src1=_nhwc(x1) if c1 else None, gn_groups=gn, residual=res))
    finally:
        N.load().tedm_conv_set_halo(1)
        N.set_cta_pairs(1)
        N.load().tedm_conv_set_tile_n(0)
    torch.cuda.synchronize()
    if gn:
        assert _rel(outs[0][1].sum(1), outs[1][1].sum(1)) < 1e-5
        assert torch.equal(outs[0][1], outs[2][1])               # pairs and single CTAs add in the same order
        outs = [o[0] for o in outs]
    assert torch.equal(outs[0], outs[2])
    assert _rel(outs[0].float(), outs[1].float()) < 2e-3         # same products, different accumulation order
    ref = _ref_conv([x0] + ([x1] if c1 else []), w, b.cpu() if use_bias else None, 1)
    if use_res:
        ref = ref + res.float().permute(0, 3, 1, 2)
    assert _rel(outs[0].float().permute(0, 3, 1, 2), ref) < 6e-3


@pytest.mark.parametrize("case", [(2, 64, 64, 128, 128, True, 0), (3, 16, 16, 512, 512, True, 0), (2, 32, 32, 256, 256, False, 0),
                                  (1, 16, 16, 64, 128, True, 64), (40, 16, 16, 128, 128, True, 0), (2, 128, 128, 128, 128, False, 0),
                                  (3, 128, 128, 64, 64, True, 0), (150, 4, 128, 64, 64, True, 0), (2, 8, 256, 64, 64, False, 0)])
def test_conv_input_groupnorm_fused_into_halo_boxes(case):
    """Block.forward's GroupNorm + (scale + 1) / shift + SiLU (reference models/unet_model.py:128-134) applied to the conv's
    halo boxes in shared memory (src0_affine; the input rows of the four-row weight-stationary kernel for 64 -> 64 on 128-pixel
    rows) against the separate pass followed by the same conv: the normalised values are rounded to bf16 by the same expression
    in both, so the outputs are bit-identical; pairs and single CTAs."""
    from tedm_b200 import native as N
    B, H, W, c, cout, with_ss, force_bn = case
    groups = 8
    assert N.conv_src_affine_supported(H, W, c, cout)
    xin = _rand((B, c, H, W), 51)
    w1, b1 = _rand((c, c, 3, 3), 52, (9 * c) ** -0.5), _rand((c,), 53, 0.1)
    w2, b2 = _rand((cout, c, 3, 3), 54, (9 * c) ** -0.5), _rand((cout,), 55, 0.1)
    gamma, beta = (1.0 + _rand((c,), 56, 0.2)).cuda(), _rand((c,), 57, 0.2).cuda()
    ss = _rand((B, 2 * c + 6), 58, 0.3).cuda() if with_ss else None
    h1, part = N.conv_igemm(_nhwc(xin), N.weight_to_krsc(w1.cuda()), 1, c, bias=b1.cuda(), gn_groups=groups)
    a1 = N.gn_silu(h1, part, gamma, beta, groups, scale_shift=ss, ss_offset=3)
    aff = N.gn_affine(part, gamma, beta, groups, H * W, scale_shift=ss, ss_offset=3)
    w2k = N.weight_to_krsc(w2.cuda())
    outs = []
    try:
        N.load().tedm_conv_set_tile_n(force_bn)
        for pairs in (1, 0):
            N.set_cta_pairs(pairs)
            two = N.conv_igemm(a1, w2k, 1, cout, bias=b2.cuda(), gn_groups=groups)
            one = N.conv_igemm(h1, w2k, 1, cout, bias=b2.cuda(), gn_groups=groups, src0_affine=aff)
            outs.append((two, one))
    finally:
        N.set_cta_pairs(1)
        N.load().tedm_conv_set_tile_n(0)
    torch.cuda.synchronize()
    for two, one in outs:
        assert _rel(one[0].float(), two[0].float()) < 1e-6, _rel(one[0].float(), two[0].float())
        assert torch.equal(one[0], two[0]) and torch.equal(one[1], two[1])
    hf = h1.float().permute(0, 3, 1, 2)
    y = F.group_norm(hf, groups, gamma, beta, 1e-5)
    if with_ss:
        y = y * (ss[:, 3:3 + c, None, None] + 1) + ss[:, 3 + c:3 + 2 * c, None, None]
    ref = F.conv2d(F.silu(y).to(torch.bfloat16).float(), w2.to(torch.bfloat16).float().cuda(), b2.cuda(), padding=1)
    assert _rel(outs[0][1][0].float().permute(0, 3, 1, 2), ref) < 8e-3


def test_conv_src_affine_is_refused_off_the_halo_path():
    from tedm_b200 import native as N
    assert not N.conv_src_affine_supported(64, 64, 64, 64) and not N.conv_src_affine_supported(8, 8, 256, 256)
    assert not N.conv_src_affine_supported(128, 128, 128, 64)
    x = torch.zeros(2, 8, 8, 64, dtype=torch.bfloat16, device="cuda")
    w = torch.zeros(64, 3, 3, 64, dtype=torch.bfloat16, device="cuda")
    with pytest.raises(RuntimeError, match="src0_affine"):
        N.conv_igemm(x, w, 1, 64, src0_affine=torch.zeros(2, 64, 2, device="cuda"))


@pytest.mark.parametrize("case", [(2, 16, 16, 128, 64, False), (1, 64, 64, 128, 64, False), (3, 32, 32, 256, 128, False), (2, 16, 16, 64, 64, True),
                                  (20, 16, 8, 128, 256, False), (5, 32, 16, 64, 128, True)])
def test_upsample_conv_halo_tiles_match_per_tap_tiles(case):
    """The folded nearest-x2 upsample conv (mode 3) in halo mode: the four parity tiles of an M tile share the 10 x 18 boxes of
    all channel blocks (2 x 2 windows each) against one box per (parity, tap); pairs and single CTAs; residual epilogue."""
    from tedm_b200 import native as N
    B, H, W, c, cout, use_res = case
    x = _rand((B, c, H, W), 61)
    w, b = _rand((cout, c, 3, 3), 62, (9 * c) ** -0.5), _rand((cout,), 63, 0.1)
    res = _nhwc(_rand((B, cout, 2 * H, 2 * W), 64)) if use_res else None
    wk = N.fold_upsample_weight(w.cuda())
    outs = []
    try:
        for halo, pairs in ((3, 1), (1, 1), (3, 0)):       # 3 = halo tiles for mode 3 too (off by default: measured slower)
            N.load().tedm_conv_set_halo(halo)
            N.set_cta_pairs(pairs)
            outs.append(N.conv_igemm(_nhwc(x), wk, 3, cout, bias=b.cuda(), residual=res))
    finally:
        N.load().tedm_conv_set_halo(1)
        N.set_cta_pairs(1)
    torch.cuda.synchronize()
    assert torch.equal(outs[0], outs[2])
    assert _rel(outs[0].float(), outs[1].float()) < 2e-3
    ref = _ref_conv([x], w, b, 3)
    if use_res:
        ref = ref + res.float().permute(0, 3, 1, 2)
    assert _rel(outs[0].float().permute(0, 3, 1, 2), ref) < 1.2e-2


def test_cta_pairs_with_n256_tiles():
    """CTA pairs on the widest N tile (256 columns: each CTA stages 128 weight rows): bit-identical to single CTAs."""
    from tedm_b200 import native as N
    x, w, b = _rand((6, 256, 16, 16), 31), _rand((512, 256, 3, 3), 32, 2304 ** -0.5), _rand((512,), 33, 0.1)
    xh, wk = _nhwc(x), N.weight_to_krsc(w.cuda())
    outs = []
    try:
        N.load().tedm_conv_set_tile_n(256)
        for pairs in (0, 2):
            N.set_cta_pairs(pairs)
            outs.append(N.conv_igemm(xh, wk, 1, 512, bias=b.cuda(), gn_groups=8))
    finally:
        N.set_cta_pairs(1)
        N.load().tedm_conv_set_tile_n(0)
    torch.cuda.synchronize()
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    assert _rel(outs[1][0].float().permute(0, 3, 1, 2), _ref_conv([x], w, b, 1)) < 6e-3


@pytest.mark.parametrize("case", [(3, 128, 128, 64, 0, True, 8), (2, 8, 256, 64, 0, False, 0), (150, 4, 128, 64, 0, True, 8),
                                  (1, 128, 128, 64, 0, True, 0), (2, 128, 128, 64, 64, True, 8), (2, 16, 128, 128, 0, False, 0),
                                  (75, 4, 128, 64, 64, True, 8), (5, 64, 64, 64, 0, True, 8), (2, 16, 64, 64, 0, False, 0),
                                  (301, 4, 64, 64, 0, True, 8)])
def test_conv_ws4_tiles_match_single_row_tiles(case):
    """3x3 into 64 channels with resident weights and four output rows per tile (every input row multiplied by the three
    vertical taps at once, N = 192) against the one-row / generic tiles (tedm_conv_set_ws(2)) and against F.conv2d: 64 and
    128 input channels, one or two sources, 64-pixel rows (two images per tile, odd batches), image borders, more tiles
    than SMs, two x tiles."""
    from tedm_b200 import native as N
    B, H, W, c0, c1, use_bias, gn = case
    ctot = c0 + c1
    x0, x1 = _rand((B, c0, H, W), 11), (_rand((B, c1, H, W), 14) if c1 else None)
    w = _rand((64, ctot, 3, 3), 12, (9 * ctot) ** -0.5)
    b = _rand((64,), 13, 0.1).cuda() if use_bias else None
    wk = N.weight_to_krsc(w.cuda())
    outs = []
    try:
        for mode in (1, 2):
            N.load().tedm_conv_set_ws(mode)
            outs.append(N.conv_igemm(_nhwc(x0), wk, 1, 64, bias=b, src1=_nhwc(x1) if c1 else None, gn_groups=gn))
    finally:
        N.load().tedm_conv_set_ws(1)
    torch.cuda.synchronize()
    four, one = outs
    if gn:
        assert four[1].shape == one[1].shape
        assert _rel(four[1].sum(1), one[1].sum(1)) < 1e-5
        if W >= 128:
            assert _rel(four[1], one[1]) < 1e-4                  # the partials keep their layout (one per 128-pixel row)
        four, one = four[0], one[0]
    assert _rel(four.float(), one.float()) < 2e-3                # same products, different accumulation order
    ref = _ref_conv([x0] + ([x1] if c1 else []), w, b.cpu() if use_bias else None, 1)
    assert _rel(four.float().permute(0, 3, 1, 2), ref) < 6e-3


@pytest.mark.parametrize("case", [(2, 128, 128, 64, 64, 64), (3, 16, 16, 256, 0, 512), (2, 32, 32, 128, 0, 256), (5, 64, 64, 128, 64, 128)])
def test_conv_residual_affine_is_resnet_block_tail(case):
    """ResnetBlock's tail `SiLU(GroupNorm(h2)) + res_conv(x)` (reference models/unet_model.py:174-175) in ONE kernel -- the 1x1
    conv reads the raw output of block2's conv as a residual that is normalised in its epilogue -- against the two passes
    (1x1 conv, then GroupNorm + SiLU + add) and against torch."""
    from tedm_b200 import native as N
    B, H, W, c0, c1, cout = case
    groups = 8
    x0, x1 = _rand((B, c0, H, W), 21), (_rand((B, c1, H, W), 22) if c1 else None)
    hin = _rand((B, cout, H, W), 23)
    w3, b3 = _rand((cout, cout, 3, 3), 24, (9 * cout) ** -0.5), _rand((cout,), 25, 0.1)
    wr, br = _rand((cout, c0 + c1, 1, 1), 26, (c0 + c1) ** -0.5), _rand((cout,), 27, 0.1)
    gamma, beta = (1.0 + _rand((cout,), 28, 0.2)).cuda(), _rand((cout,), 29, 0.2).cuda()
    h2, part = N.conv_igemm(_nhwc(hin), N.weight_to_krsc(w3.cuda()), 1, cout, bias=b3.cuda(), gn_groups=groups)
    wrk = N.weight_to_krsc(wr.cuda())
    s1 = _nhwc(x1) if c1 else None
    res = N.conv_igemm(_nhwc(x0), wrk, 0, cout, bias=br.cuda(), src1=s1)
    two = N.gn_silu(h2, part, gamma, beta, groups, residual=res)
    aff = N.gn_affine(part, gamma, beta, groups, H * W)
    one = N.conv_igemm(_nhwc(x0), wrk, 0, cout, bias=br.cuda(), src1=s1, residual=h2, residual_affine=aff)
    torch.cuda.synchronize()
    assert _rel(one.float(), two.float()) < 4e-3            # one bf16 rounding fewer than the two-pass form
    hf = h2.float().permute(0, 3, 1, 2)
    ref = F.silu(F.group_norm(hf, groups, gamma, beta, 1e-5)) + _ref_conv([x0] + ([x1] if c1 else []), wr, br, 0)
    assert _rel(one.float().permute(0, 3, 1, 2), ref) < 5e-3
    assert _rel(one.float().permute(0, 3, 1, 2), ref) <= _rel(two.float().permute(0, 3, 1, 2), ref) * 1.05


def test_conv_residual_affine_needs_tiles_inside_one_image():
    from tedm_b200 import native as N
    x = torch.zeros(4, 8, 8, 64, dtype=torch.bfloat16, device="cuda")
    w = torch.zeros(64, 1, 1, 64, dtype=torch.bfloat16, device="cuda")
    aff = torch.zeros(4, 64, 2, device="cuda")
    with pytest.raises(RuntimeError, match="inside one image"):
        N.conv_igemm(x, w, 0, 64, residual=x, residual_affine=aff)


def test_conv_fp32_output():
    from tedm_b200 import native as N
    x, w = _rand((3, 128, 16, 16), 1), _rand((128, 128, 1, 1), 2, 128 ** -0.5)
    got = N.conv_igemm(_nhwc(x), N.weight_to_krsc(w.cuda()), 0, 128, out_dtype=torch.float32)
    assert got.dtype == torch.float32
    ref = _ref_conv([x], w, None, 0)
    assert _rel(got.permute(0, 3, 1, 2), ref) < 1e-5


@pytest.mark.parametrize("case", [
    # (B, H, W, c0, c1, cout, mode, gn)
    (2, 16, 16, 64, 0, 64, 0, 0), (2, 16, 16, 128, 64, 128, 1, 8), (2, 128, 128, 64, 64, 64, 1, 8), (4, 128, 128, 64, 0, 64, 1, 8),
    (2, 64, 64, 64, 0, 128, 2, 0), (2, 16, 16, 128, 0, 64, 3, 0), (6, 8, 8, 256, 128, 128, 1, 8), (3, 16, 16, 192, 0, 64, 1, 0)])
def test_cta_pairs_equal_single_cta(case):
    """CTA pairs (tcgen05 cta_group::2: one MMA over two SMs, M = 256, each CTA staging its own pixels and half of the weight
    tile) accumulate every output in the same K order as one CTA per tile: outputs and GroupNorm partials are bit-identical.
    Forced on (mode 2) for shapes the default heuristic would leave on single CTAs too (1x1, 64 input channels, WS rows)."""
    from tedm_b200 import native as N
    B, H, W, c0, c1, cout, mode, gn = case
    k = {0: 1, 1: 3, 2: 4, 3: 3}[mode]
    ctot = c0 + c1
    x0, x1 = _rand((B, c0, H, W), 1), (_rand((B, c1, H, W), 2) if c1 else None)
    w, b = _rand((cout, ctot, k, k), 3, (ctot * k * k) ** -0.5), _rand((cout,), 4, 0.1)
    wk = N.fold_upsample_weight(w.cuda()) if mode == 3 else N.weight_to_krsc(w.cuda())
    outs = []
    try:
        N.load().tedm_conv_set_ws(2)          # single-row weight-stationary tiles (the four-row kernel is never paired)
        for pairs in (0, 2):
            N.set_cta_pairs(pairs)
            outs.append(N.conv_igemm(_nhwc(x0), wk.reshape(-1), mode, cout, bias=b.cuda(), src1=_nhwc(x1) if c1 else None, gn_groups=gn))
    finally:
        N.set_cta_pairs(1)
        N.load().tedm_conv_set_ws(1)
    torch.cuda.synchronize()
    single, paired = outs
    if gn:
        assert torch.equal(single[1], paired[1])
        single, paired = single[0], paired[0]
    assert torch.equal(single, paired)
    ref = _ref_conv([x0] + ([x1] if c1 else []), w, b, mode)
    assert _rel(paired.float().permute(0, 3, 1, 2), ref) < 6e-3
