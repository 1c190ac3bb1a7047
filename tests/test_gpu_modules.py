"""The UNet's submodules are callable on their own (VERDICT r1 missing-2/6): code that walks the module tree
the way the reference's contrastive encoders do (models/global_local_cl.py:32-50, 74-107), forward hooks on
`ups[i][2]` (models/datasetDM_model.py:16-27, 50-53), and the learned sinusoidal embedding (models/unet_model.py:96-114)."""
import numpy as np
import pytest
import torch

from oracle import tedm_oracle as O
from tests.golden.synth import synth_images, synth_state_dict, synth_timesteps

pytestmark = pytest.mark.gpu
TOL = 2e-2


def _rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _unet(mults=(1, 2, 4, 8)):
    from tedm_b200.models import Unet
    sd = synth_state_dict(O.unet_param_shapes(dim_mults=mults), 0)
    m = Unet(64, dim_mults=mults).eval()
    m.load_state_dict(sd)
    return m.cuda(), sd


def _walk(m, x, t):
    """Unet.forward re-expressed as calls of the submodules, in the style of a subclass that walks the tree."""
    temb = m.time_mlp(t) if t is not None else None
    x = m.init_conv(x)
    stem, skips = x, []
    for b1, b2, attn, down in m.downs:
        x = b1(x, temb)
        skips.append(x)
        x = attn(b2(x, temb))
        skips.append(x)
        x = down(x)
    x = m.mid_block2(m.mid_attn(m.mid_block1(x, temb)), temb)
    for b1, b2, attn, up in m.ups:
        x = b1(torch.cat((x, skips.pop()), dim=1), temb)
        x = attn(b2(torch.cat((x, skips.pop()), dim=1), temb))
        x = up(x)
    return m.final_conv(m.final_res_block(torch.cat((x, stem), dim=1), temb))


@pytest.mark.parametrize("with_t", [True, False])
def test_walking_the_submodules_equals_unet_forward(with_t):
    m, sd = _unet()
    x = synth_images(2, 32, 4).cuda()
    t = synth_timesteps(2, seed=4).cuda() if with_t else None
    with torch.no_grad():
        walked = _walk(m, x, t)
        fused = m(x, t)
    ref = O.unet_forward(sd, x.cpu(), t.cpu() if with_t else None)
    print("walk vs oracle", _rel(walked, ref), "engine vs oracle", _rel(fused, ref), "walk vs engine", _rel(walked, fused))
    assert walked.shape == ref.shape and _rel(walked, ref) < TOL and _rel(walked, fused) < TOL


def test_groupnorm_fused_into_conv_is_bit_identical_through_the_unet():
    """engine.fuse_gn_into_conv (Block's GroupNorm + SiLU applied to the next conv's input in shared memory) and
    engine.fuse_res_conv (ResnetBlock's tail in the res_conv epilogue) against the plain schedule: the first is bit-identical,
    the second rounds once less."""
    m, sd = _unet()
    x = synth_images(3, 64, 5).cuda()
    t = synth_timesteps(3, seed=5).cuda()
    eng = m.engine
    saved = (eng.fuse_gn_into_conv, eng.fuse_res_conv)
    try:
        outs = {}
        for gn, res in ((False, False), (True, False), (False, True), (True, True)):
            eng.fuse_gn_into_conv, eng.fuse_res_conv = gn, res
            with torch.no_grad():
                outs[(gn, res)] = m(x, t)
    finally:
        eng.fuse_gn_into_conv, eng.fuse_res_conv = saved
    assert torch.equal(outs[(True, False)], outs[(False, False)])
    assert torch.equal(outs[(True, True)], outs[(False, True)])
    ref = O.unet_forward(sd, x.cpu(), t.cpu())
    assert _rel(outs[(False, True)], outs[(False, False)]) < TOL      # two equally valid bf16 roundings, 0.96 % apart at the output
    assert _rel(outs[(True, True)], ref) < TOL and _rel(outs[(False, False)], ref) < TOL


def test_each_submodule_against_the_oracle():
    m, sd = _unet()
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 64, 32, 32, generator=g)
    temb = torch.randn(2, 256, generator=g)
    ident = lambda z: z
    with torch.no_grad():
        b1 = m.downs[0][0]
        assert _rel(b1(x.cuda(), temb.cuda()), O._resblock(sd, "downs.0.0.", x, temb, 8, ident)) < TOL
        assert _rel(b1(x.cuda()), O._resblock(sd, "downs.0.0.", x, None, 8, ident)) < TOL
        ss = (torch.randn(2, 64, 1, 1, generator=g) * 0.1, torch.randn(2, 64, 1, 1, generator=g) * 0.1)
        assert _rel(b1.block1(x.cuda(), (ss[0].cuda(), ss[1].cuda())), O._block(sd, "downs.0.0.block1.", x, 8, ss, ident)) < TOL
        assert _rel(m.downs[0][2](x.cuda()), O._linear_attention(sd, "downs.0.2.", x, 1e-5, ident)) < TOL
        x2 = torch.randn(2, 128, 16, 16, generator=g)                      # a two-source decoder block: 64 + 64 -> 64
        assert _rel(m.ups[3][0](x2.cuda(), temb.cuda()), O._resblock(sd, "ups.3.0.", x2, temb, 8, ident)) < TOL
        xm = torch.randn(2, 512, 16, 16, generator=g)
        assert _rel(m.mid_attn(xm.cuda()), O._mid_attention(sd, "mid_attn.", xm, 1e-5, ident)) < TOL
        ln = m.downs[0][2].fn.norm
        assert _rel(ln(x.cuda()), O._chan_layernorm(x, sd["downs.0.2.fn.norm.g"], 1e-5)) < TOL
        down = m.downs[0][3]
        assert _rel(down(x.cuda()), torch.nn.functional.conv2d(x, sd["downs.0.3.weight"], sd["downs.0.3.bias"], stride=2, padding=1)) < TOL
        img = synth_images(2, 32, 1)
        assert _rel(m.init_conv(img.cuda()), torch.nn.functional.conv2d(img, sd["init_conv.weight"], sd["init_conv.bias"], padding=3)) < TOL
    with pytest.raises(RuntimeError, match="CUDA"):
        b1(x)                                                                # no CPU fallback
    with pytest.raises(RuntimeError, match="forward-only"):
        b1(x.cuda().requires_grad_(True))


def test_forward_hooks_on_decoder_maps_fire_with_reference_layout():
    """The reference's DatasetDM registers `save_activations` hooks on ups[i][2]; the same registration against this
    Unet sees NCHW fp32 maps equal to the oracle's features."""
    m, sd = _unet()
    got = {}
    hooks = [a.register_forward_hook(lambda mod, inp, out, i=i: got.__setitem__(i, (inp[0].shape, out.detach().cpu())))
             for i, (_, _, a, _) in enumerate(m.ups)]
    x, t = synth_images(2, 32, 2), synth_timesteps(2, seed=2)
    with torch.no_grad():
        m(x.cuda(), t.cuda())
    for h in hooks:
        h.remove()
    _, feats = O.unet_forward(sd, x, t, want_features=True)
    assert sorted(got) == [0, 1, 2, 3]
    for i, f in enumerate(feats):
        shape_in, out = got[i]
        assert out.dtype == torch.float32 and out.shape == f.shape and tuple(shape_in) == tuple(f.shape)
        assert _rel(out, f) < TOL, i
    got.clear()
    with torch.no_grad():
        m(x.cuda(), t.cuda())
    assert not got                                                           # removed hooks stay removed


def test_learned_sinusoidal_embedding_matches_reference_arithmetic():
    """`learned_sinusoidal_cond=True` (unet_model.py:96-114, 279-285): [t, sin(2 pi t w), cos(2 pi t w)] -> time MLP."""
    from tedm_b200.models import Unet
    torch.manual_seed(3)
    m = Unet(64, dim_mults=(1, 2), learned_sinusoidal_cond=True, learned_sinusoidal_dim=16).eval().cuda()
    assert m.time_mlp[1].in_features == 17 and m.time_mlp[0].weights.shape == (8,)
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    x, t = synth_images(2, 32, 6), torch.tensor([3, 700])
    with torch.no_grad():
        got = m(x.cuda(), t.cuda())
    # oracle with the time embedding swapped for the learned one
    w = sd["time_mlp.0.weights"]
    ang = t[:, None].float() * w[None, :] * 2 * np.pi
    four = torch.cat((t[:, None].float(), ang.sin(), ang.cos()), dim=-1)
    e = torch.nn.functional.linear(four, sd["time_mlp.1.weight"], sd["time_mlp.1.bias"])
    temb = torch.nn.functional.linear(torch.nn.functional.gelu(e), sd["time_mlp.3.weight"], sd["time_mlp.3.bias"])
    orig = O.time_embedding
    O.time_embedding = lambda *_a, **_k: temb
    try:
        ref = O.unet_forward({k: v for k, v in sd.items() if k != "time_mlp.0.weights"}, x, t)
    finally:
        O.time_embedding = orig
    assert _rel(got, ref) < TOL
