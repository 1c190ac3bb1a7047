"""End-to-end parity of the CUDA path against the committed outputs of the live reference
(tests/golden/*.npz, made by tests/golden/make_golden.py) and against the oracle.

Tolerances (BASELINE.json north_star): schedule / q_sample / indices bit-exact; UNet activations and
features <= 2e-2 relative (bf16 compute); argmax masks agree on >= 99.9 % of pixels."""
from argparse import Namespace

import numpy as np
import pytest
import torch

from oracle import tedm_oracle as O
from tests.golden.synth import synth_noise, synth_state_dict

pytestmark = pytest.mark.gpu
T = lambda a: torch.from_numpy(np.asarray(a))
TOL = 2e-2


def _rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _nchw(f):
    return f.float().permute(0, 3, 1, 2).cpu()


@pytest.fixture(scope="module")
def ddpm():
    from tedm_b200.models import DiffusionModel
    m = DiffusionModel(Namespace(normalize=True)).eval()
    sd = synth_state_dict(O.unet_param_shapes(prefix="model."), 0)
    missing = m.load_state_dict(sd, strict=False)
    assert not missing.unexpected_keys and all(not k.startswith("model.") for k in missing.missing_keys)
    return m.cuda()


def test_unet_forward_small(golden, ddpm):
    g = golden["ddpm_small"]
    x_t, t = T(g["x_t"]).cuda(), T(g["t"]).cuda()
    with torch.no_grad():
        out, feats = ddpm.model.engine.forward(x_t, t, want_features=True)
    errs = {"out": _rel(out, g["unet_out"]), **{f"feat{i}": _rel(_nchw(f), g[f"feat{i}"]) for i, f in enumerate(feats)}}
    print("unet small rel errors:", errs)
    assert max(errs.values()) < TOL, errs
    with torch.no_grad():
        out_none = ddpm.model(x_t, None)
    assert _rel(out_none, g["unet_out_t_none"]) < TOL
    # folded-upsample path and materialised-upsample path agree with each other too
    ddpm.model.engine.fold_upsample = False
    try:
        with torch.no_grad():
            out2 = ddpm.model(x_t, t)
    finally:
        ddpm.model.engine.fold_upsample = True
    assert _rel(out2, g["unet_out"]) < TOL and _rel(out2, out) < TOL


def test_q_sample_loss_and_sampler_small(golden, ddpm):
    g = golden["ddpm_small"]
    x0, t, nz = T(g["x0"]).cuda(), T(g["t"]).cuda(), T(g["noise"]).cuda()
    x_t, _ = ddpm.forward_diffusion_model(x0 * 2 - 1, t, nz)
    assert np.array_equal(x_t.cpu().numpy(), g["x_t"])                       # bit-exact
    with torch.no_grad():
        loss = ddpm.train_step(x0, t=t, noise=nz)
    assert abs(loss.item() - float(g["ddpm_loss"])) < TOL * float(g["ddpm_loss"])
    z = synth_noise(g["x0"].shape, 1, "z").cuda()
    for ts in (500, 999, 0):
        got = ddpm.sample_timestep(T(g["x_t"]).cuda(), ts, noise=z)
        assert _rel(got, g[f"sample_t{ts}"]) < TOL, ts


def _tedm(n_steps, shared, steps):
    from tedm_b200.models import DatasetDM, tedm_classifier
    m = DatasetDM(Namespace(normalize=True, saved_diffusion_model="/nonexistent", t_steps_to_save=steps))
    if shared:
        m.classifier = tedm_classifier(n_steps)
    shapes = {**O.unet_param_shapes(prefix="diffusion_model.model."), **O.head_param_shapes(n_steps, shared)}
    missing = m.load_state_dict(synth_state_dict(shapes, 0), strict=False)
    assert not missing.unexpected_keys
    return m.eval().cuda()


# An untrained (random-init) head puts every logit at the decision threshold: the reference's logits have std 0.04
# around 0.03 and ~64 % of the pixels have |prob - 0.5| < 0.01, i.e. the logits are a near-cancelling residual of O(1)
# activations and amplify any rounding of the features ~4x.  What reduced-precision arithmetic can deliver there is
# pinned by the reference itself: tests/golden/noise_floor.json holds the reference run under
# torch.autocast(bfloat16) against the reference in fp32 on the SAME fixtures (tests/golden/make_golden_floor.py).  The
# bf16 path must be at least that close (it is ~10x closer); the fp32 mode meets >= 99.9 % on the same fixtures
# (tests/test_gpu_fp32.py).  The trained-head fixtures carry the north-star assertion for the bf16 path.
import json
import os

with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "noise_floor.json")) as _fh:
    FLOOR = json.load(_fh)


def _at_least_as_close_as_reference_bf16(name, logits_rel, mask_agree):
    fl = FLOOR[name]
    print(f"{name}: CUDA bf16 path logits rel {logits_rel:.4g} / mask {mask_agree:.5f}  vs  reference under autocast(bf16) "
          f"{fl['logits_rel']:.4g} / {fl['mask_agree']:.5f}")
    assert logits_rel <= fl["logits_rel"] and mask_agree >= fl["mask_agree"], (name, logits_rel, mask_agree, fl)


def _masks_agree_outside_threshold_band(mask, g, band=5e-3):
    diff = mask.cpu().numpy() != g["tedm_mask"]
    assert not (diff & (np.abs(g["tedm_prob"] - 0.5) >= band)).any(), "mask differs where the reference is decisive"


@pytest.mark.parametrize("tag", ["small", "full"])
def test_tedm_trained_head_masks(golden, tag):
    """North-star criterion: argmax masks agree with the reference on >= 99.9 % of the pixels."""
    g = golden[f"tedm_{tag}_trained"]
    steps = g["steps"].tolist()
    x0 = T(g["x0"]).cuda()
    noises = [T(g[f"noise{i}"]) for i in range(len(steps))]
    ted = _tedm(len(steps), True, steps)
    ted.load_state_dict({k: T(g[k]) for k in g.files if k.startswith("classifier.")}, strict=False)
    with FixedNoise([_interleaved(noises, x0.shape[0])]):
        mask, prob, logits = ted.segment(x0)
    lr = _rel(logits, g["tedm_logits"])
    agree = (mask.cpu().numpy() == g["tedm_mask"]).mean()
    print(f"tedm {tag} (trained head): logits rel {lr:.4g}, prob rel {_rel(prob, g['tedm_prob']):.4g}, mask agreement {agree:.5f}")
    assert lr < TOL
    if tag == "full":
        assert agree >= 0.999                      # the north-star criterion, at the BASELINE image size
        _at_least_as_close_as_reference_bf16("tedm_full_trained", lr, agree)
    else:
        # 2 x 32 x 32 = 2048 pixels: 0.1 % is two pixels, too coarse a grid for the percentage criterion;
        # require >= 99.8 % and that every disagreement sits inside the rounding band around prob = 0.5
        assert agree >= 0.998
        _masks_agree_outside_threshold_band(mask, g, band=2e-2)


class FixedNoise:
    def __init__(self, tensors):
        self.q = list(tensors)
    def __enter__(self):
        self.orig = torch.randn_like
        torch.randn_like = lambda x, **kw: self.q.pop(0).to(device=x.device, dtype=x.dtype)
        return self
    def __exit__(self, *a):
        torch.randn_like = self.orig


def _interleaved(noises, b):
    """per-step noise tensors (S x (B,1,H,W)) -> (B*S,1,H,W) in '(b step)' order"""
    return torch.stack(noises, dim=1).reshape(b * len(noises), *noises[0].shape[1:])


def test_tedm_and_ledm_small(golden):
    g = golden["tedm_small"]
    steps = g["steps"].tolist()
    x0 = T(g["x0"]).cuda()
    noises = [T(g[f"noise{i}"]) for i in range(len(steps))]
    ted = _tedm(len(steps), True, steps)
    with FixedNoise([_interleaved(noises, x0.shape[0])]):
        mask, prob, logits = ted.segment(x0)
    assert logits.shape == (x0.shape[0] * len(steps), 1, 32, 32)
    print("tedm small logits rel:", _rel(logits, g["tedm_logits"]))
    tol_untrained = FLOOR["tedm_full"]["logits_rel"]      # what the reference's own bf16 run achieves on an untrained head
    assert _rel(logits, g["tedm_logits"]) < tol_untrained
    assert _rel(prob, g["tedm_prob"]) < TOL
    _masks_agree_outside_threshold_band(mask, g)
    with FixedNoise([_interleaved(noises, x0.shape[0])]):
        assert _rel(ted(x0), g["tedm_logits"]) < tol_untrained               # nn.Module.__call__ path
    led = _tedm(len(steps), False, steps)
    with FixedNoise([_interleaved(noises, x0.shape[0])]):
        ll = led(x0)
    assert ll.shape == (x0.shape[0], 1, 32, 32)
    print("ledm small logits rel:", _rel(ll, g["ledm_logits"]))
    assert _rel(ll, g["ledm_logits"]) < tol_untrained
    # reference-format feature tensor (API compatibility)
    with FixedNoise([_interleaved(noises, x0.shape[0])]):
        feats = ted.extract_features(x0)
    assert feats.shape == (2, 960 * 3, 32, 32)
    sd = synth_state_dict({**O.unet_param_shapes(prefix="diffusion_model.model.")}, 0)
    sd.update(O.schedule_tables())
    with torch.no_grad():
        ref_feats = O.concat_features(O.extract_feature_maps(sd, x0.cpu(), steps, noises), 32)
    assert _rel(feats, ref_feats) < TOL


def test_tedm_full_size(golden):
    g = golden["tedm_full"]
    steps = g["steps"].tolist()
    x0 = T(g["x0"]).cuda()
    noises = [synth_noise((1, 1, 128, 128), 20 + i, "tedm") for i in range(len(steps))]
    ted = _tedm(len(steps), True, steps)
    with FixedNoise([_interleaved(noises, 1)]):
        mask, prob, logits = ted.segment(x0)
    lr = _rel(logits, g["tedm_logits"])
    agree = (mask.cpu().numpy() == g["tedm_mask"]).mean()
    print(f"tedm full (untrained head): logits rel {lr:.4g}, mask agreement {agree:.5f}")
    _at_least_as_close_as_reference_bf16("tedm_full", lr, agree)
    assert agree >= 0.99                                       # measured 0.9966; the fp32 mode carries >= 0.999 here
    _masks_agree_outside_threshold_band(mask, g)
    # one full-size UNet forward with intermediate feature checks
    dm = ted.diffusion_model
    t = torch.tensor([400], device="cuda")
    x_t, _ = dm.forward_diffusion_model(x0, t, noises[5].cuda())
    with torch.no_grad():
        out, feats = dm.model.engine.forward(x_t, t, want_features=True)
    assert _rel(out, g["unet_out_t400"]) < TOL
    for i, f in enumerate(feats):
        fn = _nchw(f)
        assert _rel(fn[:, :8], g[f"feat{i}_t400_first8ch"]) < TOL
        assert abs(fn.norm().item() - float(g[f"feat{i}_t400_norm"])) < TOL * float(g[f"feat{i}_t400_norm"])
        assert _rel(fn.mean(dim=(0, 2, 3)), g[f"feat{i}_t400_chmean"]) < 5e-2


class SameNoise:
    """torch.randn_like returns ONE resident device tensor every time: usable while a CUDA graph is captured and replayed."""
    def __init__(self, tensor):
        self.t = tensor
    def __enter__(self):
        self.orig = torch.randn_like
        torch.randn_like = lambda x, **kw: self.t
        return self
    def __exit__(self, *a):
        torch.randn_like = self.orig


def test_tedm_bench_configuration_b16_through_the_graphed_call(golden):
    """Parity AT the bench configuration (BASELINE configs[3]: B = 16 images x S = 8 timesteps, 128 x 128 -> 262 144 mask
    pixels), through the exact call bench.py times -- `segment(x, None, graph=True)`, one CUDA-graph replay -- against the
    live reference's masks with a head the reference trained (tests/golden/make_golden_floor.py)."""
    from tests.golden.synth import synth_images
    g = golden["tedm_b16_trained"]
    steps, b = g["steps"].tolist(), int(g["batch"])
    x0 = synth_images(b, 128, int(g["image_seed"])).cuda()
    noises = [synth_noise((b, 1, 128, 128), int(g["noise_seed0"]) + i, "tedm") for i in range(len(steps))]
    ted = _tedm(len(steps), True, steps)
    ted.load_state_dict({k: T(g[k]) for k in g.files if k.startswith("classifier.")}, strict=False)
    ref_mask = np.unpackbits(g["tedm_mask_packed"])[:b * 128 * 128].reshape(b, 1, 128, 128).astype(bool)
    nz = _interleaved(noises, b).cuda().float().contiguous()
    with SameNoise(nz):
        mask, prob, logits = ted.segment(x0, None, graph=True)      # warm-up + capture + first replay
        mask, prob, logits = ted.segment(x0, None, graph=True)      # a pure replay
        torch.cuda.synchronize()
        eager = ted.segment(x0)
    assert torch.equal(eager[0], mask) and torch.equal(eager[2], logits)      # the replay is the eager computation
    agree = (mask.cpu().numpy() == ref_mask).mean()
    lr = _rel(logits[:4 * len(steps)], g["tedm_logits_first4"])
    pr = np.abs(prob.cpu().numpy().astype(np.float64) - g["tedm_prob_f16"].astype(np.float64)).max()
    print(f"tedm B=16 x S=8 @128 (trained head, graph replay): logits rel {lr:.4g}, max |prob diff| {pr:.4g}, "
          f"mask agreement {agree:.5f} ({int((mask.cpu().numpy() != ref_mask).sum())} of {ref_mask.size} px)")
    assert lr < TOL
    assert agree >= 0.999                   # the north-star criterion at the bench configuration (measured 0.99936)
    _at_least_as_close_as_reference_bf16("tedm_b16_trained", lr, agree)
    # no image's result depends on its batch neighbours: image 5 alone gives the same mask
    with SameNoise(nz[5 * len(steps):6 * len(steps)].contiguous()):
        one = ted.segment(x0[5:6])
    assert (one[0].cpu().numpy() == mask[5:6].cpu().numpy()).mean() >= 0.9995


def test_batched_equals_per_image(ddpm):
    """Size-independent property at BASELINE batch size: a batch of 16 gives the same per-image result
    as 16 batches of 1 (no cross-image leakage in tiles that span images)."""
    x = torch.rand(16, 1, 32, 32, generator=torch.Generator().manual_seed(5)).cuda()
    t = torch.randint(0, 1000, (16,), generator=torch.Generator().manual_seed(6)).cuda()
    with torch.no_grad():
        full = ddpm.model(x, t)
        for i in (0, 7, 15):
            one = ddpm.model(x[i:i + 1], t[i:i + 1])
            assert _rel(one, full[i:i + 1]) < 5e-3


@pytest.mark.parametrize("size,mults", [(256, (1, 2, 4, 8)), (64, (1, 2, 4, 8)), (128, (1, 2, 4))])
def test_unet_forward_other_sizes_vs_oracle(size, mults):
    """`--img_size` / `--dim_mults` other than the defaults (config.py:33,41): 256^2 puts 1024 tokens through the mid
    attention (flash kernel) and 65536 pixels through the fused LinearAttention chunks; (1,2,4) at 128^2 gives a
    1024-token mid attention at 512... channels 256.  Inference only (the mid-attention backward handles n <= 256)."""
    from tedm_b200.models import Unet
    from tests.golden.synth import synth_images, synth_state_dict, synth_timesteps
    sd = synth_state_dict(O.unet_param_shapes(dim_mults=mults), 0)
    m = Unet(64, dim_mults=mults).eval()
    m.load_state_dict(sd)
    m.cuda()
    x, t = synth_images(1, size, 21), synth_timesteps(1, seed=3)
    with torch.no_grad():
        got = m(x.cuda(), t.cuda())
    ref = O.unet_forward(sd, x, t)
    err = _rel(got, ref)
    print(f"unet {size}x{size} mults {mults}: rel err {err:.4f}")
    assert err < TOL, err


def test_segment_graph_replay_is_bit_identical():
    """DatasetDM.segment(graph=True) (one CUDA-graph replay per call) == the eager call, for changing inputs, and re-captures
    when the weights change."""
    from tedm_b200.models import DatasetDM, tedm_classifier
    torch.manual_seed(0)
    steps = [10, 200, 600]
    m = DatasetDM(Namespace(normalize=True, saved_diffusion_model="", t_steps_to_save=steps))
    m.classifier = tedm_classifier(len(steps))
    m = m.cuda().eval()
    for seed in (1, 2, 3):
        g = torch.Generator(device="cuda").manual_seed(seed)
        x = torch.rand(2, 1, 64, 64, device="cuda", generator=g)
        nz = torch.randn(2, 1, 64, 64, device="cuda", generator=g)
        ref = [t.clone() for t in m.segment(x, nz)]
        got = m.segment(x, nz, graph=True)
        assert all(torch.equal(a, b) for a, b in zip(ref, got)), seed
    with torch.no_grad():
        m.classifier[7].bias.add_(0.5)                       # weights changed: the graph must be rebuilt, not replayed stale
    ref = [t.clone() for t in m.segment(x, nz)]
    got = m.segment(x, nz, graph=True)
    assert all(torch.equal(a, b) for a, b in zip(ref, got))
