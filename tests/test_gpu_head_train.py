"""Training-mode head (BatchNorm batch statistics, BCE loss, gradients into the head parameters) of the CUDA path
against torch autograd through the oracle's head (reference form: upsample + concat + 1x1 convs), on the CUDA path's
own extracted features -- this isolates the head arithmetic from the bf16 rounding of the UNet features, which
tests/test_gpu_e2e.py bounds separately.  Reference: models/datasetDM_model.py:57-64,80-88;
trainers/train_datasetDM.py:30-42,88-99."""
from argparse import Namespace

import pytest
import torch
import torch.nn.functional as F

from oracle import tedm_oracle as O
from tests.golden.synth import synth_images, synth_noise, synth_state_dict

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _model(steps, shared):
    from tedm_b200.models import DatasetDM, tedm_classifier
    m = DatasetDM(Namespace(normalize=True, saved_diffusion_model="", t_steps_to_save=steps))
    if shared:
        m.classifier = tedm_classifier(len(steps))
    shapes = {**O.unet_param_shapes(prefix="diffusion_model.model."), **O.head_param_shapes(len(steps), shared)}
    missing = m.load_state_dict(synth_state_dict(shapes, 0), strict=False)
    assert not missing.unexpected_keys
    m = m.cuda()
    m.train()
    m.diffusion_model.eval()
    return m


@pytest.mark.parametrize("shared,steps", [(True, [10, 400, 800]), (False, [50, 150, 250])])
def test_head_training_step_matches_autograd(shared, steps):
    b, size, s = 2, 32, len(steps)
    m = _model(steps, shared)
    x0 = synth_images(b, size, 0).cuda()
    noises = [synth_noise((b, 1, size, size), 10 + i, "tedm") for i in range(s)]
    nz = torch.stack(noises, dim=1).reshape(b * s, 1, size, size).cuda()
    orig = torch.randn_like
    torch.randn_like = lambda x, **kw: nz.to(x.dtype)
    try:
        feats_ref = m.extract_features(x0)                      # (B, 960*S, H, W) fp32 from the native feature maps
        y = (torch.rand(b, 1, size, size, generator=torch.Generator().manual_seed(1)) > 0.5).float().cuda()
        if shared:
            y = y.repeat_interleave(s, dim=0)
        rm0 = {k: v.clone() for k, v in m.classifier.state_dict().items() if "running" in k or "num_batches" in k}
        logits = m(x0)
        loss = F.binary_cross_entropy_with_logits(logits, y, reduction="none").mean(dim=(2, 3)).mean()
        loss.backward()
    finally:
        torch.randn_like = orig
    torch.cuda.synchronize()
    # reference: the oracle's head in training mode on the same features, torch autograd (fp32, on the GPU for speed)
    sd = {f"classifier.{k}": v.detach().clone() for k, v in m.classifier.state_dict().items()}
    for k, v in rm0.items():
        if "running" in k:
            sd[f"classifier.{k}"] = v.clone()
    params = {k: v.requires_grad_(True) for k, v in sd.items() if v.is_floating_point() and "running" not in k}
    ref_logits = O.head_forward(sd, feats_ref, s, shared, training=True)
    ref_loss = F.binary_cross_entropy_with_logits(ref_logits, y, reduction="none").mean(dim=(2, 3)).mean()
    ref_loss.backward()
    assert logits.shape == ref_logits.shape
    lr = _rel(logits, ref_logits)
    print(f"head train (shared={shared}): logits rel {lr:.4g}; loss {loss.item():.6f} vs {ref_loss.item():.6f}")
    assert lr < 2e-2
    assert abs(loss.item() - ref_loss.item()) < 2e-3 * abs(ref_loss.item()) + 1e-5
    # The gradients of everything below BatchNorm are near-cancelling sums (sum over the batch of d/d(BN input) is zero;
    # only the ReLU mask leaves a remainder), so they amplify the bf16 rounding of the forward operands.  To separate
    # that from arithmetic errors the same fp32 autograd is run once more with the CUDA path's forward rounding points
    # emulated (straight-through): W1 in bf16, a1 = relu(z1) in bf16, layer 2 with BatchNorm-1 folded into bf16 weights.
    o = 1 if shared else 0
    emu = {k: v.detach().clone().requires_grad_(True) for k, v in params.items()}
    ste = lambda t: t + (t.to(torch.bfloat16).float() - t).detach()
    xin = feats_ref.reshape(b * s, -1, size, size) if shared else feats_ref
    pe = lambda i, n: emu[f"classifier.{i + o}.{n}"]
    a1 = ste(F.relu(F.conv2d(xin, ste(pe(0, "weight")), pe(0, "bias"))))
    mu1, var1 = a1.mean(dim=(0, 2, 3)), a1.var(dim=(0, 2, 3), unbiased=False)
    A1 = pe(2, "weight") * (var1 + 1e-5).rsqrt()
    C1 = pe(2, "bias") - mu1 * A1
    w2 = pe(3, "weight")[:, :, 0, 0]
    z2 = F.conv2d(a1, ste(w2 * A1[None, :])[:, :, None, None], pe(3, "bias") + w2 @ C1)
    h2 = F.batch_norm(F.relu(z2), None, None, pe(5, "weight"), pe(5, "bias"), training=True, eps=1e-5)
    emu_logits = F.conv2d(h2, pe(6, "weight"), pe(6, "bias"))
    F.binary_cross_entropy_with_logits(emu_logits, y, reduction="none").mean(dim=(2, 3)).mean().backward()
    assert _rel(logits, emu_logits) < 2e-3, _rel(logits, emu_logits)
    worst_emu = worst_ref = 0.0
    for name, p in m.classifier.named_parameters():
        ref, eg = params[f"classifier.{name}"].grad, emu[f"classifier.{name}"].grad
        assert p.grad is not None, name
        r_ref, r_emu, r_between = _rel(p.grad, ref), _rel(p.grad, eg), _rel(eg, ref)
        print(f"   grad {name}: vs fp32 reference {r_ref:.4g}, vs rounding-emulated reference {r_emu:.4g} "
              f"(emulated vs fp32: {r_between:.4g}; norm {ref.norm().item():.3g})")
        worst_emu, worst_ref = max(worst_emu, r_emu), max(worst_ref, r_ref)
    assert worst_emu < 2e-2, worst_emu          # arithmetic: matches the reference given the same forward rounding
    assert worst_ref < 0.25, worst_ref          # and stays in the neighbourhood of the pure fp32 gradient
    # BatchNorm running statistics advanced like torch's (momentum 0.1, unbiased variance)
    for bi, width in ((2 + o, 128), (5 + o, 32)):
        bn = m.classifier[bi]
        assert int(bn.num_batches_tracked) == int(rm0[f"{bi}.num_batches_tracked"]) + 1
        assert bn.running_mean.shape == (width,)
    ref_bn = torch.nn.BatchNorm2d(128).cuda().train()
    ref_bn.load_state_dict({"weight": sd[f"classifier.{2 + o}.weight"].detach(), "bias": sd[f"classifier.{2 + o}.bias"].detach(),
                            "running_mean": rm0[f"{2 + o}.running_mean"], "running_var": rm0[f"{2 + o}.running_var"],
                            "num_batches_tracked": torch.tensor(0)})
    with torch.no_grad():
        xin = feats_ref.reshape(b * s, -1, size, size) if shared else feats_ref
        ref_bn(F.relu(F.conv2d(xin, sd[f"classifier.{o}.weight"], sd[f"classifier.{o}.bias"])))
    assert _rel(m.classifier[2 + o].running_mean, ref_bn.running_mean) < 1e-2
    assert _rel(m.classifier[2 + o].running_var, ref_bn.running_var) < 1e-2


def test_head_training_reduces_loss():
    """A few Adam steps on the head alone (the reference's LEDM/TEDM training loop) lower the BCE loss."""
    from tedm_b200.optim import FusedAdam
    steps = [10, 400, 800]
    m = _model(steps, True)
    opt = FusedAdam(m.classifier.parameters(), lr=1e-3)
    x0 = synth_images(4, 32, 1).cuda()
    y = (x0 > x0.mean()).float().repeat_interleave(len(steps), dim=0)
    losses = []
    for _ in range(25):
        opt.zero_grad()
        loss = F.binary_cross_entropy_with_logits(m(x0), y)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert losses[-1] < 0.7 * losses[0], losses
    m.eval()
    mask, prob, _ = m.segment(x0)
    assert mask.shape == (4, 1, 32, 32)
