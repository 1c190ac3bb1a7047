"""Supervised-segmentation edges of the path on the GPU: BCE loss + gradient, dice / precision /
recall, uint8 input transport, the baseline `Unet(timestep=None)` segmenter, the trainer loops and the checkpoint format.

Bars: loader arithmetic, thresholds and TP/FP/FN counts bit-exact against the live reference's golden vectors
(tests/golden/seg_small.npz); BCE rows / gradient within 1e-5 relative (fp32 sums in a different order); UNet-as-
segmenter logits within the bf16 activation budget (2e-2)."""
import os
import subprocess
import sys
from argparse import Namespace

import numpy as np
import pytest
import torch

from oracle import tedm_oracle as O
from tests.golden.synth import synth_images, synth_state_dict

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
T = lambda a: torch.from_numpy(np.asarray(a))


def _rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def test_bce_rows_loss_and_grad_match_reference(golden):
    from tedm_b200 import native as N
    g = golden["seg_small"]
    logits, y = T(g["logits"]).cuda(), T(g["y"]).cuda()
    loss, rows, grad = N.bce_logits(logits, y, want_grad=True)
    assert np.allclose(rows.cpu().numpy(), g["bce_rows"], rtol=1e-5, atol=1e-7)
    assert abs(loss.item() - float(g["bce_loss"])) < 1e-6
    assert _rel(grad, g["bce_grad"]) < 1e-5
    # deterministic: fixed summation order
    loss2, rows2, _ = N.bce_logits(logits, y)
    assert torch.equal(rows, rows2) and torch.equal(loss, loss2)


def test_bce_autograd_and_per_timestep_rows(golden):
    from tedm_b200.autograd import bce_with_logits_rows
    g = golden["seg_small"]
    logits = T(g["logits"]).cuda().requires_grad_(True)
    y = T(g["y"]).cuda()
    rows = bce_with_logits_rows(logits, y)
    w = torch.linspace(0.5, 2.0, rows.numel(), device="cuda").view_as(rows)         # a non-uniform reduction of the rows
    (rows * w).sum().backward()
    ref = T(g["logits"]).requires_grad_(True)
    (O.bce_rows(ref, T(g["y"]), int(g["n_steps"])) * w.cpu()).sum().backward()
    assert _rel(logits.grad, ref.grad) < 1e-5


@pytest.mark.parametrize("shape,s", [((128, 1, 128, 128), 8), ((16, 1, 128, 128), 1), ((6, 2, 33, 35), 1)])
def test_bce_and_metrics_full_size_vs_oracle(shape, s):
    """BASELINE sizes (B*S = 128 rows of 128 x 128) and a ragged, multi-channel case."""
    from tedm_b200 import native as N
    gen = torch.Generator().manual_seed(3)
    logits = torch.randn(shape, generator=gen) * 3
    y = (torch.rand((shape[0] // s,) + shape[1:], generator=gen) > 0.5).float()
    loss, rows, grad = N.bce_logits(logits.cuda(), y.cuda(), want_grad=True)
    ref_rows = O.bce_rows(logits, y, s)
    assert _rel(rows, ref_rows) < 1e-5 and abs(loss.item() - ref_rows.mean().item()) < 1e-5
    lg = logits.clone().requires_grad_(True)
    O.bce_rows(lg, y, s).mean().backward()
    assert _rel(grad, lg.grad) < 1e-5
    m = N.seg_metrics(logits.cuda(), y.cuda()).cpu()
    y_hat = torch.sigmoid(logits.cuda()).cpu() > .5            # torch's CUDA sigmoid: the formula the kernel uses
    ref = O.seg_metrics(y_hat, y, s)
    for i, k in enumerate(("dice", "precision", "recall")):
        assert torch.equal(m[..., i], ref[k]), k
    m2 = N.seg_metrics(y_hat.cuda(), y.cuda()).cpu()
    assert torch.equal(m, m2)


def test_seg_metrics_match_reference_incl_nan_rows(golden):
    from tedm_b200 import native as N
    from tedm_b200.trainers.train_baseline import dice, precision, recall
    g = golden["seg_small"]
    y = T(g["y"]).cuda()
    for pred in (T(g["y_hat"]).cuda(), T(g["logits"]).cuda()):
        m = N.seg_metrics(pred, y).cpu().numpy()
        for i, k in enumerate(("dice", "precision", "recall")):
            assert np.array_equal(m[..., i], g[k], equal_nan=True), k
    y_rep = y.repeat_interleave(int(g["n_steps"]), dim=0)
    yh = T(g["y_hat"]).cuda()
    assert np.array_equal(dice(yh, y_rep).cpu().numpy(), g["dice"], equal_nan=True)
    assert np.array_equal(precision(yh, y_rep).cpu().numpy(), g["precision"], equal_nan=True)
    assert np.array_equal(recall(yh, y_rep).cpu().numpy(), g["recall"], equal_nan=True)
    with pytest.raises(ValueError):
        N.seg_metrics(yh[:5], y)                    # 5 rows are not a multiple of 3 labels


def test_u8_transport_bit_exact(golden):
    from tedm_b200 import native as N
    g = golden["seg_small"]
    assert np.array_equal(N.u8_to_unit(T(g["u8_img"]).cuda()).cpu().numpy()[None], g["f32_img"])
    for k in ("overlap", "disjoint"):
        lab = N.u8_masks_to_label(T(g[f"u8_masks_{k}"])[None].cuda())
        assert np.array_equal(lab[0].cpu().numpy(), g[f"label_{k}"]), k
    # every byte value, a ragged length (tail path) and a BASELINE-sized batch
    allv = torch.arange(256, dtype=torch.uint8).repeat(3)[:700]
    assert torch.equal(N.u8_to_unit(allv.cuda()).cpu(), O.to_tensor_u8(allv))
    big = torch.randint(0, 256, (64, 1, 128, 128), dtype=torch.uint8)
    assert torch.equal(N.u8_to_unit(big.cuda()).cpu(), O.to_tensor_u8(big))
    masks = torch.randint(0, 256, (5, 2, 128, 128), dtype=torch.uint8)
    ref = torch.stack([O.jsrt_label(m) for m in masks])
    assert torch.equal(N.u8_masks_to_label(masks.cuda()).cpu(), ref)


def test_device_loader_matches_host_arithmetic():
    from tedm_b200.dataloaders.device_loader import DeviceLoader, SyntheticXray
    ds = SyntheticXray(10, 64, labelled=True, seed=5)
    dl = DeviceLoader(ds, 4, shuffle=False, num_workers=0, labelled=True)
    xs, ys = zip(*[(x.cpu(), y.cpu()) for x, y in dl])
    assert [x.shape[0] for x in xs] == [4, 4, 2] and len(dl) == 3
    x_all, y_all = torch.cat(xs), torch.cat(ys)
    for i in range(10):
        img, masks = ds[i]
        assert torch.equal(x_all[i], O.to_tensor_u8(img)) and torch.equal(y_all[i], O.jsrt_label(masks))
    assert dl.h2d_bytes == 10 * 3 * 64 * 64                       # uint8 over PCIe: 1 image + 2 mask planes per sample
    assert 0.02 < y_all.mean().item() < 0.6


def test_unet_as_segmenter_timestep_none_vs_oracle():
    """Baseline experiment (trainers/train_baseline.py:180-185): Unet(out_dim=1)(x) with no timestep."""
    from tedm_b200.models import Unet
    sd = synth_state_dict(O.unet_param_shapes(out_dim=1), 0)
    m = Unet(64, dim_mults=(1, 2, 4, 8), channels=1, out_dim=1).eval()
    m.load_state_dict(sd)
    m.cuda()
    x = synth_images(2, 64, 11)
    with torch.no_grad():
        got = m(x.cuda())
    ref = O.unet_forward(sd, x, None)
    assert _rel(got, ref) < 2e-2, _rel(got, ref)


def test_baseline_training_steps_reduce_loss(tmp_path):
    """A few steps of the baseline loop on synthetic pairs: loss goes down, validation metrics are finite, the
    checkpoint has the reference's keys and reloads."""
    from tedm_b200.dataloaders.device_loader import build_synthetic_dataloaders
    from tedm_b200.models import Unet
    from tedm_b200.optim import FusedAdam
    from tedm_b200.trainers import train_baseline as TB
    from tedm_b200.trainers.utils import TensorboardLogger, seed_everything
    seed_everything(0)
    cfg = Namespace(device="cuda", debug=False, log_freq=1, val_freq=6, max_steps=6, max_val_steps=1, log_dir=tmp_path,
                    shared_weights_over_timesteps=False, early_stop=False, img_size=64, batch_size=4)
    m = Unet(64, dim_mults=(1, 2, 4), channels=1, out_dim=1).cuda().train()
    opt = FusedAdam(m.parameters(), lr=2e-3)
    dls = build_synthetic_dataloaders(64, 4, labelled=True, n_train=8, n_val=4)
    first = TB.validate(cfg, m, dls["val"])
    TB.train(cfg, m, opt, dls["train"], dls["val"], TensorboardLogger(enabled=False), None, 0)
    last = TB.validate(cfg, m, dls["val"])
    assert last["val/loss"] < first["val/loss"], (first, last)
    assert all(np.isfinite(last[k]) for k in ("val/loss", "val/dice", "val/precision", "val/recall"))
    ck = torch.load(tmp_path / "best_model.pt", weights_only=False)
    assert sorted(ck) == ["config", "model_state_dict", "optimizer_state_dict", "step"]
    m2 = Unet(64, dim_mults=(1, 2, 4), channels=1, out_dim=1)
    m2.load_state_dict(ck["model_state_dict"])


def test_fused_adam_loads_torch_adam_state():
    """Reference checkpoints carry torch.optim.Adam state (train_CXR14.py:96-114): FusedAdam continues from it exactly
    like torch's Adam would."""
    from tedm_b200.optim import FusedAdam
    torch.manual_seed(0)
    ps = [torch.nn.Parameter(torch.randn(33, 7, device="cuda")), torch.nn.Parameter(torch.randn(129, device="cuda"))]
    qs = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    ref = torch.optim.Adam(ps, lr=1e-2)
    grads = [[torch.randn_like(p) for p in ps] for _ in range(3)]
    for p, g in zip(ps, grads[0]):
        p.grad = g.clone()
    ref.step()
    state = ref.state_dict()
    with torch.no_grad():
        for q, p in zip(qs, ps):
            q.copy_(p)
    fused = FusedAdam(qs, lr=1e-2)
    fused.load_state_dict(state)
    for k in (1, 2):
        for p, q, g in zip(ps, qs, grads[k]):
            p.grad, q.grad = g.clone(), g.clone()
        ref.step()
        fused.step()
    for p, q in zip(ps, qs):
        assert torch.allclose(p, q, rtol=1e-5, atol=1e-6)
    sd = fused.state_dict()
    assert float(sd["state"][0]["step"]) == 3.0 and sd["state"][0]["exp_avg"].shape == (33, 7)
    # ... and the other way round: torch.optim.Adam (the reference's optimiser) resumes from a FusedAdam checkpoint
    rs = [torch.nn.Parameter(q.detach().clone()) for q in qs]
    back = torch.optim.Adam(rs, lr=1e-2)
    import io
    buf = io.BytesIO()
    torch.save(sd, buf)                                  # through a checkpoint file, as trainers' save() / load() do
    buf.seek(0)
    back.load_state_dict(torch.load(buf, weights_only=False))
    g3 = [torch.randn_like(p) for p in ps]
    for p, q, r, g in zip(ps, qs, rs, g3):
        p.grad, q.grad, r.grad = g.clone(), g.clone(), g.clone()
    ref.step()
    fused.step()
    back.step()
    for p, q, r in zip(ps, qs, rs):
        assert torch.allclose(p, q, rtol=1e-5, atol=1e-6) and torch.allclose(p, r, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("experiment", ["TEDM", "LEDM", "baseline", "img_only"])
def test_train_cli_debug_step(tmp_path, experiment):
    """`python train.py --experiment X --debug` (reference command line): one training step + one validation pass."""
    cmd = [sys.executable, os.path.join(ROOT, "train.py"), "--experiment", experiment, "--dataset", "synthetic", "--debug",
           "--batch_size", "2", "--log_dir", str(tmp_path / "run"), "--n_sampled_imgs", "0", "--timesteps", "1000",
           "--saved_diffusion_model", "/nonexistent"]
    if experiment in ("TEDM", "LEDM"):
        cmd += ["--n_labelled_images", "6"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "Train loss" in r.stdout and "Validation loss" in r.stdout, r.stdout[-2000:]


def test_evaluate_shared_weights_matches_reference_formulae():
    """Per-timestep + ensemble report (testing_shared_weights.py:113-138) against the oracle's metrics on the same logits."""
    from tedm_b200.dataloaders.device_loader import build_synthetic_dataloaders
    from tedm_b200.evaluate import evaluate_shared_weights
    from tedm_b200.models import DatasetDM, tedm_classifier
    steps = [10, 200, 600]
    torch.manual_seed(0)
    m = DatasetDM(Namespace(normalize=True, saved_diffusion_model="", t_steps_to_save=steps, dim_mults=[1, 2, 4, 8]))
    m.classifier = tedm_classifier(len(steps))
    m = m.cuda().eval()
    dl = build_synthetic_dataloaders(128, 2, labelled=True, n_train=2, n_val=4)["val"]
    torch.manual_seed(5)
    res = evaluate_shared_weights(m, dl)
    assert res["n_images"] == 4 and sorted(res["per_timestep"]) == steps
    # same logits again (same noise seed) through the oracle's formulae
    torch.manual_seed(5)
    rows, ens = [], []
    for x, y in dl:
        logits = m(x).cpu()
        pr = torch.sigmoid(logits).reshape(x.shape[0], len(steps), 1, 128, 128)
        rows.append(torch.stack([O.seg_metrics(pr[:, i] > .5, y.cpu())["dice"] for i in range(len(steps))], dim=1))
        ens.append(O.seg_metrics(pr.mean(1) > .5, y.cpu())["dice"])
    rows, ens = torch.cat(rows), torch.cat(ens)
    for i, t in enumerate(steps):
        a, b = res["per_timestep"][t]["dice"][0], rows[:, i].mean().item()
        assert (np.isnan(a) and np.isnan(b)) or abs(a - b) < 1e-6, (t, a, b)
    a, b = res["ensemble"]["dice"][0], ens.mean().item()
    assert (np.isnan(a) and np.isnan(b)) or abs(a - b) < 1e-6


def test_val_step_and_sampling_driver():
    """val_step batches its timestep grid along the image axis (diffusion_model.py:145-156); sample_images /
    sample_plot_image drive sample_timestep for T steps (trainers/utils.py:62-98)."""
    from tedm_b200.models import DiffusionModel
    from tedm_b200.trainers.utils import sample_images, sample_plot_image
    torch.manual_seed(0)
    m = DiffusionModel(Namespace(normalize=True, timesteps=40, dim_mults=[1, 2, 4])).cuda().eval()
    x = synth_images(4, 64, 3).cuda()
    torch.manual_seed(1)
    v = m.val_step(x, t_steps=4).item()
    torch.manual_seed(2)
    ref = np.mean([m.train_step(x, t=torch.full((4,), t, device="cuda")).item() for t in range(0, 40, 10)])
    assert abs(v - ref) < 0.03 * ref, (v, ref)          # same timestep grid, independent noise draws
    final, snaps = sample_images(m, 40, 64, 2)
    assert final.shape == (2, 1, 64, 64) and len(snaps) == 8 and torch.isfinite(final).all()
    assert final.min() >= -1e-6 and final.max() <= 1 + 1e-6      # dynamic thresholding keeps x in [-1, 1] -> [0, 1]
    grid = sample_plot_image(m, 40, 64, 2, normalized=True)
    assert grid.shape == (2, 3, 2 * 66 + 2, 4 * 66 + 2)


def test_train_then_evaluate_cli(tmp_path):
    """`train.py --experiment TEDM` writes a checkpoint in the reference's format; `python -m tedm_b200.evaluate -e <dir>`
    (the reference's testing_shared_weights.py) reloads it and reports per-timestep + ensemble metrics."""
    run = tmp_path / "run"
    cmd = [sys.executable, os.path.join(ROOT, "train.py"), "--experiment", "TEDM", "--dataset", "synthetic", "--batch_size", "2",
           "--log_dir", str(run), "--n_labelled_images", "6", "--max_steps", "3", "--val_freq", "3", "--log_freq", "1",
           "--max_val_steps", "1", "--saved_diffusion_model", "/nonexistent"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    exp = tmp_path / "TEDM" / "6" / "run"
    assert (exp / "best_model.pt").exists(), r.stdout[-2000:]
    r = subprocess.run([sys.executable, "-m", "tedm_b200.evaluate", "-e", str(exp)], capture_output=True, text=True,
                       timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "synthetic_val 800 metrics" in r.stdout and "synthetic_test metrics" in r.stdout, r.stdout[-2000:]
    assert (exp / "synthetic_val_metrics.pt").exists()


def test_graphed_sampler_matches_eager_steps():
    """One reverse step replayed from the CUDA graph (schedule values from device memory) == sample_timestep, bit for bit,
    at several timesteps incl. t = 0 (no noise); the loop driver uses it by default."""
    from tedm_b200.models import DiffusionModel
    from tedm_b200.trainers.utils import GraphedSampler, sample_images
    torch.manual_seed(0)
    m = DiffusionModel(Namespace(normalize=True, timesteps=40, dim_mults=[1, 2, 4])).cuda().eval()
    gs = GraphedSampler(m, 2, 1, 64)
    gen = torch.Generator(device="cuda").manual_seed(1)
    for t in (39, 17, 1, 0):
        x = torch.randn(2, 1, 64, 64, device="cuda", generator=gen)
        z = torch.randn(2, 1, 64, 64, device="cuda", generator=gen)
        gs.x.copy_(x)
        got = gs.step(t, noise=z).clone()
        ref = m.sample_timestep(x, t, noise=z)
        assert torch.equal(got, ref), (t, (got - ref).abs().max().item())
    final, snaps = sample_images(m, 40, 64, 2)
    assert final.shape == (2, 1, 64, 64) and len(snaps) == 8 and torch.isfinite(final).all()
    assert final.min() >= -1e-6 and final.max() <= 1 + 1e-6
