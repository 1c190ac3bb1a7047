"""Two-GPU data-parallel training step == the single-GPU step on the concatenated batch (replicas + ONE
gradient all-reduce; the mean over ranks is taken inside Adam).  Skipped on boxes with fewer than two GPUs."""
import os
import socket
from argparse import Namespace

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _model():
    from oracle import tedm_oracle as O
    from tedm_b200.models import DiffusionModel
    from tests.golden.synth import synth_state_dict
    m = DiffusionModel(Namespace(normalize=True, dim_mults=[1, 2, 4])).train()
    sd = synth_state_dict(O.unet_param_shapes(dim_mults=(1, 2, 4), prefix="model."), 0)
    assert not m.load_state_dict(sd, strict=False).unexpected_keys
    return m


def _data():
    from tests.golden.synth import synth_images, synth_noise, synth_timesteps
    return synth_images(8, 64, 5), synth_timesteps(8, seed=5), synth_noise((8, 1, 64, 64), 5)


def _worker(rank, world, port, out):
    import torch.distributed as dist
    from tedm_b200.optim import FusedAdam
    from tedm_b200.trainers.utils import dp_optimizer_step
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    m = _model().cuda()
    opt = FusedAdam(m.parameters(), lr=1e-3)
    x, t, nz = _data()
    lo, hi = rank * 4, rank * 4 + 4
    loss = m.train_step(x[lo:hi].cuda(), t=t[lo:hi].cuda(), noise=nz[lo:hi].cuda())
    loss.backward()
    grad = opt.flat_grad().clone()
    dist.all_reduce(grad)                              # what dp_optimizer_step hands to Adam (with grad_scale = 1 / world)
    grad /= world
    dp_optimizer_step(opt, world)
    torch.cuda.synchronize()
    if rank == 0:
        torch.save({"flat": opt.flat_param.cpu(), "loss": loss.item(), "grad": grad.cpu()}, out)
    dist.destroy_process_group()


def test_two_gpu_step_matches_single_gpu_on_the_full_batch(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    from tedm_b200.optim import FusedAdam
    out = str(tmp_path / "dp.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    dp = torch.load(out)
    m = _model().cuda()
    opt = FusedAdam(m.parameters(), lr=1e-3)
    before = opt.flat_param.clone()
    x, t, nz = _data()
    m.train_step(x.cuda(), t=t.cuda(), noise=nz.cuda()).backward()
    grad_ref = opt.flat_grad().clone().cpu()
    opt.step()
    ref = opt.flat_param.cpu()
    # the mean gradient over the two shards IS the full-batch gradient: per-image arithmetic does not depend on the batch
    # split, only the fp32 summation order of the weight gradients does
    g_rel = ((dp["grad"].double() - grad_ref.double()).norm() / grad_ref.double().norm()).item()
    print(f"DP vs single-GPU gradient: relative difference {g_rel:.3g}")
    assert g_rel < 1e-3, g_rel
    delta_ref, delta_dp = ref - before.cpu(), dp["flat"] - before.cpu()
    # Adam's first step moves every weight by lr * sign(g) (up to eps): a gradient that is zero to rounding can land on
    # either side, and every such flip costs 2 lr -- hence the looser bound on the update than on the gradient
    agree = (torch.sign(delta_ref) == torch.sign(delta_dp)).float().mean().item()
    rel = ((delta_ref - delta_dp).norm() / delta_ref.norm()).item()
    print(f"DP vs single-GPU update: sign agreement {agree:.5f}, relative difference {rel:.4f}")
    assert agree > 0.995 and rel < 0.05, (agree, rel)


def _worker_graphed(rank, world, port, out):
    import torch.distributed as dist
    from tedm_b200.optim import FusedAdam
    from tedm_b200.train import GraphedTrainStep
    from tedm_b200.trainers.utils import dp_optimizer_step
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    m = _model().cuda()
    opt = FusedAdam(m.parameters(), lr=1e-3)
    torch.manual_seed(100 + rank)                    # rank-distinct t / noise streams, identical initial weights
    x, _, _ = _data()
    xr = x[rank * 4:rank * 4 + 4].cuda()
    step = GraphedTrainStep(m, opt, xr, warmup=2)
    for _ in range(3):
        step(xr)
    # the odd-sized last batch of an epoch goes through the eager path (trainers/train_CXR14.py) and rebinds p.grad
    opt.zero_grad()
    m.train_step(xr[:3]).backward()
    dp_optimizer_step(opt, world)
    for _ in range(3):
        step(xr)
    torch.cuda.synchronize()
    flat = opt.flat_param.clone()
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    if rank == 0:
        torch.save({"equal": all(torch.equal(gathered[0], g) for g in gathered[1:]), "step": opt._step,
                    "in_graph_reduction": step.reducer is not None}, out)
    step.close()                                    # the graphs hold captured NCCL work: release them before the group
    dist.destroy_process_group()


def test_graphed_dp_replicas_stay_bit_identical_across_an_eager_step(tmp_path):
    """after an eager odd-sized step the graphed step must still all-reduce the arena its graphs own;
    otherwise the replicas silently drift apart."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    out = str(tmp_path / "dpg.pt")
    mp.spawn(_worker_graphed, args=(2, _free_port(), out), nprocs=2, join=True)
    res = torch.load(out)
    print("graphed DP:", res)
    assert res["equal"] and res["step"] == 7, res


def _head_model():
    from oracle import tedm_oracle as O
    from tedm_b200.models import DatasetDM, tedm_classifier
    from tests.golden.synth import synth_state_dict
    steps = [10, 400, 800]
    m = DatasetDM(Namespace(normalize=True, saved_diffusion_model="", t_steps_to_save=steps))
    m.classifier = tedm_classifier(len(steps))
    shapes = {**O.unet_param_shapes(prefix="diffusion_model.model."), **O.head_param_shapes(len(steps), True)}
    assert not m.load_state_dict(synth_state_dict(shapes, 0), strict=False).unexpected_keys
    m = m.cuda().train()
    m.diffusion_model.eval()
    return m, steps


def _head_data():
    from tests.golden.synth import synth_images, synth_noise
    x = synth_images(4, 32, 9)
    return x, (x > 0.45).float(), synth_noise((4, 1, 32, 32), 9)


def _head_step(m, x, y, nz, world=1):
    from tedm_b200.autograd import bce_with_logits_rows
    from tedm_b200.optim import FusedAdam
    from tedm_b200.trainers.utils import dp_optimizer_step
    opt = FusedAdam(m.classifier.parameters(), lr=1e-3)
    loss = bce_with_logits_rows(m(x.cuda(), nz.cuda()), y.cuda()).mean()
    loss.backward()
    grads = torch.cat([p.grad.reshape(-1) for p in m.classifier.parameters()]).clone()
    dp_optimizer_step(opt, world)
    bn = [mod for mod in m.classifier if isinstance(mod, torch.nn.BatchNorm2d)]
    return {"loss": loss.item(), "grads": grads.cpu(), "flat": opt.flat_param.cpu().clone(),
            "rm": torch.cat([b.running_mean for b in bn]).cpu(), "rv": torch.cat([b.running_var for b in bn]).cpu()}


def _worker_head(rank, world, port, out, sync):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    m, _ = _head_model()
    m.sync_bn = sync
    x, y, nz = _head_data()
    lo, hi = rank * 2, rank * 2 + 2
    res = _head_step(m, x[lo:hi], y[lo:hi], nz[lo:hi], world)
    g = res["grads"].cuda()
    dist.all_reduce(g)                                  # what the optimiser step saw: the sum over ranks, / world in Adam
    res["grads"] = (g / world).cpu()
    losses = torch.tensor([res["loss"]], device="cuda")
    dist.all_reduce(losses)
    res["loss"] = losses.item() / world
    if rank == 0:
        torch.save(res, out)
    dist.destroy_process_group()


@pytest.mark.parametrize("sync", [True, False])
def test_head_training_two_gpus_vs_one(tmp_path, sync):
    """SURVEY 8e / VERDICT r1 missing-7: with `sync_bn` two replicas with 2 images each take the step the single device takes
    on all 4 (loss, parameter gradients, updated parameters and BatchNorm running statistics); without it the statistics
    are per replica, which is a different (declared) function: the test states how different."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    out = str(tmp_path / f"head_{sync}.pt")
    mp.spawn(_worker_head, args=(2, _free_port(), out, sync), nprocs=2, join=True)
    dp = torch.load(out)
    m, _ = _head_model()
    ref = _head_step(m, *_head_data())
    rel = lambda a, b: ((a.double() - b.double()).norm() / b.double().norm()).item()
    errs = {k: rel(dp[k], ref[k]) for k in ("grads", "flat", "rm", "rv")}
    errs["loss"] = abs(dp["loss"] - ref["loss"]) / abs(ref["loss"])
    print(f"head training, 2 GPUs vs 1 (sync_bn={sync}):", errs)
    if sync:
        # The frozen UNet sees 6 (image, timestep) pairs per rank against 12 on one device.  Its per-image arithmetic is the same,
        # but the fp32 context sums of the tcgen05 LinearAttention block are grouped by the tile -> CTA distribution, which
        # depends on the batch: 1e-7 differences that now and then flip a bf16 rounding of a feature.  Measured over the round:
        # rm 1.5e-7 ... 2.4e-5, rv 3e-8 ... 2e-6, grads 1.6e-4 ... 4.8e-4, flat 3.0e-4 ... 4.8e-4, loss 8e-8 ... 1.6e-6.
        assert errs["loss"] < 1e-5 and errs["rm"] < 1e-4 and errs["rv"] < 1e-4 and errs["grads"] < 2e-3 and errs["flat"] < 2e-3, errs
    else:
        # per-replica statistics: running buffers are rank 0's shard's, the loss differs in the third digit
        assert errs["loss"] < 5e-2 and errs["grads"] < 0.5, errs
