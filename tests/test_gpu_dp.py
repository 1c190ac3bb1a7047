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
    dp_optimizer_step(opt, world)
    torch.cuda.synchronize()
    if rank == 0:
        torch.save({"flat": opt.flat_param.cpu(), "loss": loss.item()}, out)
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
    opt.step()
    ref = opt.flat_param.cpu()
    delta_ref, delta_dp = ref - before.cpu(), dp["flat"] - before.cpu()
    # Adam's first step moves every weight by lr * sign(g) (up to eps): compare the update directions and sizes
    agree = (torch.sign(delta_ref) == torch.sign(delta_dp)).float().mean().item()
    rel = ((delta_ref - delta_dp).norm() / delta_ref.norm()).item()
    print(f"DP vs single-GPU update: sign agreement {agree:.5f}, relative difference {rel:.4f}")
    assert agree > 0.995 and rel < 0.05, (agree, rel)
