"""Debug driver for the graphed data-parallel step (2 ranks): prints a line per stage; dumps stacks if a stage hangs."""
import faulthandler, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist, torch.multiprocessing as mp


def worker(rank, world, port):
    faulthandler.dump_traceback_later(100, exit=True)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from tests.test_gpu_dp import _model, _data
    from tedm_b200.optim import FusedAdam
    from tedm_b200.train import GraphedTrainStep
    from tedm_b200.trainers.utils import dp_optimizer_step
    say = lambda *a: print(f"[rank {rank} t={time.time() % 1000:.1f}]", *a, flush=True)
    m = _model().cuda()
    opt = FusedAdam(m.parameters(), lr=1e-3)
    torch.manual_seed(100 + rank)
    x, _, _ = _data()
    xr = x[rank * 4:rank * 4 + 4].cuda()
    say("constructing GraphedTrainStep")
    step = GraphedTrainStep(m, opt, xr, warmup=2)
    say("captured; reducer", step.reducer is not None)
    for i in range(3):
        step(xr)
        torch.cuda.synchronize()
        say("graphed step", i)
    opt.zero_grad()
    m.train_step(xr[:3]).backward()
    say("eager backward done")
    dp_optimizer_step(opt, world)
    torch.cuda.synchronize()
    say("eager dp step done")
    for i in range(3):
        step(xr)
        torch.cuda.synchronize()
        say("graphed step after eager", i)
    flat = opt.flat_param.clone()
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    say("equal across ranks:", all(torch.equal(gathered[0], g) for g in gathered[1:]))
    dist.destroy_process_group()
    say("done")


if __name__ == "__main__":
    mp.spawn(worker, args=(2, 29533), nprocs=2, join=True)
