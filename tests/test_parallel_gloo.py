"""The N>1 host logic on CPU: world_size-2 gloo processes."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tedm_b200 import parallel as P


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, ws, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        # inference shards: disjoint, cover everything, no collective needed
        lo, hi = P.shard_range(37, rank, ws)
        # timing rule: max over ranks
        mx = P.max_over_ranks(10.0 + rank)
        # training: one averaged gradient all-reduce, bucketed
        torch.manual_seed(0)
        params = [torch.nn.Parameter(torch.zeros(s)) for s in ((300,), (17, 5), (1000,), ())]
        for i, p in enumerate(params):
            p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
        n_coll = P.allreduce_gradients(params, bucket_bytes=2000)
        out[rank] = (lo, hi, mx, n_coll, [float(p.grad.flatten()[0]) for p in params])
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_and_gradient_allreduce():
    ws, port = 2, _free_port()
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(ws, port, out), nprocs=ws, join=True)
        r0, r1 = out[0], out[1]
    assert (r0[0], r0[1], r1[0], r1[1]) == (0, 19, 19, 37)
    assert r0[2] == r1[2] == 11.0
    assert r0[3] == r1[3] and r0[3] >= 2                      # several buckets
    assert r0[4] == r1[4] == [1.5, 3.0, 4.5, 6.0]             # mean of (rank+1)*(i+1) over ranks 0,1


def test_shard_range_properties():
    for n in (0, 1, 7, 128, 1001):
        for ws in (1, 2, 3, 8):
            spans = [P.shard_range(n, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
