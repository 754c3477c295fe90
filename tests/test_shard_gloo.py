"""CPU, world_size 2, gloo: the N>1 bookkeeping of bench.py (scene sharding, max-over-ranks timing, whole-job throughput)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from graspnerf_b200.shard import shard_scenes, job_throughput, max_over_ranks


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    mine = shard_scenes(7, rank, world)
    elapsed = 10.0 * (rank + 1)                       # rank 1 is the slow one
    thr, ms = job_throughput(5, elapsed, dist)
    mx = max_over_ranks([float(rank), 3.0 - rank], dist)
    q.put((rank, mine, thr, ms, mx))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_timing():
    world, port = 2, _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    shards = [r[1] for r in res]
    assert sorted(shards[0] + shards[1]) == list(range(7)) and not set(shards[0]) & set(shards[1])
    for _, _, thr, ms, mx in res:
        assert ms == 20.0                              # max over ranks
        assert abs(thr - 2 * 5 / 0.020) < 1e-6         # all ranks' units / slowest rank's time
        assert mx == [1.0, 3.0]


def test_single_process_is_identity():
    thr, ms = job_throughput(4, 8.0)
    assert ms == 8.0 and abs(thr - 500.0) < 1e-9
    assert shard_scenes(5, 0, 1) == [0, 1, 2, 3, 4]


def _bucket_worker(rank, world, port, q):
    import os
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from graspnerf_b200.train import GradBucket
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Linear(3, 2))
    unused = torch.nn.Parameter(torch.ones(5))                   # never receives a gradient: must travel as zeros
    bucket = GradBucket(list(net.parameters()) + [unused])
    x = torch.full((2, 4), float(rank + 1))
    net(x).sum().backward()
    local = [p.grad.clone() for p in net.parameters()]
    bucket.allreduce(dist, 1.0 / world)
    q.put((rank, [g.numpy() for g in local], [p.grad.numpy().copy() for p in bucket.params]))
    dist.destroy_process_group()


def test_gradient_bucket_allreduce_two_ranks():
    """world_size 2 over gloo: the flat bucket all-reduce averages the ranks' gradients; grad-less parameters stay zero."""
    import numpy as np
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 200)
    procs = [ctx.Process(target=_bucket_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
    mean = [(a + b) / 2 for a, b in zip(res[0][1], res[1][1])]
    for r in res:
        for got, exp in zip(r[2][:4], mean):
            assert np.allclose(got, exp, rtol=1e-6, atol=1e-7)
        assert np.all(r[2][4] == 0)
