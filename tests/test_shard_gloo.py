"""CPU, world_size 2, gloo: the N>1 bookkeeping of bench.py (scene sharding, max-over-ranks timing, whole-job throughput)."""
import os

import pytest
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


class _ToyNet:
    """Stands in for the GraspNeRF mirror: a callable on a `data` dict with torch parameters (TrainStep only needs that)."""

    def __init__(self):
        import torch
        torch.manual_seed(3)
        self.mod = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 2))
        self.frozen = torch.nn.Parameter(torch.ones(3), requires_grad=False)      # like deviation_network.variance before step 1

    def parameters(self):
        return list(self.mod.parameters()) + [self.frozen]

    def __call__(self, data):
        return {'y': self.mod(data['x'])}


def _toy_loss(out, data):
    return ((out['y'] - data['t']) ** 2).mean()


def _toy_batch(n):
    import torch
    g = torch.Generator().manual_seed(11)
    return [{'x': torch.randn(4, 6, generator=g), 't': torch.randn(4, 2, generator=g)} for _ in range(n)]


def _trainstep_worker(rank, world, port, q, nscenes=4):
    import os
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from graspnerf_b200.train import TrainStep
    from graspnerf_b200.shard import shard_scenes
    net = _ToyNet()
    step = TrainStep(net, lr=1e-2, dist=dist, loss_fn=_toy_loss)
    batch = _toy_batch(nscenes)
    mine = [batch[i] for i in shard_scenes(nscenes, rank, world)]
    for _ in range(3):
        step(mine)
    q.put((rank, [p.detach().numpy().copy() for p in net.parameters()]))
    dist.destroy_process_group()


@pytest.mark.parametrize('nscenes', [4, 3])
def test_trainstep_two_ranks_equals_one_rank_on_the_global_batch(nscenes):
    """Data-parallel equivalence: 2 ranks with one gradient all-reduce per step == 1 process on the global
    batch, after three Adam steps - for an even split (2 + 2 scenes) and an UNEVEN one (2 + 1: the mean must divide by the
    global scene count, which travels in the same all-reduce); a parameter that does not require a gradient stays put."""
    import numpy as np
    import torch.multiprocessing as mp
    from graspnerf_b200.train import TrainStep
    ref = _ToyNet()
    step = TrainStep(ref, lr=1e-2, dist=None, loss_fn=_toy_loss)
    batch = _toy_batch(nscenes)
    for _ in range(3):
        step(batch)
    want = [p.detach().numpy() for p in ref.parameters()]
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29800 + (os.getpid() % 150)
    procs = [ctx.Process(target=_trainstep_worker, args=(r, 2, port + nscenes, q, nscenes)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
    for r in res:
        for got, exp in zip(r[1], want):
            assert np.allclose(got, exp, rtol=1e-5, atol=1e-6)
    assert np.array_equal(res[0][1][-1], np.ones(3, np.float32))
