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
