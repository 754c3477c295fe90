"""Scene-level data parallelism helpers: the path shards by independent scenes, one process per GPU,
no data-path collective in inference.  Only the bookkeeping (which rank takes which scenes, max-over-ranks timing) uses
torch.distributed; it is backend-agnostic so the logic is unit-tested on CPU with gloo."""
import torch


def shard_scenes(num_scenes, rank, world):
    """Rank r of W takes scenes r, r+W, r+2W, ... ('rank r of W processes takes scenes r::W')."""
    return list(range(rank, num_scenes, world))


def max_over_ranks(values, dist=None, device='cpu'):
    """Element-wise max of a list of floats over all ranks (the timing rule of the bench contract)."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


def job_throughput(units_per_rank, elapsed_ms_local, dist=None, device='cpu'):
    """Whole-job units/s = (units all ranks processed) / (max over ranks of the elapsed time)."""
    world = dist.get_world_size() if (dist is not None and dist.is_initialized()) else 1
    (ms,) = max_over_ranks([elapsed_ms_local], dist, device)
    return world * units_per_rank / (ms / 1e3), ms
