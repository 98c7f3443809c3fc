"""Multi-GPU partitioning of the path: frames of one stream are serially
dependent through the lookahead, independent streams are not, so the unit of sharding is the
stream: stream i -> rank i mod world_size, one session per stream, no data-path collective.
Timing of a multi-rank run is the max over ranks (each rank times its own device)."""
import torch
import torch.distributed as dist


def streams_of_rank(n_streams: int, rank: int, world_size: int):
    """Global stream ids owned by `rank` (round-robin, like 'stream i -> GPU i mod G')."""
    return [s for s in range(n_streams) if s % world_size == rank]


def max_over_ranks(value: float, device=None) -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: int, device=None) -> int:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return int(value)
    t = torch.tensor([value], dtype=torch.int64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return int(t.item())
