"""Data-parallel plumbing for the hot path.  The console and the losses have no cross-rank
dependency (every batch item is independent): ranks shard the batch, no
collective runs on the data path; torch.distributed is used only to agree on timing and to
average reported scalars."""
import torch
import torch.distributed as dist


def world_info():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_batch(global_batch: int, rank: int, world: int):
    """Contiguous [start, stop) slice of the batch owned by `rank` (remainder to the low ranks)."""
    base, rem = divmod(global_batch, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def max_over_ranks(value: float, device="cpu") -> float:
    """Device-timed durations are reported as the max over ranks."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def mean_over_ranks(value: torch.Tensor) -> torch.Tensor:
    """Average of a reported scalar (the reference logs with sync_dist=True, mst/system.py:343-364)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        value = value.clone()
        dist.all_reduce(value, op=dist.ReduceOp.SUM)
        value /= dist.get_world_size()
    return value


def aggregate_throughput(units_per_rank: float, ms_per_step_local: float, device="cpu") -> float:
    """Whole-job units per second: all ranks' units over the slowest rank's time."""
    _, world = world_info()
    return world * units_per_rank / (max_over_ranks(ms_per_step_local, device) / 1e3)
