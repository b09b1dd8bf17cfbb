"""Multi-GPU sharding: games are independent, so a box of G GPUs runs one process per GPU, each
owning a contiguous block of global game ids. Philox streams are keyed by the GLOBAL game id,
which makes every trajectory independent of the shard layout (1/2/4/8 GPUs give identical
games). There is no collective on the step path; the only exchange is the optional
end-of-rollout reduction of the 64-byte stats vector."""
import typing

import torch
import torch.distributed as dist


def shard_range(n_total: int, rank: int, world: int) -> typing.Tuple[int, int]:
    """(first global game id, count) owned by ``rank``: contiguous blocks, remainder spread over
    the first ranks."""
    if not 0 <= rank < world:
        raise ValueError('rank out of range')
    base, rem = divmod(n_total, world)
    count = base + (1 if rank < rem else 0)
    start = rank * base + min(rank, rem)
    return start, count


def gather_stats(stats: torch.Tensor) -> torch.Tensor:
    """Sums the int64[STAT_COUNT] counters over all ranks (NCCL on GPUs: NVLink/NVSwitch; gloo on
    CPU tensors in tests). A no-op when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    return stats


def max_over_ranks(value: float, device=None) -> float:
    """Device-timed durations are reported as the max over ranks."""
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
