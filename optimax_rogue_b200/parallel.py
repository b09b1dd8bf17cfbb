"""Multi-GPU sharding: games are independent, so a box of G GPUs runs one process per GPU, each
owning a contiguous block of global game ids. Philox streams are keyed by the GLOBAL game id,
which makes every trajectory independent of the shard layout (1/2/4/8 GPUs give identical
games). There is no collective on the step path; the only exchange is the optional
end-of-rollout reduction of the 64-byte stats vector."""
import os
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


def _parse_cpulist(text: str) -> typing.List[int]:
    cpus = []
    for part in text.strip().split(','):
        if not part:
            continue
        lo, _, hi = part.partition('-')
        cpus.extend(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa_node(device_index: int) -> typing.Optional[dict]:
    """Pins the calling process to the CPUs of the NUMA node the GPU hangs off (``/sys/bus/pci/devices/<gpu>/
    local_cpulist``), so that the thread that ticks a GPU from HOST command / result buffers runs next to it and
    the pinned buffers it allocates afterwards are first-touched on that node. With one process per GPU on a
    multi-socket box this keeps the 8 ranks' PCIe traffic and stepping threads off each other's memory controllers.
    Returns what was done ({'numa_node', 'cpus'}) or None when the topology is not exposed (a single-node VM, no
    sysfs entry): then nothing is changed."""
    try:
        props = torch.cuda.get_device_properties(device_index)
        bdf = f'{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0'
        base = f'/sys/bus/pci/devices/{bdf}'
        with open(os.path.join(base, 'numa_node')) as f:
            node = int(f.read().strip())
        with open(os.path.join(base, 'local_cpulist')) as f:
            cpus = _parse_cpulist(f.read())
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if node < 0 or not cpus or len(cpus) >= len(allowed):
            return None                    # one node (or no information): nothing to gain
        os.sched_setaffinity(0, cpus)
        return {'numa_node': node, 'cpus': len(cpus)}
    except (OSError, ValueError, AttributeError, RuntimeError):
        return None
