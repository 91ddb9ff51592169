"""Multi-GPU layout of a replay: shard the (independent) filters, one process per GPU.

Filters never interact (each reference `KalmanFilter` object is self-contained,
`Python Kalman Filter/ExtendedKalmanFilter.py:6-11`) and time is a strict recurrence, so the only
sensible partition is contiguous slices of the filter axis; there is NO collective on the hot path.
The single communication step is the optional epilogue that gathers the final `[4, N]` states
(<= 256 MB at 16 M filters) with `torch.distributed.all_gather` -- NCCL over NVLink on the GPUs,
gloo in the CPU tests.
"""
from __future__ import annotations

from typing import Iterator

import torch
import torch.distributed as dist

ALIGN = 128   # filters per CTA; shard boundaries on this grid keep every shard TMA-eligible


def shard_bounds(n_filters: int, rank: int, world: int, align: int = ALIGN) -> tuple[int, int]:
    """[begin, end) of rank's slice.  Blocks of `align` filters are dealt out as evenly as possible
    (the first `rem` ranks get one more block); the ragged tail goes to the last non-empty rank."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    blocks = (n_filters + align - 1) // align
    per, rem = divmod(blocks, world)
    b0 = rank * per + min(rank, rem)
    b1 = b0 + per + (1 if rank < rem else 0)
    return min(b0 * align, n_filters), min(b1 * align, n_filters)


def shard_sizes(n_filters: int, world: int, align: int = ALIGN) -> list[int]:
    return [e - b for b, e in (shard_bounds(n_filters, r, world, align) for r in range(world))]


def time_chunks(n_steps: int, chunk_steps: int) -> Iterator[tuple[int, int]]:
    """[t0, t1) windows covering 0..n_steps; state is carried between consecutive windows."""
    if chunk_steps <= 0:
        raise ValueError("chunk_steps must be positive")
    t = 0
    while t < n_steps:
        yield t, min(t + chunk_steps, n_steps)
        t += chunk_steps


def chunk_steps_for_budget(n_local: int, bytes_budget: int, store_trajectory: bool = False) -> int:
    """Largest time chunk whose input (+ trajectory) buffers fit `bytes_budget` on one GPU."""
    per_step = n_local * 4 * (9 + (4 if store_trajectory else 0))
    return max(1, bytes_budget // max(per_step, 1))


def gather_states(x_local: torch.Tensor, n_filters: int, group=None, align: int = ALIGN) -> torch.Tensor:
    """All-gather per-rank `[k, n_local]` state slices into the full `[k, n_filters]` tensor
    (every rank receives it).  Slices may have different lengths: they are padded to the longest."""
    world = dist.get_world_size(group)
    sizes = shard_sizes(n_filters, world, align)
    k = x_local.shape[0]
    rank = dist.get_rank(group)
    if x_local.shape[1] != sizes[rank]:
        raise ValueError(f"rank {rank}: local slice has {x_local.shape[1]} filters, expected {sizes[rank]}")
    longest = max(sizes)
    padded = torch.zeros((k, longest), dtype=x_local.dtype, device=x_local.device)
    padded[:, :sizes[rank]] = x_local
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    return torch.cat([p[:, :s] for p, s in zip(parts, sizes)], dim=1)


def max_over_ranks(value: float, device, group=None) -> float:
    """MAX-reduce a host scalar (used for timing: a multi-GPU step takes as long as its slowest rank)."""
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def _parse_cpulist(text: str) -> set[int]:
    cpus: set[int] = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_local_cpus(device_index: int) -> set[int]:
    """CPUs of the NUMA node the GPU's PCIe root hangs off (sysfs `local_cpulist`); empty if unknown."""
    import os

    try:
        props = torch.cuda.get_device_properties(device_index)
        bdf = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/local_cpulist") as fh:
            cpus = _parse_cpulist(fh.read())
        return cpus & set(os.sched_getaffinity(0))
    except Exception:
        return set()


def bind_host_to_gpu(device_index: int) -> str | None:
    """Pin this process to the CPUs next to its GPU BEFORE it allocates pinned host buffers.

    The host-buffer replay (`replay_host`) streams 36 B per filter-step over PCIe; with one process per GPU
    on a two-socket box, a staging buffer that lives on the other socket makes every H2D copy cross the
    inter-socket link, which several ranks then share.  Linux allocates (and CUDA pins) pages on the node of the
    allocating thread, so binding first keeps each rank's stream on its own socket.  Returns a description of
    what was done (for logs), or None when the topology is not exposed."""
    import os

    cpus = gpu_local_cpus(device_index)
    if not cpus:
        return None
    os.sched_setaffinity(0, cpus)
    return f"{len(cpus)} cpus local to gpu {device_index} ({min(cpus)}-{max(cpus)})"
