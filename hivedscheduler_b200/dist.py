"""Rank plumbing for the N>1 bench legs: one process per GPU, torch.distributed for the barrier and the
max-over-ranks timing.  The path itself shards by virtual cluster INSIDE a GPU (one CTA per group of VCs);
across GPUs round 1 runs independent replicas, so no data-path collective exists."""
from __future__ import annotations

import os
from typing import List, Sequence, Tuple


def dist_env() -> Tuple[int, int, int]:
    """(rank, world_size, local_rank) from the torchrun environment (defaults: single process)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def max_over_ranks(values: Sequence[float], device: str = "cpu") -> List[float]:
    """Element-wise MAX of per-rank timings (identity without an initialised process group)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t.tolist()]


def aggregate_throughput(units_per_rank: int, steps: int, seconds_max: float, world: int) -> float:
    """Whole-job throughput of `world` replicas: all units of all ranks over the slowest rank's time."""
    return world * units_per_rank * steps / seconds_max


def vc_owner(vc: int, n_partitions: int) -> int:
    """The partition (CTA today, GPU next) that owns a virtual cluster's events — same rule as the engine."""
    return vc % n_partitions
