"""Population sharding across the GPUs of one box.

Robots are independent, so the population partitions into contiguous robot ranges, one per rank, and the
hot path needs no inter-GPU traffic.  The only exchange is the end-of-job gather of the per-robot fitness
scalars (a few floats per robot), done with ``torch.distributed.all_gather`` (NCCL over NVLink on the
GPUs, gloo in the CPU tests).  Mirrors what ``evaluate_all`` does host-side in the reference
(evosoro/tools/evaluation.py:72-158: fan the population out, collect one fitness file per robot).
"""
from __future__ import annotations

from typing import Tuple

import numpy as np


def shard_range(n_robots: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous range [lo, hi) of robot indices owned by ``rank``: ceil(n/world) per rank, last ranks may
    get fewer (or none)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    per = (n_robots + world - 1) // world
    lo = min(n_robots, rank * per)
    return lo, min(n_robots, lo + per)


def gather_fitness(local: np.ndarray, n_robots: int, dist=None, device=None) -> np.ndarray:
    """All-gather the per-robot fitness rows of every shard into the population-ordered array
    ``[n_robots, k]`` on every rank.  ``dist`` is an initialised ``torch.distributed`` (or None for a
    single process); shards may be ragged (padded to ceil(n/world) rows on the wire)."""
    local = np.ascontiguousarray(local, dtype=np.float64)
    if local.ndim == 1:
        local = local[:, None]
    if dist is None or dist.get_world_size() == 1:
        assert local.shape[0] == n_robots
        return local
    import torch
    world = dist.get_world_size()
    per = (n_robots + world - 1) // world
    pad = np.zeros((per, local.shape[1]))
    pad[: local.shape[0]] = local
    t = torch.from_numpy(pad)
    if device is not None:
        t = t.to(device)
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    rows = []
    for r in range(world):
        lo, hi = shard_range(n_robots, world, r)
        rows.append(out[r][: hi - lo].cpu().numpy())
    return np.concatenate(rows, axis=0)


def max_over_ranks(x: float, dist=None, device=None) -> float:
    """Timing reduction of the bench contract: the slowest rank defines the step time."""
    if dist is None or dist.get_world_size() == 1:
        return float(x)
    import torch
    t = torch.tensor([x], dtype=torch.float64)
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])
