"""Sequence-level sharding across GPUs: one process per GPU, whole sequences per rank, no collective
on the hot path -- only a final gather of per-sequence results.  The reference
has an unused helper for this (datasets/dataloader.py:39-52, splits the dataset by local_rank)."""
from __future__ import annotations

from typing import Any, List, Sequence

import torch.distributed as dist


def assign_sequences(costs: Sequence[float], world_size: int) -> List[List[int]]:
    """Longest-processing-time-first assignment of sequence indices to ranks (greedy, deterministic).

    ``costs[i]`` is any monotone proxy of sequence i's work (frames x objects x pixels)."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    load = [0.0] * world_size
    shards: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        shards[r].append(i)
        load[r] += costs[i]
    return shards


def gather_results(local: Any) -> List[Any]:
    """Every rank contributes one picklable object; every rank gets the list ordered by rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [local]
    out: List[Any] = [None] * dist.get_world_size()
    dist.all_gather_object(out, local)
    return out


def merge_by_index(per_rank: List[dict]) -> dict:
    """Per-rank {sequence index: result} dicts -> one dict; a duplicate index is a sharding bug."""
    merged: dict = {}
    for d in per_rank:
        for k, v in d.items():
            if k in merged:
                raise RuntimeError(f'sequence {k} was processed by two ranks')
            merged[k] = v
    return merged
