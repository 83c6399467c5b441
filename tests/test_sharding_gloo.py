"""world_size-2 gloo test of the multi-GPU host logic (sequence sharding + final gather)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from swem_b200.sharding import assign_sequences, gather_results, merge_by_index


def test_assignment_is_balanced_and_complete():
    costs = [36 * 6 * 3600, 20 * 1 * 1590, 30 * 3 * 1620, 25 * 2 * 1590, 33 * 5 * 1620, 21 * 4 * 3600, 28 * 1 * 1620]
    for world in (1, 2, 4, 8):
        shards = assign_sequences(costs, world)
        assert sorted(i for s in shards for i in s) == list(range(len(costs)))
        loads = [sum(costs[i] for i in s) for s in shards]
        assert max(loads) <= sum(costs) / world + max(costs)
    assert assign_sequences(costs, 2) == assign_sequences(costs, 2)          # deterministic


def _worker(rank, world, port, costs):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        mine = assign_sequences(costs, world)[rank]
        local = {i: {'frames': int(costs[i]), 'checksum': float(torch.arange(costs[i]).sum())} for i in mine}
        merged = merge_by_index(gather_results(local))
        assert sorted(merged) == list(range(len(costs)))
        assert all(merged[i]['frames'] == costs[i] for i in merged)
        t = torch.tensor([float(rank + 1)])
        dist.all_reduce(t, op=dist.ReduceOp.MAX)                               # the bench's max-over-ranks timing
        assert t.item() == world
    finally:
        dist.destroy_process_group()


def test_two_rank_gather_gloo():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, [5, 9, 3, 7, 2]), nprocs=2, join=True)
