"""world_size-2 CPU tests (gloo) of the multi-GPU host logic: capacity sharding, round-robin
episode placement, the gradient all-reduce helper and the start-up broadcast.  The kernels
themselves need a GPU; what is covered here is everything rank-dependent that runs on the host."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, out_dir: str):
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    sys.path[:0] = [str(root), str(root / 'advanced-soft-actor-critic_b200')]
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from asac_b200 import dist as adist
    try:
        assert adist.world() == (rank, world)
        # capacity sharding: the global ring rounded down to a power of two, split evenly
        assert adist.shard_capacity(524288, world) == 262144
        assert adist.shard_capacity(1000, world) == 256
        assert adist.shard_capacity(3, world) == 2
        # whole episodes go to one shard, round robin: every episode has exactly one owner
        owners = [adist.episode_owner(i, world) for i in range(10)]
        mine = torch.tensor([1.0 if o == rank else 0.0 for o in owners])
        total = mine.clone()
        dist.all_reduce(total)
        assert torch.equal(total, torch.ones(10))
        assert abs(int(mine.sum()) - 10 // world) <= 1
        # gradient exchange: SUM all-reduce in place, the Adam kernel applies 1 / world
        g = torch.arange(8, dtype=torch.float32) * (rank + 1)
        adist.all_reduce_sum_(g)
        want = torch.arange(8, dtype=torch.float32) * sum(r + 1 for r in range(world))
        assert torch.equal(g, want)
        mean = g * (1.0 / world)
        assert torch.allclose(mean, torch.arange(8, dtype=torch.float32) * (world + 1) / 2)
        # start-up: every rank adopts rank 0's parameters and Adam moments
        params = [torch.full((5,), float(rank + 1)), torch.full((3,), float(10 * rank))]
        adist.broadcast_(params, src=0)
        assert torch.equal(params[0], torch.ones(5)) and torch.equal(params[1], torch.zeros(3))
        # replicas that apply the same averaged gradient to the same start stay bit-identical
        w = params[0].clone()
        local = torch.full((5,), float(rank))
        adist.all_reduce_sum_(local)
        w -= 0.1 * local / world
        gathered = [torch.zeros(5) for _ in range(world)]
        dist.all_gather(gathered, w)
        assert all(torch.equal(gathered[0], t) for t in gathered)
        np.save(os.path.join(out_dir, f'ok{rank}.npy'), np.array([1]))
    finally:
        dist.destroy_process_group()


def test_world_size_2_host_logic(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert (tmp_path / f'ok{r}.npy').exists(), f'rank {r} did not finish'


def test_single_process_helpers_are_noops():
    from asac_b200 import dist as adist
    assert adist.world() == (0, 1)
    t = torch.ones(4)
    assert adist.all_reduce_sum_(t) is t and torch.equal(t, torch.ones(4))
    adist.broadcast_([t])
    assert adist.shard_capacity(524288, 1) == 524288 and adist.episode_owner(7, 1) == 0
