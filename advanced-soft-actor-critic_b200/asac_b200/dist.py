"""Multi-GPU plumbing of the learner (new capability; the reference picks ONE GPU,
algorithm/sac_base.py:237-243).  One process per GPU, ``torch.distributed`` over NCCL.

* replay: sharded by capacity — rank r owns ``capacity // world`` ring slots with its own
  segment tree; whole episodes go to one shard (round-robin) so sampled windows never cross
  shards.  No data-path collective: sampling, gather, priority update and write-back are local.
* learner: data parallel over the batch with replicated weights.  The only exchange is the
  gradient: after each of the three backward passes the reduced flat gradient buffer
  (critics / policy / log-alpha) is all-reduced (SUM) and the Adam kernel applies 1/world.

These helpers are pure host logic and run unchanged on the gloo backend (CPU tests).
"""
from __future__ import annotations

import math

import torch
import torch.distributed as dist


def world() -> tuple[int, int]:
    """(rank, world_size); (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_capacity(capacity: int, world_size: int) -> int:
    """Per-rank ring size: the global capacity rounded down to a power of two
    (replay_buffer.py:264), split evenly, each shard again a power of two."""
    total = int(2 ** math.floor(math.log2(capacity)))
    per = max(total // world_size, 2)
    return int(2 ** math.floor(math.log2(per)))


def episode_owner(episode_counter: int, world_size: int) -> int:
    """Round-robin shard of the episode_counter-th episode."""
    return episode_counter % world_size


def all_reduce_sum_(t: torch.Tensor, group=None) -> torch.Tensor:
    """In-place SUM all-reduce of a flat gradient buffer (a no-op without a process group)."""
    if world()[1] > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def broadcast_(tensors: list[torch.Tensor], src: int = 0, group=None) -> None:
    """Makes every rank start from rank ``src``'s weights / optimizer state."""
    if world()[1] > 1:
        for t in tensors:
            dist.broadcast(t, src=src, group=group)


def all_agree(flag: bool, device=None, group=None) -> bool:
    """True iff ``flag`` is true on EVERY rank (MIN all-reduce; the tensor lives on the CPU for gloo, on
    ``device`` for NCCL).  Used for decisions every rank must take identically: whether to step, which
    gradient exchange to run."""
    if world()[1] == 1:
        return bool(flag)
    on_cpu = dist.get_backend(group) == 'gloo' or device is None
    t = torch.tensor([1 if flag else 0], dtype=torch.int32, device='cpu' if on_cpu else device)
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    return bool(int(t.item()))


class PeerGradientExchange:
    """Symmetric NVLink-mapped receive buffer for the in-kernel gradient exchange
    (include/asac_b200.h, AsacPeerTable).  Built on ``torch.distributed._symmetric_memory`` (CUDA
    VMM handles exchanged through the process group's store): every rank allocates the buffer with
    the same size, the rendezvous maps all of them into every process.  Raises when the mapping
    cannot be established (the caller then keeps the NCCL all-reduce path)."""

    def __init__(self, recv_words: int, device: torch.device, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        group = dist.group.WORLD if group is None else group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.recv = symm_mem.empty(int(recv_words), dtype=torch.int64, device=device)
        self.recv.zero_()  # epoch 0 never matches: optimizer epochs start at 1
        torch.cuda.synchronize(device)
        self._hdl = symm_mem.rendezvous(self.recv, group.group_name)
        self.recv_ptrs = [int(p) for p in self._hdl.buffer_ptrs]
        if len(self.recv_ptrs) != self.world:
            raise RuntimeError('symmetric memory rendezvous returned an unexpected number of peers')
        dist.barrier(group)  # every rank's buffer is zeroed and mapped before the first kernel

    def table(self):
        from . import _lib
        t = _lib.AsacPeerTable()
        t.world, t.rank = self.world, self.rank
        for i in range(self.world):
            t.recv[i] = self.recv_ptrs[i]
        t.recv_words = self.recv.numel()
        return t
