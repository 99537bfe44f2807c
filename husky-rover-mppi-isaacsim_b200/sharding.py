"""Multi-GPU partitioning of the MPPI step (one process per GPU, torch.distributed for the plumbing).

Two modes, both new work (the reference is single-GPU):

* sample sharding (BASELINE config 3): rank g owns the contiguous global samples [k_begin, k_begin + K_local).
  The Philox counter is the GLOBAL sample id, so the noise -- and therefore every cost -- is independent of the
  number of ranks.  The only exchange is one all-gather of the softmax partial {M, S, argmin, S2, A1[T], A2[T]}
  (4 + 2T floats per rank, 816 B at T = 100); every rank then folds the partials in rank order with the same
  kernel, so all ranks hold the identical updated nominal without a broadcast.
* rover sharding (BASELINE config 4): rovers are independent controllers; rank g owns a contiguous block of
  rovers and nothing is exchanged.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced split of `total` units: returns (begin, count).  The first (total % world) ranks
    get one extra unit; empty shards are possible only when total < world."""
    if world < 1 or not (0 <= rank < world) or total < 0:
        raise ValueError("bad shard arguments")
    base, rem = divmod(total, world)
    begin = rank * base + min(rank, rem)
    return begin, base + (1 if rank < rem else 0)


def exchange_partials(mine: torch.Tensor, group=None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """All-gather of one softmax partial per rank -> [world, stride] in rank order.  Works on CUDA tensors
    (NCCL over NVLink) and on CPU tensors (gloo; used by the CPU tests of the protocol)."""
    world = dist.get_world_size(group)
    if out is None:
        out = torch.empty((world, mine.numel()), dtype=mine.dtype, device=mine.device)
    if mine.is_cuda:
        dist.all_gather_into_tensor(out, mine.reshape(-1), group=group)
    else:
        chunks = [out[r] for r in range(world)]
        dist.all_gather(chunks, mine.reshape(-1), group=group)
    return out


class SampleShardedStepper:
    """One logical controller with K_total samples spread over the ranks of `group`."""

    def __init__(self, core, K_total: int, group=None):
        self.core, self.group = core, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.k_begin, k_local = shard_range(K_total, self.world, self.rank)
        if k_local != core.K:
            raise ValueError(f"core was created for K={core.K} but this rank's shard has {k_local} samples")
        n = core.partial_floats()
        self.mine = torch.zeros(n, dtype=torch.float32, device=core.device)
        self.all = torch.zeros((self.world, n), dtype=torch.float32, device=core.device)

    def step(self, state, proj, seed: int, offset: int, noise=None, stream=None):
        self.core.step_partial(state, self.mine, self.k_begin, proj, noise, seed, offset, stream)
        if self.world > 1:
            exchange_partials(self.mine, self.group, self.all)
            parts = self.all
        else:
            parts = self.mine
        self.core.combine_partials(state, parts, self.world, stream)
